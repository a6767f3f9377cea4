// Issue rate of the legacy warp-level tensor path (mma.sync.m16n8k16 bf16 -> f32, SASS HMMA.16816.F32.BF16) on one
// B200: independent accumulator chains per warp, 1..16 warps per SM sub-partition. Prints cycles per HMMA per
// sub-partition and the implied dense TFLOP/s. Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hmma_rate tools/probes/hmma_rate.cu && /tmp/hmma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void hmma_kernel(int iters, float* out, long long* cycles) {
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                "{%0, %1, %2, %3};"
                : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CHAINS>
void run(int warps_per_sm, int sms) {
    const int iters = 4096;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * warps_per_sm * 32);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    hmma_kernel<CHAINS><<<sms, warps_per_sm * 32>>>(16, out, cyc);
    cudaEventRecord(e0);
    hmma_kernel<CHAINS><<<sms, warps_per_sm * 32>>>(iters, out, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[1];
    cudaMemcpy(h, cyc, sizeof(long long), cudaMemcpyDeviceToHost);
    const double per_smsp = double(warps_per_sm) / 4.0 * iters * CHAINS;   // HMMAs per sub-partition
    const double tflops = double(sms) * warps_per_sm * iters * CHAINS * 4096.0 / (ms * 1e-3) / 1e12;
    printf("chains %d  warps/SM %2d: %7.2f cycles per HMMA per sub-partition (SM clock)   %7.1f TFLOP/s  (%.3f ms)\n",
           CHAINS, warps_per_sm, double(h[0]) / per_smsp, tflops, ms);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    for (int w : {4, 8, 16, 32}) run<1>(w, sms);
    for (int w : {4, 8, 16, 32}) run<4>(w, sms);
    for (int w : {4, 8, 16, 32}) run<8>(w, sms);
    return 0;
}
