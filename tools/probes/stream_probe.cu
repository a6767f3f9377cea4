// Micro-benchmark behind DESIGN.md 4.3: how fast can 148 SMs stream a [Nd, V] fp32 matrix under different work
// assignments and copy mechanisms? Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// A: one row per CTA at a time (rows round-robin), plain 16/8-byte loads into registers, `U` loads in flight per thread
template <int U>
__global__ void __launch_bounds__(1024) row_per_cta_ldg(const float* __restrict__ d, int Nd, int V, int contiguous, float* out) {
    float acc = 0.f;
    const int rows_per = (Nd + gridDim.x - 1) / gridDim.x;
    for (int r = 0; r < rows_per; ++r) {
        const int j = contiguous ? blockIdx.x * rows_per + r : blockIdx.x + r * gridDim.x;
        if (j >= Nd) break;
        const float2* row = reinterpret_cast<const float2*>(d + size_t(j) * V);
        const int n2 = V / 2;
        for (int i0 = threadIdx.x; i0 < n2; i0 += blockDim.x * U) {
            float2 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = (i0 + u * blockDim.x < n2) ? __ldg(row + i0 + u * blockDim.x) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += x[u].x + x[u].y;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

// C: the whole grid walks the matrix linearly (torch.sum-like)
template <int U>
__global__ void __launch_bounds__(512) linear_ldg(const float4* __restrict__ d, size_t n4, float* out) {
    float acc = 0.f;
    const size_t stride = size_t(gridDim.x) * blockDim.x;
    for (size_t i0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = (i0 + u * stride < n4) ? __ldg(d + i0 + u * stride) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += x[u].x + x[u].y + x[u].z + x[u].w;
    }
    if (acc == 123.456f) out[0] = acc;
}

// E: row parts through shared memory with cp.async (what scores_docrow_kernel does), S stages
template <int S>
__global__ void __launch_bounds__(512, 1) row_per_cta_cpasync(const float* __restrict__ d, int Nd, int V, int parts, float* out) {
    extern __shared__ __align__(128) float st[];
    const int part_len = ((V + parts - 1) / parts + 3) & ~3;
    const int stage_floats = part_len + 8;
    const int my_rows = (Nd - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
    const int total = my_rows * parts;
    auto issue = [&](int t) {
        const int j = blockIdx.x + (t / parts) * gridDim.x, p = t % parts;
        const int c0 = p * part_len, len = min(V, c0 + part_len) - c0;
        const float* src = d + size_t(j) * V + c0;
        const int a = int(reinterpret_cast<uintptr_t>(src) & 15) >> 2;
        const int head = min(len, (4 - a) & 3);
        const int n16 = (len - head) >> 2;
        float* stage = st + (t % S) * stage_floats;
        const unsigned dst0 = unsigned(__cvta_generic_to_shared(stage + a + head));
        for (int i = threadIdx.x; i < n16; i += 512)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + unsigned(i) * 16u), "l"(src + head + 4 * i) : "memory");
    };
    for (int t = 0; t < S - 1; ++t) { if (t < total) issue(t); asm volatile("cp.async.commit_group;" ::: "memory"); }
    float acc = 0.f;
    for (int t = 0; t < total; ++t) {
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 2) : "memory");
        __syncthreads();
        if (t + S - 1 < total) issue(t + S - 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        acc += st[(t % S) * stage_floats + 8 + (threadIdx.x & 127)];
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float timeit(F f, float* flush, size_t flush_n) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f, tot = 0.f;
    for (int it = 0; it < 8; ++it) {
        linear_ldg<4><<<1184, 512>>>(reinterpret_cast<const float4*>(flush), flush_n / 4, flush);   // clean-L2 flush (read)
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) { tot += ms; if (ms < best) best = ms; }
    }
    return tot / 6;
}

int main() {
    const int Nd = 2048, V = 30522;
    float *d, *out, *flush;
    const size_t n = size_t(Nd) * V, flush_n = size_t(64) << 20;
    CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&flush, flush_n * 4));
    CK(cudaMemset(d, 0, n * 4)); CK(cudaMemset(flush, 0, flush_n * 4));
    const double gb = double(n) * 4 / 1e9;
    auto rep = [&](const char* name, float ms) { printf("%-58s %8.1f us %8.1f GB/s  %.3f\n", name, ms * 1e3, gb / (ms / 1e3), gb / (ms / 1e3) / 6547.8); };
    rep("C linear grid-stride float4 x4, 1184x512", timeit([&] { linear_ldg<4><<<1184, 512>>>(reinterpret_cast<const float4*>(d), n / 4, out); }, flush, flush_n));
    rep("C linear grid-stride float4 x8, 592x512", timeit([&] { linear_ldg<8><<<592, 512>>>(reinterpret_cast<const float4*>(d), n / 4, out); }, flush, flush_n));
    rep("A row/CTA round-robin, 148x512, 8 loads", timeit([&] { row_per_cta_ldg<8><<<148, 512>>>(d, Nd, V, 0, out); }, flush, flush_n));
    rep("A row/CTA round-robin, 148x1024, 8 loads", timeit([&] { row_per_cta_ldg<8><<<148, 1024>>>(d, Nd, V, 0, out); }, flush, flush_n));
    rep("A row/CTA round-robin, 148x1024, 16 loads", timeit([&] { row_per_cta_ldg<16><<<148, 1024>>>(d, Nd, V, 0, out); }, flush, flush_n));
    rep("A row/CTA round-robin, 296x512, 8 loads", timeit([&] { row_per_cta_ldg<8><<<296, 512>>>(d, Nd, V, 0, out); }, flush, flush_n));
    rep("A row/CTA round-robin, 592x512, 8 loads", timeit([&] { row_per_cta_ldg<8><<<592, 512>>>(d, Nd, V, 0, out); }, flush, flush_n));
    rep("A row/CTA round-robin, 2048x512 (one row each)", timeit([&] { row_per_cta_ldg<8><<<2048, 512>>>(d, Nd, V, 0, out); }, flush, flush_n));
    rep("D row/CTA contiguous blocks, 148x512, 8 loads", timeit([&] { row_per_cta_ldg<8><<<148, 512>>>(d, Nd, V, 1, out); }, flush, flush_n));
    rep("D row/CTA contiguous blocks, 148x1024, 16 loads", timeit([&] { row_per_cta_ldg<16><<<148, 1024>>>(d, Nd, V, 1, out); }, flush, flush_n));
    CK(cudaFuncSetAttribute(row_per_cta_cpasync<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(row_per_cta_cpasync<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(row_per_cta_cpasync<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    auto smem = [&](int parts, int S) { return size_t(S) * ((((V + parts - 1) / parts + 3) & ~3) + 8) * 4; };
    rep("E cp.async 2 parts x 3 stages, 148x512 (kernel as shipped)", timeit([&] { row_per_cta_cpasync<3><<<148, 512, smem(2, 3)>>>(d, Nd, V, 2, out); }, flush, flush_n));
    rep("E cp.async 4 parts x 6 stages, 148x512", timeit([&] { row_per_cta_cpasync<6><<<148, 512, smem(4, 6)>>>(d, Nd, V, 4, out); }, flush, flush_n));
    rep("E cp.async 8 parts x 6 stages, 148x512", timeit([&] { row_per_cta_cpasync<6><<<148, 512, smem(8, 6)>>>(d, Nd, V, 8, out); }, flush, flush_n));
    rep("E cp.async 4 parts x 3 stages, 296x512 (2 CTAs/SM)", timeit([&] { row_per_cta_cpasync<3><<<296, 512, smem(4, 3)>>>(d, Nd, V, 4, out); }, flush, flush_n));
    rep("E cp.async 8 parts x 3 stages, 592x512 (4 CTAs/SM)", timeit([&] { row_per_cta_cpasync<3><<<592, 512, smem(8, 3)>>>(d, Nd, V, 8, out); }, flush, flush_n));
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
