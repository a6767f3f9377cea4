// Globals of libsparse_b200.so: last-error string, launch counter, device info.
#include "common.h"

#include <mutex>

namespace sb200 {

thread_local char g_last_error[512] = {0};
std::atomic<unsigned long long> g_launches{0};

int num_sms() {
    static int cached[64];
    static std::once_flag once;
    std::call_once(once, [] {
        for (int& c : cached) c = 0;
    });
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

bool device_flag_test_and_set(int slot) {
    static std::atomic<unsigned long long> flags[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    const unsigned long long bit = 1ull << (slot & 63);
    return (flags[dev].fetch_or(bit) & bit) != 0;
}

}  // namespace sb200

extern "C" int sb200_abi_version(void) { return SB200_ABI_VERSION; }
extern "C" const char* sb200_last_error(void) { return sb200::g_last_error; }
extern "C" unsigned long long sb200_launch_count(void) { return sb200::g_launches.load(); }
