// Error slot, ABI version and device check for libpasta_b200.so (include/pasta_b200.h).
#include <atomic>
#include "pg_common.cuh"

namespace pg {

char* error_slot() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_slot(), 512, fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

__global__ void probe_kernel(int* out) { if (out) *out = 100; }

}  // namespace pg

extern "C" int pg_abi_version(void) { return PG_ABI_VERSION; }

extern "C" int64_t pg_launch_count(void) { return (int64_t)pg::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* pg_last_error(void) { return pg::error_slot(); }

extern "C" int pg_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return pg::fail(PG_ERR_NO_DEVICE, "no CUDA device"); }
    cudaDeviceProp prop;
    PG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return pg::fail(PG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a images only", dev, prop.major, prop.minor);
    cudaFuncAttributes attr;
    cudaError_t e = cudaFuncGetAttributes(&attr, pg::probe_kernel);
    if (e != cudaSuccess) { cudaGetLastError(); return pg::fail(PG_ERR_NO_DEVICE, "sm_100a kernel image not loadable: %s", cudaGetErrorString(e)); }
    return PG_OK;
}
