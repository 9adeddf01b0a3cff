// Channel-blocked fp16 activations ("C8": [N][C/8][H*W][8], include/pasta_b200.h PG_LAYOUT_C8) <-> dense NCHW.
//
// The tcgen05 convolution reads a C8 tensor with the Tensor Memory Accelerator straight into its operand layout and its epilogue writes C8 directly,
// so inside a chain of our own layers nothing ever converts.  These two kernels are the boundary of such a chain: tensors produced by other code
// (fp32 NCHW, the reference's layout at every API boundary) enter through pg_nchw_to_c8, results leave through pg_c8_to_nchw.
// HBM-bound streaming: one thread per (pixel, 8-channel block) -- eight coalesced plane reads, one 16-byte store (and the reverse).
#include "pg_common.cuh"

namespace pg {

constexpr int kLyThreads = 256;

__device__ __forceinline__ unsigned int h2pack(float a, float b) {
    unsigned int r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

template <class T>
__global__ void __launch_bounds__(kLyThreads) nchw_to_c8_kernel(const T* __restrict__ x, uint4* __restrict__ y, int C, int CB, long long hw) {
    const long long blk = blockIdx.y;                      // n * CB + cb
    const int n = (int)(blk / CB), cb = (int)(blk - (long long)n * CB);
    const T* xp = x + ((size_t)n * C + (size_t)cb * 8) * hw;
    uint4* yp = y + (size_t)blk * hw;
    const int nval = C - cb * 8 < 8 ? C - cb * 8 : 8;
    for (long long i = (long long)blockIdx.x * kLyThreads + threadIdx.x; i < hw; i += (long long)gridDim.x * kLyThreads) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) v[c] = c < nval ? to_acc<T>(__ldg(xp + (size_t)c * hw + i)) : 0.f;
        yp[i] = make_uint4(h2pack(v[0], v[1]), h2pack(v[2], v[3]), h2pack(v[4], v[5]), h2pack(v[6], v[7]));
    }
}

template <class T>
__global__ void __launch_bounds__(kLyThreads) c8_to_nchw_kernel(const uint4* __restrict__ x, T* __restrict__ y, int C, int CB, long long hw) {
    const long long blk = blockIdx.y;
    const int n = (int)(blk / CB), cb = (int)(blk - (long long)n * CB);
    const uint4* xp = x + (size_t)blk * hw;
    T* yp = y + ((size_t)n * C + (size_t)cb * 8) * hw;
    const int nval = C - cb * 8 < 8 ? C - cb * 8 : 8;
    for (long long i = (long long)blockIdx.x * kLyThreads + threadIdx.x; i < hw; i += (long long)gridDim.x * kLyThreads) {
        const uint4 q = __ldg(xp + i);
        const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (c < nval) {
                const __half h = __ushort_as_half((unsigned short)((w[c >> 1] >> ((c & 1) * 16)) & 0xffffu));
                yp[(size_t)c * hw + i] = from_acc<T, float>(__half2float(h));
            }
        }
    }
}

static dim3 layout_grid(long long hw, long long blocks) {
    long long gx = (hw + kLyThreads * 4 - 1) / (kLyThreads * 4);
    if (gx < 1) gx = 1;
    if (gx > 4096) gx = 4096;
    return dim3((unsigned)gx, (unsigned)blocks);
}

}  // namespace pg

extern "C" int pg_nchw_to_c8(const void* x, void* y, int64_t N, int64_t C, int64_t hw, int32_t x_dtype, void* stream) {
    using namespace pg;
    const int64_t CB = (C + 7) / 8;
    PG_REQUIRE(N >= 0 && C >= 1 && hw >= 1 && N * CB <= 65535, "nchw_to_c8: bad sizes (N * ceil(C / 8) must fit gridDim.y)");
    PG_REQUIRE(x_dtype == PG_F32 || x_dtype == PG_F16, "nchw_to_c8: x must be float32 or float16");
    if (N == 0) return PG_OK;
    PG_REQUIRE(x && y && aligned16(y), "nchw_to_c8: x and y must be device pointers, y 16-byte aligned");
    const dim3 grid = layout_grid(hw, N * CB);
    if (x_dtype == PG_F32) nchw_to_c8_kernel<float><<<grid, kLyThreads, 0, (cudaStream_t)stream>>>((const float*)x, (uint4*)y, (int)C, (int)CB, (long long)hw);
    else                   nchw_to_c8_kernel<__half><<<grid, kLyThreads, 0, (cudaStream_t)stream>>>((const __half*)x, (uint4*)y, (int)C, (int)CB, (long long)hw);
    return launch_status("nchw_to_c8", 1);
}

extern "C" int pg_c8_to_nchw(const void* x, void* y, int64_t N, int64_t C, int64_t hw, int32_t y_dtype, void* stream) {
    using namespace pg;
    const int64_t CB = (C + 7) / 8;
    PG_REQUIRE(N >= 0 && C >= 1 && hw >= 1 && N * CB <= 65535, "c8_to_nchw: bad sizes (N * ceil(C / 8) must fit gridDim.y)");
    PG_REQUIRE(y_dtype == PG_F32 || y_dtype == PG_F16, "c8_to_nchw: y must be float32 or float16");
    if (N == 0) return PG_OK;
    PG_REQUIRE(x && y && aligned16(x), "c8_to_nchw: x and y must be device pointers, x 16-byte aligned");
    const dim3 grid = layout_grid(hw, N * CB);
    if (y_dtype == PG_F32) c8_to_nchw_kernel<float><<<grid, kLyThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (float*)y, (int)C, (int)CB, (long long)hw);
    else                   c8_to_nchw_kernel<__half><<<grid, kLyThreads, 0, (cudaStream_t)stream>>>((const uint4*)x, (__half*)y, (int)C, (int)CB, (long long)hw);
    return launch_status("c8_to_nchw", 1);
}
