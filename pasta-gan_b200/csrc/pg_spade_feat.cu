// Garment-feature completion of SynthesisNetworkFull.get_spade_feat (reference training/networks.py:5777-5800) as two streaming kernels:
//   sum[n,c]        = sum_hw feat[n,c,hw] * valid[n,hw]                                  (pg_masked_plane_sum)
//   out[n,c0+c,hw]  = feat[n,c,hw] * (1 - rest[n,hw]) + fill[n,c] * rest[n,hw]           (pg_masked_fill, written into a channel slice of the
//                                                                                          concatenated upper|lower buffer: no torch.cat)
// replacing five elementwise / reduction passes over the 134 MB feature map and the concatenation copy.  HBM-bound: 4 B/elem and 8 B/elem.
#include "pg_common.cuh"

namespace pg {

constexpr int kSfThreads = 256;

__global__ void __launch_bounds__(kSfThreads) masked_plane_sum_kernel(const float* __restrict__ feat, const float* __restrict__ mask, float* __restrict__ out,
                                                                      int C, long long hw, int vec) {
    const long long plane = blockIdx.x;
    const float* fp = feat + (size_t)plane * hw;
    const float* mp = mask + (size_t)(plane / C) * hw;
    float s = 0.f;
    if (vec) {
        const float4* f4 = reinterpret_cast<const float4*>(fp);
        const float4* m4 = reinterpret_cast<const float4*>(mp);
        const long long n4 = hw >> 2;
        for (long long i = threadIdx.x; i < n4; i += 4 * kSfThreads) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const long long j = i + (long long)u * kSfThreads;
                a[u] = j < n4 ? __ldg(f4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                b[u] = j < n4 ? __ldg(m4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) s += (a[u].x * b[u].x + a[u].y * b[u].y) + (a[u].z * b[u].z + a[u].w * b[u].w);
        }
    } else {
        for (long long i = threadIdx.x; i < hw; i += kSfThreads) s = fmaf(__ldg(fp + i), __ldg(mp + i), s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ float red[kSfThreads / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < kSfThreads / 32; i++) t += red[i];
        out[plane] = t;
    }
}

template <bool HALF>
__global__ void __launch_bounds__(kSfThreads) masked_fill_kernel(const float* __restrict__ feat, const float* __restrict__ rest, const float* __restrict__ fill,
                                                                 float* __restrict__ out, int C, long long hw, long long out_batch_stride, int vec) {
    const long long plane = blockIdx.y;                       // n * C + c
    const int n = (int)(plane / C), c = (int)(plane - (long long)n * C);
    const float* fp = feat + (size_t)plane * hw;
    const float* rp = rest + (size_t)n * hw;
    float* op = HALF ? nullptr : out + (size_t)n * out_batch_stride + (size_t)c * hw;
    __half* oh = HALF ? reinterpret_cast<__half*>(out) + (size_t)n * out_batch_stride + (size_t)c * hw : nullptr;
    auto h2 = [](float a, float b) { unsigned int r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; };
    const float m = __ldg(fill + plane);
    if (vec) {
        const long long n4 = hw >> 2;
        for (long long i = (long long)blockIdx.x * kSfThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kSfThreads) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(fp) + i), r = __ldg(reinterpret_cast<const float4*>(rp) + i);
            float4 y;
            y.x = a.x * (1.f - r.x) + m * r.x; y.y = a.y * (1.f - r.y) + m * r.y;
            y.z = a.z * (1.f - r.z) + m * r.z; y.w = a.w * (1.f - r.w) + m * r.w;
            if (HALF) reinterpret_cast<uint2*>(oh)[i] = make_uint2(h2(y.x, y.y), h2(y.z, y.w));
            else      reinterpret_cast<float4*>(op)[i] = y;
        }
    } else {
        for (long long i = (long long)blockIdx.x * kSfThreads + threadIdx.x; i < hw; i += (long long)gridDim.x * kSfThreads) {
            const float r = __ldg(rp + i);
            const float y = __ldg(fp + i) * (1.f - r) + m * r;
            if (HALF) reinterpret_cast<unsigned short*>(oh)[i] = (unsigned short)(h2(y, 0.f) & 0xffffu);
            else      op[i] = y;
        }
    }
}

// The same fill written as channel-blocked fp16 (C8): one thread per (pixel, 8-channel block), eight coalesced plane reads, one 16-byte store into
// blocks [cb_offset, cb_offset + C / 8) of out [N][cb_total][hw][8].
__global__ void __launch_bounds__(kSfThreads) masked_fill_c8_kernel(const float* __restrict__ feat, const float* __restrict__ rest, const float* __restrict__ fill,
                                                                    uint4* __restrict__ out, int C, long long hw, int cb_total, int cb_offset) {
    const int CB = C / 8;
    const long long blk = blockIdx.y;                       // n * CB + cb
    const int n = (int)(blk / CB), cb = (int)(blk - (long long)n * CB);
    const float* fp = feat + ((size_t)n * C + (size_t)cb * 8) * hw;
    const float* rp = rest + (size_t)n * hw;
    uint4* op = out + ((size_t)n * cb_total + cb_offset + cb) * hw;
    auto h2 = [](float a, float b) { unsigned int r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; };
    float m[8];
#pragma unroll
    for (int c = 0; c < 8; c++) m[c] = __ldg(fill + (size_t)n * C + cb * 8 + c);
    for (long long i = (long long)blockIdx.x * kSfThreads + threadIdx.x; i < hw; i += (long long)gridDim.x * kSfThreads) {
        const float r = __ldg(rp + i);
        float y[8];
#pragma unroll
        for (int c = 0; c < 8; c++) y[c] = __ldg(fp + (size_t)c * hw + i) * (1.f - r) + m[c] * r;
        op[i] = make_uint4(h2(y[0], y[1]), h2(y[2], y[3]), h2(y[4], y[5]), h2(y[6], y[7]));
    }
}

}  // namespace pg

extern "C" int pg_masked_fill_c8(const float* feat, const float* rest, const float* fill, void* out, int64_t N, int64_t C, int64_t hw,
                                 int64_t cb_total, int64_t cb_offset, void* stream) {
    using namespace pg;
    PG_REQUIRE(N >= 0 && C >= 8 && C % 8 == 0 && hw >= 1 && cb_offset >= 0 && cb_offset + C / 8 <= cb_total, "masked_fill_c8: bad sizes (C %% 8 == 0)");
    if (N == 0) return PG_OK;
    PG_REQUIRE(feat && rest && fill && out && aligned16(out), "masked_fill_c8: feat, rest, fill and out must be device pointers, out 16-byte aligned");
    PG_REQUIRE(N * (C / 8) <= 65535, "masked_fill_c8: N * C / 8 must fit gridDim.y");
    long long gx = (hw + kSfThreads * 2 - 1) / (kSfThreads * 2);
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)(N * (C / 8)));
    masked_fill_c8_kernel<<<grid, kSfThreads, 0, (cudaStream_t)stream>>>(feat, rest, fill, (uint4*)out, (int)C, (long long)hw, (int)cb_total, (int)cb_offset);
    return launch_status("masked_fill_c8", 1);
}

extern "C" int pg_masked_plane_sum(const float* feat, const float* mask, float* out, int64_t N, int64_t C, int64_t hw, void* stream) {
    using namespace pg;
    PG_REQUIRE(N >= 0 && C >= 1 && hw >= 1 && N * C <= INT32_MAX, "masked_plane_sum: bad sizes");
    if (N == 0) return PG_OK;
    PG_REQUIRE(feat && mask && out, "masked_plane_sum: feat, mask and out must be device pointers");
    const int vec = (hw % 4 == 0) && aligned16(feat) && aligned16(mask);
    masked_plane_sum_kernel<<<(unsigned)(N * C), kSfThreads, 0, (cudaStream_t)stream>>>(feat, mask, out, (int)C, (long long)hw, vec);
    return launch_status("masked_plane_sum", 1);
}

extern "C" int pg_masked_fill(const float* feat, const float* rest, const float* fill, float* out, int64_t N, int64_t C, int64_t hw,
                              int64_t out_batch_stride, int32_t out_dtype, void* stream) {
    using namespace pg;
    PG_REQUIRE(N >= 0 && C >= 1 && hw >= 1 && N * C <= 65535 * 64LL && out_batch_stride >= C * hw, "masked_fill: bad sizes");
    if (N == 0) return PG_OK;
    PG_REQUIRE(feat && rest && fill && out, "masked_fill: feat, rest, fill and out must be device pointers");
    PG_REQUIRE(N * C <= 65535, "masked_fill: N * C must fit gridDim.y");
    PG_REQUIRE(out_dtype == PG_F32 || out_dtype == PG_F16, "masked_fill: out must be float32 or float16");
    const bool half = out_dtype == PG_F16;
    const int vec = (hw % 4 == 0) && (out_batch_stride % 4 == 0) && aligned16(feat) && aligned16(rest) && (half ? ((uintptr_t)out & 7) == 0 : aligned16(out));
    const long long per = vec ? hw / 4 : hw;
    unsigned gx = (unsigned)((per + kSfThreads * 4 - 1) / (kSfThreads * 4));
    if (gx < 1) gx = 1;
    dim3 grid(gx, (unsigned)(N * C));
    if (half) masked_fill_kernel<true><<<grid, kSfThreads, 0, (cudaStream_t)stream>>>(feat, rest, fill, out, (int)C, (long long)hw, (long long)out_batch_stride, vec);
    else      masked_fill_kernel<false><<<grid, kSfThreads, 0, (cudaStream_t)stream>>>(feat, rest, fill, out, (int)C, (long long)hw, (long long)out_batch_stride, vec);
    return launch_status("masked_fill", 1);
}
