// Instance-norm statistics for the SPADE blocks (reference Spade_Norm_Block.forward, training/networks.py:4371-4379: nn.InstanceNorm2d(affine=False),
// biased variance, eps 1e-5).  The normalisation itself is applied inside the epilogue of the gamma|beta convolution (pg_conv2d_igemm_spade_run);
// this kernel only produces mean[n,c] and rstd[n,c] in ONE streaming pass over x (HBM-bound: 4 bytes per element).
#include "pg_common.cuh"

namespace pg {

constexpr int kStatThreads = 256;

// One CTA per (n, c) plane.  Shifted sums around the plane's first element K keep the single-pass variance well conditioned:
//   mean = K + S1 / n,  var = (S2 - S1^2 / n) / n  with  S1 = sum(x - K), S2 = sum((x - K)^2)
__global__ void __launch_bounds__(kStatThreads) instance_stats_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ rstd,
                                                                      long long hw, float eps, int vec) {
    const float* xp = x + (size_t)blockIdx.x * hw;
    const float K = __ldg(xp);
    float s1 = 0.f, s2 = 0.f;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(xp);
        const long long n4 = hw >> 2;
        for (long long i = threadIdx.x; i < n4; i += 8 * kStatThreads) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const long long j = i + (long long)u * kStatThreads; v[u] = j < n4 ? __ldg(x4 + j) : make_float4(K, K, K, K); }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float a = v[u].x - K, b = v[u].y - K, c = v[u].z - K, d = v[u].w - K;
                s1 += (a + b) + (c + d);
                s2 = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, s2))));
            }
        }
    } else {
        for (long long i = threadIdx.x; i < hw; i += kStatThreads) { const float a = __ldg(xp + i) - K; s1 += a; s2 = fmaf(a, a, s2); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    __shared__ float r1[kStatThreads / 32], r2[kStatThreads / 32];
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < kStatThreads / 32; i++) { t1 += r1[i]; t2 += r2[i]; }
        const float inv = 1.f / (float)hw;
        const float m = t1 * inv;
        const float var = fmaxf(fmaf(-m, t1, t2) * inv, 0.f);
        mean[blockIdx.x] = K + m;
        rstd[blockIdx.x] = rsqrtf(var + eps);
    }
}

}  // namespace pg

extern "C" int pg_instance_norm_stats(const float* x, float* mean, float* rstd, int64_t planes, int64_t hw, float eps, void* stream) {
    using namespace pg;
    PG_REQUIRE(planes >= 0 && hw >= 1 && planes <= INT32_MAX, "instance_norm_stats: bad sizes");
    if (planes == 0) return PG_OK;
    PG_REQUIRE(x && mean && rstd, "instance_norm_stats: x, mean and rstd must be device pointers");
    const int vec = (hw % 4 == 0) && aligned16(x);
    instance_stats_kernel<<<(unsigned)planes, kStatThreads, 0, (cudaStream_t)stream>>>(x, mean, rstd, (long long)hw, eps, vec);
    return launch_status("instance_norm_stats", 1);
}
