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

// nn.InstanceNorm2d(affine=False) followed by an activation, y = act((x - mean) * rstd) * gain, for planes that fit shared memory (<= 24 K elements):
// ONE pass over HBM -- the plane is staged in shared memory, reduced (same shifted sums as above), normalised and written.  This is the
// Linear -> InstanceNorm -> LeakyReLU tail of the style encoder's `Dense` blocks (reference training/networks.py:594-611), which torch runs as three
// passes (batch_norm_collect_statistics, batch_norm_transform_input, leaky_relu).
__global__ void __launch_bounds__(kStatThreads) instance_norm_act_kernel(const float* __restrict__ x, float* __restrict__ y, int hw, float eps,
                                                                         float slope, float gain, int vec) {
    extern __shared__ float plane[];
    const float* xp = x + (size_t)blockIdx.x * hw;
    float* yp = y + (size_t)blockIdx.x * hw;
    const float K = __ldg(xp);
    float s1 = 0.f, s2 = 0.f;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(xp);
        float4* p4 = reinterpret_cast<float4*>(plane);
        for (int i = threadIdx.x; i < (hw >> 2); i += kStatThreads) {
            const float4 v = __ldg(x4 + i);
            p4[i] = v;
            const float a = v.x - K, b = v.y - K, c = v.z - K, d = v.w - K;
            s1 += (a + b) + (c + d);
            s2 = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, s2))));
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kStatThreads) { const float v = __ldg(xp + i); plane[i] = v; const float a = v - K; s1 += a; s2 = fmaf(a, a, s2); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    __shared__ float r1[kStatThreads / 32], r2[kStatThreads / 32], stat[2];
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < kStatThreads / 32; i++) { t1 += r1[i]; t2 += r2[i]; }
        const float inv = 1.f / (float)hw;
        const float m = t1 * inv;
        const float var = fmaxf(fmaf(-m, t1, t2) * inv, 0.f);
        stat[0] = K + m; stat[1] = rsqrtf(var + eps);
    }
    __syncthreads();
    const float mean = stat[0], rstd = stat[1];
    auto f = [&](float v) { const float t = (v - mean) * rstd; return (fmaxf(t, 0.f) + slope * fminf(t, 0.f)) * gain; };
    if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(plane);
        float4* y4 = reinterpret_cast<float4*>(yp);
        for (int i = threadIdx.x; i < (hw >> 2); i += kStatThreads) { const float4 v = p4[i]; y4[i] = make_float4(f(v.x), f(v.y), f(v.z), f(v.w)); }
    } else {
        for (int i = threadIdx.x; i < hw; i += kStatThreads) yp[i] = f(plane[i]);
    }
}

}  // namespace pg

extern "C" int pg_instance_norm_act(const float* x, float* y, int64_t planes, int64_t hw, float eps, int32_t act, float alpha, float gain, void* stream) {
    using namespace pg;
    PG_REQUIRE(planes >= 0 && hw >= 1 && planes <= INT32_MAX, "instance_norm_act: bad sizes");
    PG_REQUIRE(hw <= 24 * 1024, "instance_norm_act: planes of more than 24576 elements do not fit shared memory (use pg_instance_norm_stats + the consumer's epilogue)");
    PG_REQUIRE(act == PG_ACT_LINEAR || act == PG_ACT_RELU || act == PG_ACT_LRELU, "instance_norm_act: act must be linear / relu / lrelu");
    if (planes == 0) return PG_OK;
    PG_REQUIRE(x && y, "instance_norm_act: x and y must be device pointers");
    const int vec = (hw % 4 == 0) && aligned16(x) && aligned16(y);
    const float slope = act == PG_ACT_LINEAR ? 1.f : (act == PG_ACT_RELU ? 0.f : alpha);
    const size_t smem = (size_t)hw * sizeof(float);
    if (smem > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(instance_norm_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    instance_norm_act_kernel<<<(unsigned)planes, kStatThreads, smem, (cudaStream_t)stream>>>(x, y, (int)hw, eps, slope, gain, vec);
    return launch_status("instance_norm_act", 1);
}

extern "C" int pg_instance_norm_stats(const float* x, float* mean, float* rstd, int64_t planes, int64_t hw, float eps, void* stream) {
    using namespace pg;
    PG_REQUIRE(planes >= 0 && hw >= 1 && planes <= INT32_MAX, "instance_norm_stats: bad sizes");
    if (planes == 0) return PG_OK;
    PG_REQUIRE(x && mean && rstd, "instance_norm_stats: x, mean and rstd must be device pointers");
    const int vec = (hw % 4 == 0) && aligned16(x);
    instance_stats_kernel<<<(unsigned)planes, kStatThreads, 0, (cudaStream_t)stream>>>(x, mean, rstd, (long long)hw, eps, vec);
    return launch_status("instance_norm_stats", 1);
}
