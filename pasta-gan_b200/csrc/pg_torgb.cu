// ToRGB skip path for sm_100a: img_out = upsample2d(img_in, f) + clamp( sum_c W[o,c] * s[n,c] * x[n,c] + b[o] ).
//
// Replaces, in one streaming pass over the feature map, what the reference does with four operator calls per synthesis block
// (training/networks.py:5709-5715 + ToRGBLayerFull.forward :5601-5611): upfirdn2d.upsample2d on the running RGB image (up = 2, 4x4 FIR,
// pad (2,1,2,1), gain 4 — upfirdn2d.py:308-343), modulated 1x1 conv without demodulation (:37-94), bias_act(linear, clamp) and img.add_(y).
// The op is HBM-bound: it reads the C-channel feature map once (268 MB at 256^2, batch 16) to produce 3 (RGB) or 6 (parsing) channels.
//   * one thread = 4 consecutive pixels (float4 loads along w, a warp covers 512 contiguous bytes per channel) x all O outputs;
//   * per-sample modulated weights W[o,c] * s[n,c] are built once per CTA in shared memory;
//   * the channel loop is unrolled 8x so 8 independent 16-byte loads are in flight per thread;
//   * the 2x2 polyphase taps of the up-sampled previous image and the residual add live in the epilogue.
#include "pg_common.cuh"

namespace pg {

constexpr int kRgbMaxO = 8;

struct ToRgbParams {
    const float* x; const float* w; const float* styles; const float* bias; const float* img_in; const float* fir; float* out;
    int N, C, O, H, W; float clamp; int has_img; int x_c8;
};

// bias + clamp, the 2x2 live taps of upsample2d(img_in) and the store of one quad of 4 adjacent pixels, all O output channels
template <int O>
__device__ __forceinline__ void torgb_epilogue_quad(const ToRgbParams& p, const int n, const int qd, const float (&acc)[O][4]) {
    const int HW = p.H * p.W;
    const int pix = qd << 2;
    const int Y = pix / p.W, X0 = pix - Y * p.W;
#pragma unroll
    for (int o = 0; o < O; o++) {
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float v = acc[o][k] + (p.bias ? __ldg(p.bias + o) : 0.f);
            if (p.clamp >= 0.f) v = fminf(fmaxf(v, -p.clamp), p.clamp);
            r[k] = v;
        }
        if (p.has_img) {
            // upsample2d(img_in): zero-insert x2, pad (2,1,2,1), 4x4 FIR as a true convolution, gain 4  ->  2x2 live taps per output
            const int h2 = p.H >> 1, w2 = p.W >> 1;
            const float* im = p.img_in + ((size_t)n * O + o) * h2 * w2;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int X = X0 + k;
                const int uy0 = Y - 2, ux0 = X - 2;
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if ((uy0 + i) & 1) continue;
                    const int iy = (uy0 + i) >> 1;
                    if (iy < 0 || iy >= h2) continue;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if ((ux0 + j) & 1) continue;
                        const int ix = (ux0 + j) >> 1;
                        if (ix < 0 || ix >= w2) continue;
                        s = fmaf(__ldg(p.fir + (3 - i) * 4 + (3 - j)), __ldg(im + iy * w2 + ix), s);
                    }
                }
                r[k] += 4.f * s;
            }
        }
        *reinterpret_cast<float4*>(p.out + ((size_t)n * O + o) * HW + pix) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

// Small images (<= 64^2): the kernel above gives each thread a quad of pixels and the whole channel loop -- at 4^2 that is four busy threads per
// sample walking 512 channels one memory round trip after another (~70 us whatever the size).  Here the eight warps of a CTA split the channels
// (warp w takes c = w, w + 8, ...), a CTA covers 32 quads, and the partial sums meet in shared memory: 8x shorter dependent chains, 8x the CTAs.
template <int O>
__global__ void __launch_bounds__(256) torgb_skip_splitc_kernel(ToRgbParams p) {
    extern __shared__ float wmod[];                         // [C][O] for this sample, then the partial sums [8 warps][32 lanes][O * 4]
    float* red = wmod + p.C * O;
    const int n = blockIdx.y;
    const int HW = p.H * p.W;
    for (int i = threadIdx.x; i < p.C * O; i += blockDim.x) {
        const int c = i / O, o = i - c * O;
        wmod[i] = p.w[o * p.C + c] * (p.styles ? p.styles[(size_t)n * p.C + c] : 1.f);
    }
    __syncthreads();
    const int quads = HW >> 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qd = blockIdx.x * 32 + lane;
    const bool live = qd < quads;
    const float4* src = reinterpret_cast<const float4*>(p.x + (size_t)n * p.C * HW) + (live ? qd : 0);
    float acc[O][4];
#pragma unroll
    for (int o = 0; o < O; o++) { acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f; }
    for (int c0 = warp; c0 < p.C; c0 += 32) {               // four channels of this warp (stride 8) in flight
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = (live && c0 + 8 * u < p.C) ? __ldg(src + (size_t)(c0 + 8 * u) * quads) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (c0 + 8 * u >= p.C) break;
#pragma unroll
            for (int o = 0; o < O; o++) {
                const float wv = wmod[(c0 + 8 * u) * O + o];
                acc[o][0] = fmaf(wv, v[u].x, acc[o][0]); acc[o][1] = fmaf(wv, v[u].y, acc[o][1]);
                acc[o][2] = fmaf(wv, v[u].z, acc[o][2]); acc[o][3] = fmaf(wv, v[u].w, acc[o][3]);
            }
        }
    }
    float* mine = red + ((size_t)warp * 32 + lane) * (O * 4);
#pragma unroll
    for (int o = 0; o < O; o++)
#pragma unroll
        for (int k = 0; k < 4; k++) mine[o * 4 + k] = acc[o][k];
    __syncthreads();
    if (warp == 0 && live) {
#pragma unroll
        for (int o = 0; o < O; o++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float t = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; w8++) t += red[((size_t)w8 * 32 + lane) * (O * 4) + o * 4 + k];      // fixed order: deterministic
                acc[o][k] = t;
            }
        torgb_epilogue_quad<O>(p, n, qd, acc);
    }
}

template <int O>
__global__ void __launch_bounds__(256) torgb_skip_kernel(ToRgbParams p) {
    extern __shared__ float wmod[];                         // [C][O] for this sample
    const int n = blockIdx.y;
    const int HW = p.H * p.W;
    for (int i = threadIdx.x; i < p.C * O; i += blockDim.x) {
        const int c = i / O, o = i - c * O;
        wmod[i] = p.w[o * p.C + c] * (p.styles ? p.styles[(size_t)n * p.C + c] : 1.f);
    }
    __syncthreads();
    const int quads = HW >> 2;                              // H*W is a multiple of 4 (checked on the host)
    const float* xn = p.x + (size_t)n * p.C * HW;
    for (int qd = blockIdx.x * blockDim.x + threadIdx.x; qd < quads; qd += gridDim.x * blockDim.x) {
        float acc[O][4];
#pragma unroll
        for (int o = 0; o < O; o++) { acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f; }
        const float4* src = reinterpret_cast<const float4*>(xn) + qd;
        int c = 0;
        for (; c + 8 <= p.C; c += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = __ldg(src + (size_t)(c + u) * quads);
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int o = 0; o < O; o++) {
                    const float wv = wmod[(c + u) * O + o];
                    acc[o][0] = fmaf(wv, v[u].x, acc[o][0]); acc[o][1] = fmaf(wv, v[u].y, acc[o][1]);
                    acc[o][2] = fmaf(wv, v[u].z, acc[o][2]); acc[o][3] = fmaf(wv, v[u].w, acc[o][3]);
                }
            }
        }
        for (; c < p.C; c++) {
            const float4 v = __ldg(src + (size_t)c * quads);
#pragma unroll
            for (int o = 0; o < O; o++) {
                const float wv = wmod[c * O + o];
                acc[o][0] = fmaf(wv, v.x, acc[o][0]); acc[o][1] = fmaf(wv, v.y, acc[o][1]);
                acc[o][2] = fmaf(wv, v.z, acc[o][2]); acc[o][3] = fmaf(wv, v.w, acc[o][3]);
            }
        }
        torgb_epilogue_quad<O>(p, n, qd, acc);
    }
}

// The same operation over a channel-blocked fp16 feature map (PG_LAYOUT_C8, [N][C/8][H*W][8]): one thread = one pixel, a 16-byte load per channel
// block (a warp reads 512 contiguous bytes), four blocks in flight; half the bytes of the fp32 form.
template <int O>
__global__ void __launch_bounds__(256) torgb_skip_c8_kernel(ToRgbParams p) {
    extern __shared__ float wmod[];                         // [C][O] for this sample
    const int n = blockIdx.y;
    const int HW = p.H * p.W;
    for (int i = threadIdx.x; i < p.C * O; i += blockDim.x) {
        const int c = i / O, o = i - c * O;
        wmod[i] = p.w[o * p.C + c] * (p.styles ? p.styles[(size_t)n * p.C + c] : 1.f);
    }
    __syncthreads();
    const int CB = p.C >> 3;
    const uint4* xn = reinterpret_cast<const uint4*>(p.x) + (size_t)n * CB * HW;
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
        float acc[O];
#pragma unroll
        for (int o = 0; o < O; o++) acc[o] = 0.f;
        auto fma8 = [&](const uint4& q, int cb) {
            const unsigned int wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wds[j]));
                const float* w0 = wmod + (cb * 8 + 2 * j) * O;
#pragma unroll
                for (int o = 0; o < O; o++) acc[o] = fmaf(w0[O + o], f.y, fmaf(w0[o], f.x, acc[o]));
            }
        };
        int cb = 0;
        for (; cb + 4 <= CB; cb += 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = __ldg(xn + (size_t)(cb + u) * HW + pix);
#pragma unroll
            for (int u = 0; u < 4; u++) fma8(v[u], cb + u);
        }
        for (; cb < CB; cb++) fma8(__ldg(xn + (size_t)cb * HW + pix), cb);
        const int Y = pix / p.W, X = pix - Y * p.W;
#pragma unroll
        for (int o = 0; o < O; o++) {
            float v = acc[o] + (p.bias ? __ldg(p.bias + o) : 0.f);
            if (p.clamp >= 0.f) v = fminf(fmaxf(v, -p.clamp), p.clamp);
            if (p.has_img) {
                const int h2 = p.H >> 1, w2 = p.W >> 1;
                const float* im = p.img_in + ((size_t)n * O + o) * h2 * w2;
                const int uy0 = Y - 2, ux0 = X - 2;
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if ((uy0 + i) & 1) continue;
                    const int iy = (uy0 + i) >> 1;
                    if (iy < 0 || iy >= h2) continue;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if ((ux0 + j) & 1) continue;
                        const int ix = (ux0 + j) >> 1;
                        if (ix < 0 || ix >= w2) continue;
                        s = fmaf(__ldg(p.fir + (3 - i) * 4 + (3 - j)), __ldg(im + iy * w2 + ix), s);
                    }
                }
                v += 4.f * s;
            }
            p.out[((size_t)n * O + o) * HW + pix] = v;
        }
    }
}

template <int O>
static int launch_torgb(const ToRgbParams& p, cudaStream_t s) {
    if (p.x_c8) {
        const size_t smem = (size_t)p.C * O * sizeof(float);
        const int HW = p.H * p.W;
        int bx = (HW + 255) / 256;
        if (bx > kNumSMs * 8) bx = kNumSMs * 8;
        if (smem > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(torgb_skip_c8_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        torgb_skip_c8_kernel<O><<<dim3(bx, p.N), 256, smem, s>>>(p);
        return launch_status("torgb_skip(c8)");
    }
    const size_t smem = (size_t)p.C * O * sizeof(float);
    const int quads = p.H * p.W / 4;
    if (quads <= 1024) {                                   // <= 64^2: channel-split kernel
        const size_t smem2 = smem + (size_t)8 * 32 * O * 4 * sizeof(float);
        if (smem2 > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(torgb_skip_splitc_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        torgb_skip_splitc_kernel<O><<<dim3((quads + 31) / 32, p.N), 256, smem2, s>>>(p);
        return launch_status("torgb_skip(split-c)");
    }
    int bx = (quads + 255) / 256;
    if (bx > kNumSMs * 8) bx = kNumSMs * 8;
    if (smem > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(torgb_skip_kernel<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    torgb_skip_kernel<O><<<dim3(bx, p.N), 256, smem, s>>>(p);
    return launch_status("torgb_skip");
}

}  // namespace pg

static int torgb_entry(const float* x, const float* w, const float* styles, const float* bias, const float* img_in, const float* fir,
                       float* out, int32_t N, int32_t C, int32_t O, int32_t H, int32_t W, float clamp, int x_c8, void* stream) {
    using namespace pg;
    PG_REQUIRE(!x_c8 || C % 8 == 0, "torgb_skip: a channel-blocked input needs C %% 8 == 0");
    PG_REQUIRE(N >= 0 && C >= 1 && H >= 1 && W >= 1, "torgb_skip: bad sizes");
    PG_REQUIRE(O >= 1 && O <= kRgbMaxO, "torgb_skip: 1..%d output channels supported (got %d)", kRgbMaxO, O);
    PG_REQUIRE((H * W) % 4 == 0 && W % 4 == 0, "torgb_skip: W must be a multiple of 4");
    PG_REQUIRE(img_in == nullptr || (fir != nullptr && H % 2 == 0 && W % 2 == 0), "torgb_skip: the skip image needs the 4x4 FIR and even H, W");
    PG_REQUIRE((int64_t)N * C * H * W <= INT32_MAX, "torgb_skip: x is too large");
    if (N == 0) return PG_OK;
    PG_REQUIRE(x && w && out, "torgb_skip: x, w and out must be device pointers");
    PG_REQUIRE(aligned16(x) && aligned16(out), "torgb_skip: x and out must be 16-byte aligned");
    ToRgbParams p;
    p.x = x; p.w = w; p.styles = styles; p.bias = bias; p.img_in = img_in; p.fir = fir; p.out = out;
    p.N = N; p.C = C; p.O = O; p.H = H; p.W = W; p.clamp = clamp; p.has_img = img_in != nullptr; p.x_c8 = x_c8;
    cudaStream_t s = (cudaStream_t)stream;
    switch (O) {
        case 1: return launch_torgb<1>(p, s); case 2: return launch_torgb<2>(p, s); case 3: return launch_torgb<3>(p, s);
        case 4: return launch_torgb<4>(p, s); case 5: return launch_torgb<5>(p, s); case 6: return launch_torgb<6>(p, s);
        case 7: return launch_torgb<7>(p, s); case 8: return launch_torgb<8>(p, s);
    }
    return fail(PG_ERR_UNSUPPORTED, "torgb_skip: O=%d", O);
}

extern "C" int pg_torgb_skip(const float* x, const float* w, const float* styles, const float* bias, const float* img_in, const float* fir,
                             float* out, int32_t N, int32_t C, int32_t O, int32_t H, int32_t W, float clamp, void* stream) {
    return torgb_entry(x, w, styles, bias, img_in, fir, out, N, C, O, H, W, clamp, 0, stream);
}

extern "C" int pg_torgb_skip_c8(const void* x_c8, const float* w, const float* styles, const float* bias, const float* img_in, const float* fir,
                                float* out, int32_t N, int32_t C, int32_t O, int32_t H, int32_t W, float clamp, void* stream) {
    return torgb_entry((const float*)x_c8, w, styles, bias, img_in, fir, out, N, C, O, H, W, clamp, 1, stream);
}
