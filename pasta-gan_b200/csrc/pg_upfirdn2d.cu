// upfirdn2d for sm_100a: pad -> zero-insert upsample -> 2-D FIR -> decimate, optionally with the
// bias_act epilogue of the calling layer fused in.
//
// Replaces the reference's upfirdn2d_kernel_small / upfirdn2d_kernel_large
// (torch_utils/ops/upfirdn2d.cu:29-200, launcher upfirdn2d.cpp:16-94).  Two kernels:
//
//   band kernel     NCHW-contiguous planes, up = 1, down in {1, 2}, 4x4 filter, non-negative pad.
//                   This is >95 % of the bytes PASTA-GAN moves through upfirdn2d (the (2H+1)^2 ->
//                   (2H)^2 filter after every up-sampling conv, the pad-2 filter before every
//                   stride-2 conv, the down-2 skip filter, and their backward forms).  One CTA owns a
//                   full-width band of output rows of one plane: because rows are full width, the
//                   input rows it needs are ONE contiguous span of global memory regardless of the odd
//                   (2H+1) row length, so loads are perfectly coalesced; the span is staged in shared
//                   memory once (halo rows re-read from L2 only), each thread then slides a 4-row
//                   register window down one output column (16 taps, 4 shared loads per output), and
//                   stores are coalesced along the row.
//   generic kernel  any filter size, any up/down, negative padding, flips, any strides
//                   (channels_last included): one thread per output element gathers its taps through
//                   the read-only path.  Serves AugmentPipe-style filters, the 3-channel RGB up-2 and
//                   every tiny plane (< 32 px wide) where launch latency, not bandwidth, is the cost.
#include <limits.h>
#include "pg_common.cuh"

namespace pg {

struct Epilogue {           // y = clamp(act(v + b[c]) * act_gain)
    const void* b; int act; float alpha, act_gain, clamp; int enabled;
};

template <class S>
__device__ __forceinline__ S apply_epilogue(const Epilogue& e, S v, S bias) {
    v += bias;
    if (e.act == PG_ACT_RELU)  v = v > (S)0 ? v : (S)0;
    if (e.act == PG_ACT_LRELU) v = v > (S)0 ? v : v * (S)e.alpha;
    v *= (S)e.act_gain;
    if (e.clamp >= 0.f) { const S c = (S)e.clamp; v = v > c ? c : (v < -c ? -c : v); }
    return v;
}

// packed fp32 pairs (FFMA2 / FMUL2 on sm_100a): two lanes of fp32 math per issue slot -- the band kernels are issue-bound, not FMA-pipe-bound
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}

// Epilogue constants of a band CTA: y = clamp(act(v) * gain) with the bias already inside v.  For slopes in [0, 1] (relu, lrelu) and a positive
// gain, act(v) * gain == max(v * gain, v * slope * gain): two packed multiplies and a max per pair instead of compare / select / multiply.
struct BandEpilogue {
    float2 g2, sg2; float clamp; int mode;        // mode 0: nothing, 1: v * gain, 2: max form, 3: general select form
    float alpha, gain;
};
__device__ __forceinline__ BandEpilogue make_band_epilogue(const Epilogue& e) {
    BandEpilogue b;
    const float slope = e.act == PG_ACT_RELU ? 0.f : (e.act == PG_ACT_LRELU ? e.alpha : 1.f);
    b.alpha = slope; b.gain = e.act_gain; b.clamp = e.clamp;
    b.g2 = make_float2(e.act_gain, e.act_gain); b.sg2 = make_float2(slope * e.act_gain, slope * e.act_gain);
    if (e.act == PG_ACT_LINEAR) b.mode = e.act_gain == 1.f ? 0 : 1;
    else b.mode = (slope >= 0.f && slope <= 1.f && e.act_gain > 0.f) ? 2 : 3;
    return b;
}
__device__ __forceinline__ float2 band_act(const BandEpilogue& b, float2 v) {
    if (b.mode == 2) { const float2 t = f2mul(v, b.g2), u = f2mul(v, b.sg2); v = make_float2(fmaxf(t.x, u.x), fmaxf(t.y, u.y)); }
    else if (b.mode == 1) v = f2mul(v, b.g2);
    else if (b.mode == 3) { v.x = (v.x > 0.f ? v.x : v.x * b.alpha) * b.gain; v.y = (v.y > 0.f ? v.y : v.y * b.alpha) * b.gain; }
    if (b.clamp >= 0.f) { v.x = fminf(fmaxf(v.x, -b.clamp), b.clamp); v.y = fminf(fmaxf(v.y, -b.clamp), b.clamp); }
    return v;
}

struct UpfirdnParams {
    const void* x; const float* f; void* y;
    int N, C, inH, inW, outH, outW;
    int64_t xs[4], ys[4];          // element strides {n, c, h, w}
    int fh, fw; int64_t fsh, fsw;
    int upx, upy, downx, downy, padx0, pady0, flip; float gain;
    Epilogue epi;
    // band kernel only
    int band_rows, bands_per_plane, tile_rows, pitch, rows_per_group;
    int flat_stage, async_stage; uint32_t inw_magic;      // narrow planes: stage the band's (contiguous) input span as one flat stream; ceil(2^32 / inW)
};

// floor division / modulo for possibly negative numerators
__device__ __forceinline__ int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }
__device__ __forceinline__ int posmod(int a, int b)   { int m = a % b; return m < 0 ? m + b : m; }

// ------------------------------------------------------------------------------------ generic
// UP == 1: every tap lands on a sample (no divisions in the tap loops); UP == 2: divisions become shifts.  Index arithmetic is 32-bit: the entry point
// rejects tensors with more than INT32_MAX elements (as the reference does, upfirdn2d.cpp:22-23,36).
template <class T, bool EPI, int UP>      // UP: 1, 2 = compile-time up factor (both directions), 0 = run-time
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(UpfirdnParams p) {
    typedef typename Acc<T>::type S;
    const T* __restrict__ x = (const T*)p.x;
    T* __restrict__ y = (T*)p.y;
    const unsigned total = (unsigned)((int64_t)p.N * p.C * p.outH * p.outW);
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int ox = (int)(idx % (unsigned)p.outW);
        unsigned r = idx / (unsigned)p.outW;
        const int oy = (int)(r % (unsigned)p.outH); r /= (unsigned)p.outH;
        const int c = (int)(r % (unsigned)p.C);
        const int n = (int)(r / (unsigned)p.C);
        // position of tap (0,0) in the zero-inserted (un-padded) image
        const int ux0 = ox * p.downx - p.padx0;
        const int uy0 = oy * p.downy - p.pady0;
        const T* xp = x + n * p.xs[0] + c * p.xs[1];
        S acc = (S)0;
        if (UP == 1) {
            for (int i = 0; i < p.fh; i++) {
                const int iy = uy0 + i;
                if (iy < 0 || iy >= p.inH) continue;
                const float* frow = p.f + (p.flip ? i : p.fh - 1 - i) * p.fsh;
                const T* xrow = xp + iy * p.xs[2];
                for (int j = 0; j < p.fw; j++) {
                    const int ix = ux0 + j;
                    if (ix < 0 || ix >= p.inW) continue;
                    acc += (S)__ldg(frow + (p.flip ? j : p.fw - 1 - j) * p.fsw) * to_acc<T>(__ldg(xrow + ix * p.xs[3]));
                }
            }
        } else {
            // first tap that lands on a real sample: (u0 + j) % up == 0
            const int upx = UP ? UP : p.upx, upy = UP ? UP : p.upy;
            const int j0 = posmod(-ux0, upx);
            const int i0 = posmod(-uy0, upy);
            for (int i = i0; i < p.fh; i += upy) {
                const int iy = (uy0 + i) / upy;              // exact: numerator is a multiple of upy
                if (iy < 0 || iy >= p.inH) continue;
                const int fi = p.flip ? i : p.fh - 1 - i;
                for (int j = j0; j < p.fw; j += upx) {
                    const int ix = (ux0 + j) / upx;
                    if (ix < 0 || ix >= p.inW) continue;
                    const int fj = p.flip ? j : p.fw - 1 - j;
                    acc += (S)__ldg(p.f + fi * p.fsh + fj * p.fsw) * to_acc<T>(__ldg(xp + iy * p.xs[2] + ix * p.xs[3]));
                }
            }
        }
        acc *= (S)p.gain;
        if (EPI) acc = apply_epilogue<S>(p.epi, acc, p.epi.b ? to_acc<T>(__ldg((const T*)p.epi.b + c)) : (S)0);
        y[n * p.ys[0] + c * p.ys[1] + oy * p.ys[2] + ox * p.ys[3]] = from_acc<T, S>(acc);
    }
}

// epilogue + store of 4 adjacent outputs of one row
template <class T, bool EPI>
__device__ __forceinline__ void store_quad(const UpfirdnParams& p, T* dst, float (&out)[4], int x0, float bias, bool vec_store) {
    if (EPI) {
#pragma unroll
        for (int c = 0; c < 4; c++) out[c] = apply_epilogue<float>(p.epi, out[c], bias);
    }
    if (vec_store) {
        *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++) if (x0 + c < p.outW) dst[c] = from_acc<T, float>(out[c]);
    }
}

// ------------------------------------------------------------------------------------ band (4x4, up=1)
// grid.x = plane * bands_per_plane + band; 256 threads; dynamic smem = tile_rows * pitch floats.
template <class T, int D, bool EPI>
__global__ void __launch_bounds__(256) upfirdn2d_band_kernel(UpfirdnParams p) {
    constexpr int F = 4;
    extern __shared__ float tile[];
    const T* __restrict__ x = (const T*)p.x;
    T* __restrict__ y = (T*)p.y;
    const int plane = blockIdx.x / p.bands_per_plane;
    const int band  = blockIdx.x - plane * p.bands_per_plane;
    const int oy0   = band * p.band_rows;
    const int rows  = min(p.band_rows, p.outH - oy0);
    const int iy0   = oy0 * D - p.pady0;                       // input row of tile row 0
    const int need  = (rows - 1) * D + F;                      // tile rows actually used
    const T* xp = x + (size_t)plane * p.inH * p.inW;

    // taps, flipped for true convolution, gain folded in
    float k[F][F];
#pragma unroll
    for (int i = 0; i < F; i++)
#pragma unroll
        for (int j = 0; j < F; j++)
            k[i][j] = __ldg(p.f + (p.flip ? i : F - 1 - i) * p.fsh + (p.flip ? j : F - 1 - j) * p.fsw) * p.gain;

    // stage: tile[r][c] = x[iy0 + r][c - padx0], zero outside the image.  One warp per row, 9 independent coalesced loads in flight per
    // thread (288 columns) before any shared-memory store: with 4 resident CTAs that is ~36 KB of reads in flight per SM.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KC = 9;
    if (p.flat_stage) {
        // Narrow planes (row shorter than the 288 columns one warp keeps in flight): the band is full width, so the input rows it needs are ONE
        // contiguous span of the plane.  Zero the tile (padding columns / rows outside the image), then stream the span with every lane busy
        // and 9 independent loads in flight per thread, scattering element i to tile[row][padx0 + col] with a magic-number division.
        for (int i = threadIdx.x; i < need * p.pitch; i += 256) tile[i] = 0.f;
        __syncthreads();
        const int r_lo = iy0 < 0 ? -iy0 : 0;
        const int r_hi = min(need, p.inH - iy0);
        const int total = (r_hi - r_lo) * p.inW;
        const T* src = xp + (ptrdiff_t)(iy0 + r_lo) * p.inW;
        for (int base = 0; base < total; base += 256 * KC) {
            float v[KC];
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const int i = base + (int)threadIdx.x + 256 * kk;
                v[kk] = i < total ? (float)to_acc<T>(__ldg(src + i)) : 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const int i = base + (int)threadIdx.x + 256 * kk;
                const int r = (int)__umulhi((uint32_t)i, p.inw_magic), c = i - r * p.inW + p.padx0;
                if (i < total && c < p.pitch) tile[(r_lo + r) * p.pitch + c] = v[kk];
            }
        }
    } else if (sizeof(T) == 4 && p.async_stage) {
        // fp32 rows by 4-byte cp.async (LDGSTS): global -> shared without a register round trip.  One warp per row; the interior of a row
        // (columns padx0 .. padx0 + inW) is copied in unrolled groups of 8 x 32 elements with immediate offsets, so a staged element costs about
        // one instruction; padding columns and rows outside the image are zeroed with plain stores.
        const int c_hi = p.padx0 + p.inW < p.pitch ? p.padx0 + p.inW : p.pitch;      // first padding column on the right
        for (int r = warp; r < need; r += 8) {
            const int iy = iy0 + r;
            float* dst = tile + r * p.pitch;
            if (iy < 0 || iy >= p.inH) {
                for (int c = lane * 4; c < p.pitch; c += 128) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                continue;
            }
            if (lane < p.padx0) dst[lane] = 0.f;
            for (int c = c_hi + lane; c < p.pitch; c += 32) dst[c] = 0.f;
            const float* src = reinterpret_cast<const float*>(xp) + (ptrdiff_t)iy * p.inW + lane;
            uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + p.padx0 + lane);
            const int n = c_hi - p.padx0;                                            // elements to copy
            int c = 0;
            for (; c + 256 <= n; c += 256, src += 256, d += 1024) {
#pragma unroll
                for (int kk = 0; kk < 8; kk++) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 128u * kk), "l"(src + 32 * kk) : "memory");
            }
            for (; c + 128 <= n; c += 128, src += 128, d += 512) {
#pragma unroll
                for (int kk = 0; kk < 4; kk++) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 128u * kk), "l"(src + 32 * kk) : "memory");
            }
            for (; c + lane < n; c += 32, src += 32, d += 128) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else
    for (int r = warp; r < need; r += 8) {
        const int iy = iy0 + r;
        const bool row_ok = (iy >= 0) && (iy < p.inH);
        const T* src = xp + (ptrdiff_t)iy * p.inW - p.padx0;
        float* dst = tile + r * p.pitch;
        for (int cb = 0; cb < p.pitch; cb += 32 * KC) {
            float v[KC];
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const int c = cb + lane + 32 * kk;
                const int ix = c - p.padx0;
                v[kk] = (row_ok && c < p.pitch && ix >= 0 && ix < p.inW) ? (float)to_acc<T>(__ldg(src + c)) : 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const int c = cb + lane + 32 * kk;
                if (c < p.pitch) dst[c] = v[kk];
            }
        }
    }
    __syncthreads();

    // compute: work item = (row group, quad of 4 adjacent output columns).  Rows are read from the tile as aligned float4 (the pitch is a
    // multiple of 4 and output quad x0 starts at tile column x0 * D), horizontally filtered once per input row, and the four most recent
    // filtered rows are kept in a register ring for the vertical pass: 8 FMAs and < 1 shared-memory load per output when the filter is an
    // outer product (it always is for PASTA-GAN's [1,3,3,1]); a general 4x4 filter takes the 16-FMA path over the raw rows.
    constexpr int NV4 = (3 * D + 4 + 3) / 4;                   // float4 loads per tile row per item: 2 (D = 1) or 3 (D = 2)
    const int quads = (p.outW + 3) >> 2;
    const int rpg = p.rows_per_group;
    const int groups = (rows + rpg - 1) / rpg;
    const int items = groups * quads;
    const int c_ch = plane % p.C;
    float bias = 0.f;
    if (EPI && p.epi.b) bias = (float)to_acc<T>(__ldg((const T*)p.epi.b + c_ch));
    const BandEpilogue bepi = make_band_epilogue(p.epi);
    T* yp = y + (size_t)plane * p.outH * p.outW;
    const bool vec_store = sizeof(T) == 4 && (p.outW & 3) == 0 && ((((size_t)p.outH * p.outW) & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);

    // rank-1 test: k[i][j] == ky[i] * kx[j] with kx = k[0][:], ky = k[:][0] / k[0][0]
    float kx[F], ky[F];
    bool separable = k[0][0] != 0.f;
    {
        float kmax = 0.f;
#pragma unroll
        for (int i = 0; i < F; i++)
#pragma unroll
            for (int j = 0; j < F; j++) kmax = fmaxf(kmax, fabsf(k[i][j]));
#pragma unroll
        for (int i = 0; i < F; i++) { kx[i] = k[0][i]; ky[i] = separable ? k[i][0] / k[0][0] : 0.f; }
#pragma unroll
        for (int i = 0; i < F; i++)
#pragma unroll
            for (int j = 0; j < F; j++) separable = separable && fabsf(k[i][j] - ky[i] * kx[j]) <= 1e-6f * kmax;
    }

    for (int item = threadIdx.x; item < items; item += 256) {
        const int g = item / quads;
        const int x0 = (item - g * quads) << 2;
        const int r0 = g * rpg;
        const int nr = min(rpg, rows - r0);
        const float* rowp = tile + (r0 * D) * p.pitch + x0 * D;
        float out[4];
        if (separable) {
            // ring of horizontally filtered rows as packed pairs; slot of input row j (relative to the group's first row) is j & 3, so four
            // unrolled steps return to the same slot assignment and no register ever moves
            float2 H[4][2];
            auto hfilter = [&](const float* rp, float2 (&hr)[2]) {
                float v[NV4 * 4];
#pragma unroll
                for (int q = 0; q < NV4; q++) {
                    const float4 t = reinterpret_cast<const float4*>(rp)[q];
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                float a[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    a[c] = kx[0] * v[c * D];
#pragma unroll
                    for (int j = 1; j < F; j++) a[c] = fmaf(kx[j], v[c * D + j], a[c]);
                }
                hr[0] = make_float2(a[0], a[1]); hr[1] = make_float2(a[2], a[3]);
            };
#pragma unroll
            for (int i = 0; i < F - D; i++) hfilter(rowp + i * p.pitch, H[i]);
            rowp += (F - D) * p.pitch;
            T* yrow = yp + (size_t)(oy0 + r0) * p.outW + x0;
            const float2 ky2[F] = {make_float2(ky[0], ky[0]), make_float2(ky[1], ky[1]), make_float2(ky[2], ky[2]), make_float2(ky[3], ky[3])};
            const float2 bias2 = make_float2(bias, bias);
#define PG_BAND_STEP(K)                                                                                                              \
            if (r + (K) < nr) {                                                                                                      \
                _Pragma("unroll") for (int i = F - D; i < F; i++) hfilter(rowp + (i - (F - D)) * p.pitch, H[(D * (K) + i) & 3]);     \
                rowp += D * p.pitch;                                                                                                 \
                float2 o[2];                                                                                                         \
                _Pragma("unroll") for (int h2 = 0; h2 < 2; h2++) {                                                                   \
                    float2 acc = EPI ? f2fma(ky2[0], H[(D * (K)) & 3][h2], bias2) : f2mul(ky2[0], H[(D * (K)) & 3][h2]);             \
                    _Pragma("unroll") for (int i = 1; i < F; i++) acc = f2fma(ky2[i], H[(D * (K) + i) & 3][h2], acc);                \
                    o[h2] = EPI ? band_act(bepi, acc) : acc;                                                                         \
                }                                                                                                                    \
                if (vec_store) *reinterpret_cast<float4*>(yrow) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);                       \
                else {                                                                                                               \
                    const float ov[4] = {o[0].x, o[0].y, o[1].x, o[1].y};                                                            \
                    _Pragma("unroll") for (int c = 0; c < 4; c++) if (x0 + c < p.outW) yrow[c] = from_acc<T, float>(ov[c]);          \
                }                                                                                                                    \
                yrow += p.outW;                                                                                                      \
            }
            for (int r = 0; r < nr; r += 4) { PG_BAND_STEP(0) PG_BAND_STEP(1) PG_BAND_STEP(2) PG_BAND_STEP(3) }
#undef PG_BAND_STEP
        } else {
            float w[F][NV4 * 4];                               // raw rows
            auto load = [&](const float* rp, float (&wr)[NV4 * 4]) {
#pragma unroll
                for (int q = 0; q < NV4; q++) {
                    const float4 t = reinterpret_cast<const float4*>(rp)[q];
                    wr[4 * q] = t.x; wr[4 * q + 1] = t.y; wr[4 * q + 2] = t.z; wr[4 * q + 3] = t.w;
                }
            };
#pragma unroll
            for (int i = 0; i < F - D; i++) load(rowp + i * p.pitch, w[i + D]);
            rowp += (F - D) * p.pitch;
            for (int r = 0; r < nr; r++) {
#pragma unroll
                for (int i = 0; i < F - D; i++)
#pragma unroll
                    for (int c = 0; c < NV4 * 4; c++) w[i][c] = w[i + D][c];
#pragma unroll
                for (int i = F - D; i < F; i++) load(rowp + (i - (F - D)) * p.pitch, w[i]);
                rowp += D * p.pitch;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float a = 0.f;
#pragma unroll
                    for (int i = 0; i < F; i++)
#pragma unroll
                        for (int j = 0; j < F; j++) a = fmaf(k[i][j], w[i][c * D + j], a);
                    out[c] = a;
                }
                store_quad<T, EPI>(p, yp + (size_t)(oy0 + r0 + r) * p.outW + x0, out, x0, bias, vec_store);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ band (4x4, up=2): polyphase
// Zero-insert up-2 + 4x4 FIR (reference tile kernel upfirdn2d.cu:252 <2,2,1,1,4,4,64,16,1>): the RGB skip image of every synthesis block
// (networks.py:5711 -> upfirdn2d.py:308-343, pad (2,1,2,1), gain 4) and the backward of every down-2 (upfirdn2d.py:251-261).  Only the taps that land
// on a sample are evaluated: output (oy, ox) with b = ox - padx0 reads input columns ceil(b/2), ceil(b/2)+1 with taps (b&1), (b&1)+2 -- 4 FMAs per
// output instead of 16 tests.  One CTA owns a full-width band of output rows of one plane; the input rows it needs (a quarter of the output's bytes)
// are staged in shared memory with a zero border, each thread walks a quad of 4 adjacent output columns down the band holding the two live input
// rows (4 columns each) in registers, and stores aligned float4.  grid.x = plane * bands_per_plane + band; smem = tile_rows * pitch floats.
template <int PR, int PODD>
__device__ __forceinline__ void up2_quad(const float (&k)[4][4], const float (&a)[4], const float (&b)[4], float (&out)[4]) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int pc = (c + PODD) & 1;                          // horizontal tap parity of this output column
        const int q = (c + PODD + pc) >> 1;                     // register column of its first tap
        out[c] = fmaf(k[PR][pc], a[q], fmaf(k[PR][pc + 2], a[q + 1], fmaf(k[PR + 2][pc], b[q], k[PR + 2][pc + 2] * b[q + 1])));
    }
}

template <class T, bool EPI>
__global__ void __launch_bounds__(256) upfirdn2d_up2_band_kernel(UpfirdnParams p) {
    constexpr int F = 4;
    extern __shared__ float tile[];
    const T* __restrict__ x = (const T*)p.x;
    T* __restrict__ y = (T*)p.y;
    const int plane = blockIdx.x / p.bands_per_plane;
    const int band  = blockIdx.x - plane * p.bands_per_plane;
    const int oy0   = band * p.band_rows;
    const int rows  = min(p.band_rows, p.outH - oy0);
    const int padl  = (p.padx0 + 1) >> 1;                       // tile column of input column 0
    // input rows of the band: first = ceil((oy0 - pady0) / 2), last = ceil((oy0 + rows - 1 - pady0) / 2) + 1
    const int by0 = oy0 - p.pady0;
    const int iy0 = (by0 + (by0 & 1)) >> 1;                     // arithmetic shift: exact, numerator even
    const int byl = oy0 + rows - 1 - p.pady0;
    const int need = ((byl + (byl & 1)) >> 1) + 2 - iy0;        // staged rows
    const T* xp = x + (size_t)plane * p.inH * p.inW;

    float k[F][F];
#pragma unroll
    for (int i = 0; i < F; i++)
#pragma unroll
        for (int j = 0; j < F; j++)
            k[i][j] = __ldg(p.f + (p.flip ? i : F - 1 - i) * p.fsh + (p.flip ? j : F - 1 - j) * p.fsw) * p.gain;

    // stage: zero the tile, then stream the (contiguous) span of input rows with 8 independent loads in flight per thread
    for (int i = threadIdx.x; i < need * p.pitch; i += 256) tile[i] = 0.f;
    __syncthreads();
    {
        const int r_lo = iy0 < 0 ? -iy0 : 0;
        const int r_hi = min(need, p.inH - iy0);
        const int total = (r_hi - r_lo) * p.inW;
        const T* src = xp + (ptrdiff_t)(iy0 + r_lo) * p.inW;
        constexpr int KC = 8;
        for (int base = 0; base < total; base += 256 * KC) {
            float v[KC];
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const int i = base + (int)threadIdx.x + 256 * kk;
                v[kk] = i < total ? (float)to_acc<T>(__ldg(src + i)) : 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < KC; kk++) {
                const int i = base + (int)threadIdx.x + 256 * kk;
                const int r = (int)__umulhi((uint32_t)i, p.inw_magic), c = i - r * p.inW + padl;
                if (i < total) tile[(r_lo + r) * p.pitch + c] = v[kk];
            }
        }
    }
    __syncthreads();

    const int quads = (p.outW + 3) >> 2;
    const int rpg = p.rows_per_group;
    const int groups = (rows + rpg - 1) / rpg;
    const int items = groups * quads;
    float bias = 0.f;
    if (EPI && p.epi.b) bias = (float)to_acc<T>(__ldg((const T*)p.epi.b + plane % p.C));
    T* yp = y + (size_t)plane * p.outH * p.outW;
    const bool vec_store = sizeof(T) == 4 && (p.outW & 3) == 0 && ((((size_t)p.outH * p.outW) & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    const int podd = p.padx0 & 1;                               // parity of (x0 - padx0) for every quad (x0 is a multiple of 4)
    // horizontal tap pairs and register columns of the quad's 4 outputs (see header comment): even base: (0,1) (1,2) (1,2) (2,3); odd: (1,2) (1,2) (2,3) (2,3)
    for (int item = threadIdx.x; item < items; item += 256) {
        const int g = item / quads;
        const int x0 = (item - g * quads) << 2;
        const int r0 = g * rpg;
        const int nr = min(rpg, rows - r0);
        // tile column of register column 0: floor((x0 - padx0) / 2) + padl == x0 / 2 for either parity of padx0
        const float* colp = tile + (x0 >> 1);
        float a[4], b[4];                                       // input rows ra, ra + 1
        int ra = INT_MIN;
        for (int r = 0; r < nr; r++) {
            const int oy = oy0 + r0 + r;
            const int by = oy - p.pady0;
            const int pr = by & 1;
            const int rn = ((by + pr) >> 1) - iy0;              // staged row of the first vertical tap
            if (rn != ra) {
                if (rn == ra + 1) {
#pragma unroll
                    for (int c = 0; c < 4; c++) a[c] = b[c];
                } else {
                    const float2 t0 = *reinterpret_cast<const float2*>(colp + rn * p.pitch), t1 = *reinterpret_cast<const float2*>(colp + rn * p.pitch + 2);
                    a[0] = t0.x; a[1] = t0.y; a[2] = t1.x; a[3] = t1.y;
                }
                const float2 u0 = *reinterpret_cast<const float2*>(colp + (rn + 1) * p.pitch), u1 = *reinterpret_cast<const float2*>(colp + (rn + 1) * p.pitch + 2);
                b[0] = u0.x; b[1] = u0.y; b[2] = u1.x; b[3] = u1.y;
                ra = rn;
            }
            float out[4];
            // pr and podd are uniform over the CTA (even rows_per_group), so the four variants are branches with compile-time tap indices
            if (pr) { if (podd) up2_quad<1, 1>(k, a, b, out); else up2_quad<1, 0>(k, a, b, out); }
            else    { if (podd) up2_quad<0, 1>(k, a, b, out); else up2_quad<0, 0>(k, a, b, out); }
            store_quad<T, EPI>(p, yp + (size_t)oy * p.outW + x0, out, x0, bias, vec_store);
        }
    }
}

// ------------------------------------------------------------------------------------ host side
static bool contiguous_nchw(const int32_t sz[4], const int64_t st[4]) {
    return st[3] == 1 && st[2] == sz[3] && st[1] == (int64_t)sz[2] * sz[3] && (st[0] == (int64_t)sz[1] * sz[2] * sz[3] || sz[0] == 1);
}

template <class T, bool EPI>
static int launch_upfirdn2d(UpfirdnParams p, bool band_ok, bool up2_ok, cudaStream_t stream) {
    if (up2_ok) {
        int band_rows = 16 * 256 / ((p.outW + 3) / 4);                  // 256 work items of 16 rows each (narrow planes: taller bands)
        band_rows = band_rows < 32 ? 32 : (band_rows > 256 ? 256 : band_rows);
        if (band_rows > p.outH) band_rows = p.outH;
        p.bands_per_plane = (p.outH + band_rows - 1) / band_rows;
        // few planes (the 3-channel RGB skip): shorter bands so that the grid still covers the SMs
        while (band_rows > 8 && (int64_t)p.N * p.C * p.bands_per_plane < 2 * kNumSMs) { band_rows /= 2; p.bands_per_plane = (p.outH + band_rows - 1) / band_rows; }
        band_rows = (p.outH + p.bands_per_plane - 1) / p.bands_per_plane;
        p.band_rows = band_rows;
        p.tile_rows = band_rows / 2 + 3;
        p.pitch = ((p.outW + 3) / 4) * 2 + 4;                    // register column 3 of the last quad: (outW/4 - 1) * 2 + 3, even pitch for float2 reads
        if (p.pitch < p.inW + ((p.padx0 + 1) >> 1) + 1) p.pitch = (p.inW + ((p.padx0 + 1) >> 1) + 2) & ~1;
        const int quads = (p.outW + 3) / 4;
        int rpg = band_rows * quads / 256;
        p.rows_per_group = (rpg < 4 ? 4 : (rpg > 16 ? 16 : rpg)) & ~1;      // even: the row parity of step r is then uniform over the CTA
        p.inw_magic = (uint32_t)((0x100000000ull + (uint64_t)p.inW - 1) / (uint64_t)p.inW);
        const size_t smem = (size_t)p.tile_rows * p.pitch * sizeof(float);
        const int64_t blocks = (int64_t)p.N * p.C * p.bands_per_plane;
        if (smem <= 96 * 1024 && blocks <= INT32_MAX) {
            auto kern = upfirdn2d_up2_band_kernel<T, EPI>;
            if (smem > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)blocks, 256, smem, stream>>>(p);
            return launch_status("upfirdn2d(up2 band)");
        }
    }
    if (band_ok) {
        const int D = p.downx;
        // band height: as tall as a ~74 KB tile allows (three CTAs per SM), at most 32 rows.  Taller bands re-read fewer halo rows and, more
        // importantly, let a work item walk 16 rows: the three rows that prime its register window are then 19 % extra horizontal filtering and
        // shared-memory reads instead of 38 % at 8 rows (the kernel is bound by issue slots and shared-memory wavefronts, not by HBM)
        const int quads0 = (p.outW + 3) / 4;
        int band_rows = 16 * 256 / quads0;                               // 256 work items of 16 rows each
        band_rows = band_rows < 32 ? 32 : (band_rows > 256 ? 256 : band_rows);
        while (band_rows > 8 && (size_t)((band_rows - 1) * D + 4) * p.pitch * sizeof(float) > 74 * 1024) band_rows /= 2;
        if (band_rows > p.outH) band_rows = p.outH;
        // balanced bands (a 33-row plane is 17 + 16 rows, not 32 + 1); a plane a few rows taller than one band stays one band
        if (p.outH <= band_rows + 8 && (size_t)((p.outH - 1) * D + 4) * p.pitch * sizeof(float) <= 48 * 1024) band_rows = p.outH;
        p.bands_per_plane = (p.outH + band_rows - 1) / band_rows;
        band_rows = (p.outH + p.bands_per_plane - 1) / p.bands_per_plane;
        p.band_rows = band_rows;
        p.tile_rows = (band_rows - 1) * D + 4;
        // rows per work item: aim for >= 256 items per CTA (one per thread) but at least 2 rows to amortise the window priming
        const int quads = (p.outW + 3) / 4;
        int rpg = band_rows * quads / 256;
        p.rows_per_group = rpg < 2 ? 2 : (rpg > 16 ? 16 : rpg);
        p.async_stage = (sizeof(T) == 4 && p.inW >= 100) ? 1 : 0;       // fp32 rows of >= 100 columns by cp.async (measured: 6-25 % faster); narrower planes keep the flat stream       // fp32 rows by cp.async; narrow planes keep the flat stream (every lane busy)
        p.flat_stage = (!p.async_stage && p.pitch <= 192) ? 1 : 0;
        p.inw_magic = (uint32_t)((0x100000000ull + (uint64_t)p.inW - 1) / (uint64_t)p.inW);
        const size_t smem = (size_t)p.tile_rows * p.pitch * sizeof(float);
        const int64_t blocks = (int64_t)p.N * p.C * p.bands_per_plane;
        if (smem <= 96 * 1024 && blocks <= INT32_MAX) {
            auto kern = (D == 1) ? upfirdn2d_band_kernel<T, 1, EPI> : upfirdn2d_band_kernel<T, 2, EPI>;
            if (smem > 48 * 1024) PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)blocks, 256, smem, stream>>>(p);
            return launch_status("upfirdn2d(band)");
        }
    }
    const int64_t total = (int64_t)p.N * p.C * p.outH * p.outW;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 64;
    if (blocks > cap) blocks = cap;
    if (p.upx == 1 && p.upy == 1)      upfirdn2d_generic_kernel<T, EPI, 1><<<(unsigned)blocks, 256, 0, stream>>>(p);
    else if (p.upx == 2 && p.upy == 2) upfirdn2d_generic_kernel<T, EPI, 2><<<(unsigned)blocks, 256, 0, stream>>>(p);
    else                               upfirdn2d_generic_kernel<T, EPI, 0><<<(unsigned)blocks, 256, 0, stream>>>(p);
    return launch_status("upfirdn2d(generic)");
}

static int upfirdn2d_entry(const void* x, const float* f, void* y,
                           const int32_t in_size[4], const int64_t in_stride[4],
                           const int32_t out_size[4], const int64_t out_stride[4],
                           int32_t fh, int32_t fw, int64_t fsh, int64_t fsw,
                           int32_t upx, int32_t upy, int32_t downx, int32_t downy,
                           int32_t padx0, int32_t padx1, int32_t pady0, int32_t pady1,
                           int32_t flip, float gain, const Epilogue& epi, int32_t dtype, void* stream) {
    PG_REQUIRE(in_size && in_stride && out_size && out_stride, "size/stride arrays must not be NULL");
    PG_REQUIRE(fh >= 1 && fw >= 1, "f must be at least 1x1");
    PG_REQUIRE(upx >= 1 && upy >= 1, "upsampling factor must be at least 1");
    PG_REQUIRE(downx >= 1 && downy >= 1, "downsampling factor must be at least 1");
    for (int i = 0; i < 4; i++) PG_REQUIRE(in_size[i] >= 0 && out_size[i] >= 0, "negative size");
    const int64_t in_numel = (int64_t)in_size[0] * in_size[1] * in_size[2] * in_size[3];
    PG_REQUIRE(in_numel <= INT32_MAX, "x is too large");
    PG_REQUIRE((int64_t)fh * fw <= INT32_MAX, "f is too large");
    const int outW = (in_size[3] * upx + padx0 + padx1 - fw + downx) / downx;
    const int outH = (in_size[2] * upy + pady0 + pady1 - fh + downy) / downy;
    PG_REQUIRE(outW >= 1 && outH >= 1, "output must be at least 1x1");
    PG_REQUIRE(out_size[0] == in_size[0] && out_size[1] == in_size[1] && out_size[2] == outH && out_size[3] == outW,
               "out_size must be [%d, %d, %d, %d]", in_size[0], in_size[1], outH, outW);
    const int64_t out_numel = (int64_t)out_size[0] * out_size[1] * outH * outW;
    PG_REQUIRE(out_numel <= INT32_MAX, "output is too large");
    if (out_numel == 0) return PG_OK;
    PG_REQUIRE(x && f && y, "x, f and y must be device pointers");
    if (epi.enabled)
        PG_REQUIRE(epi.act == PG_ACT_LINEAR || epi.act == PG_ACT_RELU || epi.act == PG_ACT_LRELU,
                   "fused epilogue supports linear / relu / lrelu only (act=%d)", epi.act);

    UpfirdnParams p;
    p.x = x; p.f = f; p.y = y;
    p.N = in_size[0]; p.C = in_size[1]; p.inH = in_size[2]; p.inW = in_size[3]; p.outH = outH; p.outW = outW;
    for (int i = 0; i < 4; i++) { p.xs[i] = in_stride[i]; p.ys[i] = out_stride[i]; }
    p.fh = fh; p.fw = fw; p.fsh = fsh; p.fsw = fsw;
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0; p.flip = flip ? 1 : 0; p.gain = gain;
    p.epi = epi;
    p.band_rows = p.bands_per_plane = p.tile_rows = p.rows_per_group = 0; p.flat_stage = p.async_stage = 0; p.inw_magic = 0;
    // columns the band tile must hold: taps of the last output column reach (outW-1)*D + 3
    // tile columns: the last quad of outputs starts at 4*(ceil(outW/4)-1)*D and reads 4*NV4 floats; multiple of 4 for aligned float4 reads
    p.pitch = 4 * ((outW + 3) / 4 - 1) * downx + (downx == 1 ? 8 : 12);

    const bool band_ok = dtype != PG_F64 && fh == 4 && fw == 4 && upx == 1 && upy == 1 && downx == downy && (downx == 1 || downx == 2) &&
                         padx0 >= 0 && pady0 >= 0 && outW >= 8 && (int64_t)outH * outW >= 256 &&      // smaller planes: per-CTA setup costs more than the generic kernel
                        
                         contiguous_nchw(in_size, in_stride) && contiguous_nchw(out_size, out_stride);
    // polyphase up-2 band kernel: 4x4 filter, up 2 in both directions, no decimation, non-negative left / top padding, every tap row / column the
    // band touches inside the staged tile (right / bottom padding may be anything: samples past the image are the tile's zero border)
    const bool up2_ok = dtype != PG_F64 && fh == 4 && fw == 4 && upx == 2 && upy == 2 && downx == 1 && downy == 1 && padx0 >= 0 && pady0 >= 0 &&
                        outW >= 8 && (int64_t)outH * outW >= 256 && contiguous_nchw(in_size, in_stride) && contiguous_nchw(out_size, out_stride);
    cudaStream_t s = (cudaStream_t)stream;
    const bool e = epi.enabled != 0;
    switch (dtype) {
        case PG_F32: return e ? launch_upfirdn2d<float, true>(p, band_ok, up2_ok, s)  : launch_upfirdn2d<float, false>(p, band_ok, up2_ok, s);
        case PG_F16: return e ? launch_upfirdn2d<__half, true>(p, band_ok, up2_ok, s) : launch_upfirdn2d<__half, false>(p, band_ok, up2_ok, s);
        case PG_F64: return e ? launch_upfirdn2d<double, true>(p, false, false, s)    : launch_upfirdn2d<double, false>(p, false, false, s);
    }
    return fail(PG_ERR_INVALID_ARGUMENT, "unsupported dtype %d", dtype);
}

}  // namespace pg

extern "C" int pg_upfirdn2d(const void* x, const float* f, void* y,
                            const int32_t in_size[4], const int64_t in_stride[4],
                            const int32_t out_size[4], const int64_t out_stride[4],
                            int32_t fh, int32_t fw, int64_t f_stride_h, int64_t f_stride_w,
                            int32_t upx, int32_t upy, int32_t downx, int32_t downy,
                            int32_t padx0, int32_t padx1, int32_t pady0, int32_t pady1,
                            int32_t flip, float gain, int32_t dtype, void* stream) {
    pg::Epilogue e; e.b = nullptr; e.act = PG_ACT_LINEAR; e.alpha = 0.f; e.act_gain = 1.f; e.clamp = -1.f; e.enabled = 0;
    return pg::upfirdn2d_entry(x, f, y, in_size, in_stride, out_size, out_stride, fh, fw, f_stride_h, f_stride_w,
                               upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain, e, dtype, stream);
}

extern "C" int pg_upfirdn2d_bias_act(const void* x, const float* f, const void* b, void* y,
                                     const int32_t in_size[4], const int64_t in_stride[4],
                                     const int32_t out_size[4], const int64_t out_stride[4],
                                     int32_t fh, int32_t fw, int64_t f_stride_h, int64_t f_stride_w,
                                     int32_t upx, int32_t upy, int32_t downx, int32_t downy,
                                     int32_t padx0, int32_t padx1, int32_t pady0, int32_t pady1,
                                     int32_t flip, float gain,
                                     int32_t act, float alpha, float act_gain, float clamp,
                                     int32_t dtype, void* stream) {
    pg::Epilogue e; e.b = b; e.act = act; e.alpha = alpha; e.act_gain = act_gain; e.clamp = clamp; e.enabled = 1;
    return pg::upfirdn2d_entry(x, f, y, in_size, in_stride, out_size, out_stride, fh, fw, f_stride_h, f_stride_w,
                               upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain, e, dtype, stream);
}
