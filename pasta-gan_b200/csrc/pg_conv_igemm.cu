// Implicit-GEMM convolution for sm_100a: tcgen05.mma with TMEM accumulators, fp32 NCHW in / out.
//
// Replaces the cuDNN calls the reference makes for every convolution (torch_utils/ops/conv2d_gradfix.py:35-43 via
// conv2d_resample.py:29-54) and the separate modulation / demodulation / noise / bias_act passes around them
// (training/networks.py:37-94 modulated_conv2d, :296-315 SynthesisLayer.forward, :170-179 Conv2dLayer.forward,
// :4342-4354 Spade_Conv2dLayer.forward).  One kernel computes
//
//     y[n,o,h,w] = clamp( act( d[n,o] * sum_{c,kh,kw} W[o,c,kh,kw] * (s[n,c] * in_act(x[n,c,h+kh-p,w+kw-p])) + noise[h,w] + b[o] ) * gain )
//
// for stride-1 "same" 3x3 and 1x1 convolutions (per-sample style s and demodulation d optional: s == 1, d == 1 is the
// plain Conv2dLayer).  GEMM view: M = pixels of one sample, N = Cout, K = taps x Cin.
//
// Data layout / algorithm
//   * The image of one sample is addressed as a flattened strip with pitch PW = W + 1: the extra column is the zero
//     padding shared by the right edge of row h and the left edge of row h+1.  An output tile is BM = 128 * NACC
//     consecutive strip positions; filter tap (kh, kw) of a 3x3 kernel reads strip position m + (kh-1)*PW + (kw-1).
//   * A operand (activations): converter warps read fp32 NCHW from global/L2 (coalesced along w), apply the input
//     activation and the per-sample style, convert to fp16/bf16 and store [plane = 8 channels][position][8 x 2 B] in
//     shared memory.  This is the K-major SWIZZLE_NONE canonical layout with SBO = 128 B, i.e. GEMM rows are exactly
//     16 B apart, so the nine taps are nine *descriptor start addresses* into the same staged tile: the im2col
//     matrix is never materialised and every input element is staged once per 16-channel chunk.
//   * B operand (weights): a prepack kernel writes fp16/bf16 tiles in the same canonical layout to a workspace
//     ([n-tile][chunk][tap][plane][n][8]); one cp.async.bulk per chunk brings all taps of a 16-channel chunk in.
//   * tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN (<= 256), K = 16, issued by one thread; NACC accumulators
//     of BN fp32 columns live in TMEM (NACC * BN <= 512 columns).
//   * Epilogue: tcgen05.ld 32x32b.x16 -> demodulate, add noise and bias, activation, gain, clamp -> coalesced fp32
//     NCHW stores.  The convolution result never round-trips through HBM before bias_act.
//   * mbarrier pipelines: A full/empty (converters <-> MMA), B full/empty (bulk copy <-> MMA), accumulator full.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>
#include "pg_common.cuh"

namespace pg {

constexpr int kConvWarps   = 8;                 // A converters, later the epilogue
constexpr int kConvThreads = 64 + 32 * kConvWarps;
constexpr int kKC          = 16;                // channels per pipeline chunk == one UMMA K step

struct ConvParams {
    const float* x; const float* x2; const float* residual; int cin1;   // channels [0, cin1) come from x, [cin1, Cin) from x2 (fused concat)
    const void* wpack; const float* styles; const float* dcoefs; const float* noise; const float* bias; float* y;
    long long noise_bstride;
    int N, Cin, Cout, H, W, ks;
    int ntiles_n;              // N tiles (persistent kernel: tile list = n-tile major)
    int band_tw, nbands, Wimg; // column bands (wide images): band_tw output columns per band (0 = off), bands per image, real image width.
                               // In band mode W = band_tw + 4 (two halo columns each side, real data) and PW = W: a tile then covers several rows
    int PW, Lp, tiles_per_img, NACC, BN, nchunks, ntaps, PA;      // PA: staged strip positions (multiple of 32)
    int SA, SB, tps; uint32_t a_stage_bytes, b_slot_bytes, b_tile_bytes;   // B ring: SB slots of `tps` taps each
    int in_act; float in_alpha, in_gain; int act; float alpha, gain, clamp; int fmt;
    uint32_t idesc; uint32_t tmem_cols;
    // up-2 polyphase mode: BN virtual channels = phases x couts (see launch code); 0 = plain
    int up2; int cout_real; int up2_pair;      // up2_pair: paired-phase epilogue (see epilogue_tile_up2_pair)
    int down2, cin_real, hin, win;   // down-2 mode: the A operand is the 4-plane space-to-depth view of x [N,cin_real,hin,win]
    const float* sp_x; const float* sp_mean; const float* sp_rstd; int spade;   // SPADE epilogue: y = act((x-mean)*rstd*(1+gamma)+beta)*gain
    long long* dbg;            // optional per-CTA phase timestamps (pg_debug_set_buffer), NULL in production
    int in_half, out_half;     // x / y are fp16 NCHW (intermediates between our own layers); everything else stays fp32
    int cgroups;               // converter warp groups that take alternate chunks (lean loader): several chunks' load round trips in flight
    int pipe, ldmode, dbgmode, lean;          // converter knobs (tuning): register double-buffering on/off, L1::no_allocate loads
    int im2col; uint32_t kk_magic, ks_magic;   // im2col mode: real kernel size (0 = off); ceil(2^32 / k^2), ceil(2^32 / k)
    long long wpack_sample_stride;             // bytes between the packed weight sets of consecutive samples (groups = N form, networks.py:84-94); 0 = shared weights
    // TMA A operand (x is fp16 in the channel-blocked layout [N][C/8][H][W][8], see include/pasta_b200.h PG_LAYOUT_C8): the staged strip is a
    // tensor-map box of tma_rows image rows x PW positions x 2 channel blocks, written by cp.async.bulk.tensor; no converter warps
    int tma_a, tma_rows, tma_cb, tma_cb2;      // on/off, rows per box, channel blocks per sample of x (Cin1 / 8) and of x2 (0: no second input)
    uint32_t a_tx_bytes;                       // bytes one A box delivers (the stage itself is rounded up to 128 B)
    uint32_t a_lbo16;                          // A descriptor leading-dimension byte offset >> 4 (distance between the two 8-channel planes of a stage)
    int y_c8, cb_out;                          // y is fp16 channel-blocked [N][cb_out][H][W][8]
    int res_c8;                                // the residual is fp16 channel-blocked (same shape as a channel-blocked y)
    int vec2; uint32_t w_magic;   // aligned 8-byte loader (see the converter section); ceil(2^32 / W)
    uint32_t pw_magic;         // ceil(2^32 / PW): q / PW == umulhi(q, pw_magic) for the strip positions that occur
    int tma_down2;             // TMA A operand of the down-2 form: x is channel-blocked [N][cin_real/8][hin][win][8]; chunk ci is parity (a, b) = ci / (cin_real/16) of two
                               // channel blocks, brought by a 4-D box with element stride 2 along rows and columns (space-to-depth done by the tensor map)
};

// Phase timestamps, load / store / fence suppression and the loader experiments exist only in -DPG_DEBUG builds (tools/conv_timeline.py builds its
// own library); in the release library PG_DBG / PG_DBGMODE / PG_LDMODE / PG_PIPE are compile-time constants and the code behind them is removed.
#ifdef PG_DEBUG
#define PG_DBG(p)     ((p).dbg != nullptr)
#define PG_DBGMODE(p) ((p).dbgmode)
#define PG_LDMODE(p)  ((p).ldmode)
#define PG_PIPE(p)    ((p).pipe)
#define PG_TS(slot) do { if (PG_DBG(p)) p.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); } while (0)
#define PG_PUT(slot, v) do { if (PG_DBG(p)) p.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = (v); } while (0)
#else
#define PG_DBG(p)     false
#define PG_DBGMODE(p) 0
#define PG_LDMODE(p)  0
#define PG_PIPE(p)    0
#define PG_TS(slot) do { } while (0)
#define PG_PUT(slot, v) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// 3-D tensor-map tile load (TMA): box -> shared memory, completion counted in bytes on `bar`; out-of-bounds elements arrive as zeros
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

// Shared-memory matrix descriptors (K-major, SWIZZLE_NONE, rows 16 B apart: SBO = 128 B, K chunks LBO bytes apart, descriptor version 1) are
// assembled inline in mma_issue_loop: constant high word, low word = (LBO >> 4) << 16 | (start address >> 4).

// The whole MMA warp runs the issue loop in uniform control flow and the tcgen05 instructions sit in `if (elected)` regions, `elected` coming from
// elect.sync: ptxas then knows a single lane is active and moves descriptors to uniform registers with plain R2URs.  Under `if (lane == 0)` it
// emits an ELECT / R2UR / BRA.U.ANY waterfall loop before every MMA that costs more than the MMA itself (149 vs 64 cycles, tools/micro).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// fp32 -> tf32 (10-bit mantissa, round to nearest, ties away): the tensor core would otherwise just drop the low 13 bits
__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp16 output element: the same round-to-nearest / saturate-to-finite conversion the consuming layer's loader would apply to the fp32 value
__device__ __forceinline__ void store_half(float* y, size_t off, float v) {
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    reinterpret_cast<unsigned short*>(y)[off] = h;
}

// eight fp32 -> one 16-byte row of the channel-blocked fp16 layout (the conversion the consuming layer's loader would apply)
__device__ __forceinline__ uint4 pack_half8(const float* v) {
    uint4 r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(v[1]), "f"(v[0]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(v[3]), "f"(v[2]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.z) : "f"(v[5]), "f"(v[4]));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.w) : "f"(v[7]), "f"(v[6]));
    return r;
}

// two fp32 -> packed fp16x2 / bf16x2 (a in the low half), round-to-nearest, saturating to the largest finite value
__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt) {
    uint32_t r;
    if (fmt == 0) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    else          asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// ---------------------------------------------------------------------------------------------- weight prepack
// w [Cout, Cin, ks, ks] fp32  ->  [n-tile][chunk][tap][plane(2)][BN][8] fp16/bf16, zero padded in Cout and Cin.
// flip_weight != 0: cross-correlation (taps as stored); 0: true convolution (taps mirrored).
// up2 != 0: the 3x3 weights are first convolved with the 4x4 FIR (gain 4) into a 6x6 composite and split into the four
// output-parity 3x3 kernels of the polyphase form; virtual channel v = phase * Cout + o.
struct PackParams {
    const float* w; const float* fir; void* out; int Cout, Cin, ks, BN, nchunks, ntaps, ntiles, flip_weight, fmt, up2, down2, im2col; float w_scale;
    // batched form (blockIdx.y = sample): per-sample weight sets `w_bstride` elements apart (0: one shared set) and / or per-sample input-channel
    // scales styles[sample, Cin] folded into the weights (the modulation of networks.py:64-66 applied while packing)
    long long w_bstride; const float* styles;
};

__device__ float composite_tap(const PackParams& p, int o, int c, int a, int b, int th, int tw) {
    // out[2y+a][2x+b] = sum_{th,tw} Kc[2*th+1-a][2*tw+1-b] * xpad1[y+th][x+tw],   Kc = w' (*) (4 k)   (6x6 full convolution)
    // k = f flipped (upfirdn2d applies f as a true convolution), w' = w mirrored when flip_weight == 0 (true convolution with w,
    // SynthesisLayer.conv0) or w as stored when flip_weight != 0.  Checked against the oracle to 2e-15 in fp64.
    const int u = 2 * th + 1 - a, v = 2 * tw + 1 - b;
    float acc = 0.f;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const int fi = u - i, fj = v - j;
            if (fi < 0 || fi > 3 || fj < 0 || fj > 3) continue;
            const int wi = p.flip_weight ? i : 2 - i, wj = p.flip_weight ? j : 2 - j;
            acc += p.w[((size_t)(o * p.Cin + c) * 3 + wi) * 3 + wj] * p.fir[(3 - fi) * 4 + (3 - fj)] * 4.f;
        }
    return acc;
}

// down-2 (FIR pad 2 -> 3x3 stride 2, conv2d_resample.py:119-122) as a 'same' 3x3 stride-1 conv over the space-to-depth planes of x:
//   Kd = w' (*) k (6x6 full convolution, k = f flipped, w' = w for correlation / w mirrored for true convolution)
//   virtual channel (a, c, b), tap (r, s)  ->  Kd[o, c, 2r+a, 2s+b]                         (checked against the oracle to 4e-15 in fp64)
__device__ float down2_tap(const PackParams& p, int o, int c, int a, int b, int r, int s_) {
    const int u = 2 * r + a, v = 2 * s_ + b;
    float acc = 0.f;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const int fp = u - i, fq = v - j;
            if (fp < 0 || fp > 3 || fq < 0 || fq > 3) continue;
            const int wi = p.flip_weight ? i : 2 - i, wj = p.flip_weight ? j : 2 - j;
            acc += p.w[((size_t)(o * p.Cin + c) * 3 + wi) * 3 + wj] * p.fir[(3 - fp) * 4 + (3 - fq)];
        }
    return acc;
}

__global__ void conv_prepack_kernel(PackParams p, long long out_sample_stride_halves) {
    const size_t total = (size_t)p.ntiles * p.nchunks * p.ntaps * 2 * p.BN * 8;
    const int nvirt = p.up2 ? 4 * p.Cout : p.Cout;
    const int sample = blockIdx.y;
    p.w += (size_t)sample * p.w_bstride;
    p.out = (void*)((uint16_t*)p.out + (size_t)sample * out_sample_stride_halves);      // (stride counted in 2-byte units for every format)
    const float* sty = p.styles ? p.styles + (size_t)sample * p.Cin : nullptr;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t r = idx;
        const int epp = p.fmt == 2 ? 4 : 8, npl = p.fmt == 2 ? 4 : 2;      // elements per 16-byte row, planes per 16-channel chunk
        const int e = r % epp; r /= epp;
        const int nl = r % p.BN; r /= p.BN;
        const int j = r % npl; r /= npl;
        const int tap = r % p.ntaps; r /= p.ntaps;
        const int ci = r % p.nchunks; r /= p.nchunks;
        const int jn = (int)r;
        const int v = jn * p.BN + nl, c = ci * kKC + j * epp + e;
        float val = 0.f;
        if (p.down2) {
            // c runs over the 4 * Cin virtual channels (row parity a, then real channel, then column parity b: the order in which the
            // activation loader's 8-byte reads deliver them); p.Cin is the real channel count
            if (v < p.Cout && c < 4 * p.Cin) {
                if (p.down2 == 2) {      // channel-blocked input (TMA): parity-major, virtual channel = (2a + b) * Cin + c_real
                    const int q = c / p.Cin;
                    val = down2_tap(p, v, c - q * p.Cin, q >> 1, q & 1, tap / 3, tap % 3);
                } else {
                    const int a = c / (2 * p.Cin), rem = c - a * 2 * p.Cin;      // virtual channel = a * 2Cin + 2 * c_real + b
                    val = down2_tap(p, v, rem >> 1, a, rem & 1, tap / 3, tap % 3);
                }
            }
        } else if (p.im2col) {
            // one "tap", Cin * ks * ks virtual channels in the weight tensor's memory order; a true convolution mirrors the taps
            const int kk = p.ks * p.ks;
            if (v < p.Cout && c < p.Cin * kk) {
                const int cr = c / kk, r2 = c - cr * kk;
                val = p.w[(size_t)(v * p.Cin + cr) * kk + (p.flip_weight ? r2 : kk - 1 - r2)];
            }
        } else if (v < nvirt && c < p.Cin) {
            const int kh = tap / p.ks, kw = tap % p.ks;
            if (p.up2) {
                const int phase = v / p.Cout, o = v % p.Cout;
                val = composite_tap(p, o, c, phase >> 1, phase & 1, kh, kw);
            } else {
                const int wh = p.flip_weight ? kh : p.ks - 1 - kh, ww = p.flip_weight ? kw : p.ks - 1 - kw;
                val = p.w[((size_t)(v * p.Cin + c) * p.ks + wh) * p.ks + ww];
            }
        }
        val *= p.w_scale;
        if (sty && val != 0.f) {
            // real input channel of GEMM channel c (down-2: (row parity, channel, column parity) order; folded taps: channel-major)
            const int cr = p.down2 == 2 ? c % p.Cin : p.down2 ? ((c % (2 * p.Cin)) >> 1) : (p.im2col ? c / (p.ks * p.ks) : c);
            val *= sty[cr];
        }
        if (p.fmt == 0)      ((__half*)p.out)[idx] = __float2half_rn(fminf(fmaxf(val, -65504.f), 65504.f));
        else if (p.fmt == 1) ((__nv_bfloat16*)p.out)[idx] = __float2bfloat16_rn(val);
        else                 ((uint32_t*)p.out)[idx] = to_tf32(val);
    }
}

// Plain layers (no resampling, no folded taps), tiled: one CTA packs 32 output channels x one 16-channel chunk x all taps.  Its source is 32 runs of
// 16 * k * k contiguous floats, read coalesced into shared memory once; its destination is, per (tap, plane), 32 consecutive 16-byte rows (512
// contiguous bytes).  The element-per-thread kernel above reads the weight tensor with a k * k element stride between neighbouring threads and
// decodes six indices per element: on the drop-in path, where the reference re-creates `weight * weight_gain` (and the per-sample modulated weights)
// on every call and every convolution therefore packs its weights anew, that kernel was a quarter of the step (7.7 of 31 ms).
constexpr int kPackRows = 32;
__global__ void __launch_bounds__(256) conv_prepack_tiled_kernel(PackParams p, long long out_sample_stride_halves) {
    extern __shared__ float psrc[];                                   // [kPackRows][16 * kk + 1]
    const int kk = p.ks * p.ks, rowlen = 16 * kk, pitch = rowlen + 1;
    const int rblocks = (p.BN + kPackRows - 1) / kPackRows;
    int b = blockIdx.x;
    const int rb = b % rblocks; b /= rblocks;
    const int ci = b % p.nchunks; const int jn = b / p.nchunks;
    const int sample = blockIdx.y;
    const float* w = p.w + (size_t)sample * p.w_bstride;
    const float* sty = p.styles ? p.styles + (size_t)sample * p.Cin : nullptr;
    const int c0 = ci * kKC, nl0 = rb * kPackRows;
    const int cvalid = p.Cin - c0 < 16 ? p.Cin - c0 : 16;            // channels of this chunk that exist
    for (int i = threadIdx.x; i < kPackRows * rowlen; i += 256) {
        const int r = i / rowlen, t = i - r * rowlen;
        const int v = jn * p.BN + nl0 + r;
        float val = 0.f;
        if (nl0 + r < p.BN && v < p.Cout && t < cvalid * kk) {
            val = __ldg(w + ((size_t)v * p.Cin + c0) * kk + t) * p.w_scale;
            if (sty) val *= sty[c0 + t / kk];
        }
        psrc[r * pitch + t] = val;
    }
    __syncthreads();
    const int epp = p.fmt == 2 ? 4 : 8, npl = p.fmt == 2 ? 4 : 2;
    uint16_t* out = (uint16_t*)p.out + (size_t)sample * out_sample_stride_halves;
    const size_t blk = ((size_t)(jn * p.nchunks + ci) * p.ntaps) * (size_t)(16 * p.BN);       // first element of this (n-tile, chunk) block
    const int per_tap = 16 * kPackRows;                               // elements this CTA writes per tap
    for (int i = threadIdx.x; i < p.ntaps * per_tap; i += 256) {
        const int tap = i / per_tap; int q = i - tap * per_tap;
        const int j = q / (kPackRows * epp); q -= j * (kPackRows * epp);
        const int r = q / epp, e = q - r * epp;
        if (nl0 + r >= p.BN) continue;
        const int cc = j * epp + e;
        const float val = psrc[r * pitch + cc * kk + (p.flip_weight ? tap : kk - 1 - tap)];
        const size_t dst = blk + ((size_t)(tap * npl + j) * p.BN + nl0 + r) * epp + e;
        if (p.fmt == 0)      ((__half*)out)[dst] = __float2half_rn(fminf(fmaxf(val, -65504.f), 65504.f));
        else if (p.fmt == 1) ((__nv_bfloat16*)out)[dst] = __float2bfloat16_rn(val);
        else                 ((uint32_t*)out)[dst] = to_tf32(val);
    }
}

// Up-2 composite weights, one thread per (output channel, input channel): the 6x6 composite Kc = w' (*) (4 k) is formed once in registers (144 FMAs)
// and its 36 entries are the 4 output-parity phases x 9 taps of the polyphase form (composite_tap above evaluates each of the 36 with its own 9-term
// loop, index arithmetic and loads: ~13x the instructions).  This matters where the weights are packed per call: the reference's fused modulated
// convolution hands conv2d_resample a fresh [N*O, I, 3, 3] tensor every step (networks.py:84-94), 550 M packed elements per generator forward.
__global__ void __launch_bounds__(256) conv_prepack_up2_kernel(PackParams p, long long out_sample_stride_halves) {
    const int cin_pad = p.nchunks * kKC;
    const long long total = (long long)p.Cout * cin_pad;
    const int sample = blockIdx.y;
    const float* w = p.w + (size_t)sample * p.w_bstride;
    const float* sty = p.styles ? p.styles + (size_t)sample * p.Cin : nullptr;
    uint16_t* out = (uint16_t*)p.out + (size_t)sample * out_sample_stride_halves;
    float ff[4][4];                                                    // ff[a][b] = 4 * fir[3 - a][3 - b]  (upfirdn2d applies f as a true convolution; gain 4)
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) ff[a][b] = 4.f * __ldg(p.fir + (3 - a) * 4 + (3 - b));
    const int epp = p.fmt == 2 ? 4 : 8, npl = p.fmt == 2 ? 4 : 2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cin_pad), o = (int)(idx / cin_pad);
        float kc[6][6];
#pragma unroll
        for (int u = 0; u < 6; u++)
#pragma unroll
            for (int v = 0; v < 6; v++) kc[u][v] = 0.f;
        if (c < p.Cin) {
            const float* wp = w + ((size_t)o * p.Cin + c) * 9;
            const float scale = p.w_scale * (sty ? sty[c] : 1.f);
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const float wv = __ldg(wp + (p.flip_weight ? i * 3 + j : (2 - i) * 3 + (2 - j))) * scale;
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) kc[i + a][j + b] = fmaf(wv, ff[a][b], kc[i + a][j + b]);
                }
        }
        const int ci = c / kKC, cc = c - ci * kKC, j = cc / epp, e = cc - j * epp;
#pragma unroll
        for (int phase = 0; phase < 4; phase++) {
            const int v = phase * p.Cout + o, jn = v / p.BN, nl = v - jn * p.BN;
            const size_t base = ((size_t)(jn * p.nchunks + ci) * p.ntaps) * (size_t)(16 * p.BN) + ((size_t)j * p.BN + nl) * epp + e;
#pragma unroll
            for (int th = 0; th < 3; th++)
#pragma unroll
                for (int tw = 0; tw < 3; tw++) {
                    const float val = kc[2 * th + 1 - (phase >> 1)][2 * tw + 1 - (phase & 1)];
                    const size_t dst = base + (size_t)((th * 3 + tw) * npl) * p.BN * epp;
                    if (p.fmt == 0)      ((__half*)out)[dst] = __float2half_rn(fminf(fmaxf(val, -65504.f), 65504.f));
                    else if (p.fmt == 1) ((__nv_bfloat16*)out)[dst] = __float2bfloat16_rn(val);
                    else                 ((uint32_t*)out)[dst] = to_tf32(val);
                }
        }
    }
}

struct MmaRing { int sa, sb; uint32_t pa, pb; };

// MMA issue loop of one CTA, executed by ALL lanes of the MMA warp (uniform control flow); `issue` != 0 on the one lane that issues.
// KS = kernel size (1 or 3); B ring slots hold KS taps each; NACC accumulators of BN columns.  Everything the tcgen05.mma needs is either a
// kernel parameter, a compile-time constant or a warp-uniform loop counter, so one MMA costs two uniform adds.
template <int KS, int NACC>
__device__ __forceinline__ void mma_issue_loop(const ConvParams& p, uint32_t a_base, uint32_t b_base, uint32_t a_full, uint32_t a_empty,
                                               uint32_t b_full, uint32_t b_empty, uint32_t acc_full, uint32_t tmem_base, uint32_t issue, MmaRing& ring,
                                               const uint32_t a_off16 = 0u) {
    // a_off16: first staged 16-byte row of this tile inside a stage (TMA mode: the box starts at a row boundary of the image, the tile does not)
    const uint32_t hi = (128u >> 4) | (1u << 14);                         // SBO = 128 B, descriptor version 1 (bits 32..47)
    const uint32_t a_lo_const = (p.a_lbo16 & 0x3FFF) << 16;               // LBO = distance between the two 8-channel planes of a stage (>> 4)
    const uint32_t b_lo_const = ((uint32_t)p.BN & 0x3FFF) << 16;          // LBO = BN * 16 B  (>> 4)
    const uint32_t b_tile16 = p.b_tile_bytes >> 4;
    const uint32_t bn = (uint32_t)p.BN;
    const uint32_t pw = (uint32_t)p.PW;
    const uint32_t idesc = p.idesc;
    const int nchunks = p.nchunks, SA = p.SA, SB = p.SB;
    int sa = ring.sa, sb = ring.sb; uint32_t pa = ring.pa, pb = ring.pb;   // stage / slot cursors continue across the tiles of a persistent CTA
    long long wait_a = 0, wait_b = 0;              // instrumentation (registers; written once at the end)
    for (int ci = 0; ci < nchunks; ci++) {
        long long t0 = PG_DBG(p) ? clock64() : 0;
        mbar_wait(a_full + 8u * (uint32_t)sa, pa);
        tc_fence_after();
        if (PG_DBG(p)) { if (ci == 0) { if (issue) PG_TS(2); } else wait_a += clock64() - t0; }
        const uint32_t a_lo = (a_lo_const | ((a_base + (uint32_t)sa * p.a_stage_bytes) >> 4)) + a_off16;
#pragma unroll
        for (int kh = 0; kh < KS; kh++) {
            t0 = PG_DBG(p) ? clock64() : 0;
            mbar_wait(b_full + 8u * (uint32_t)sb, pb);
            tc_fence_after();
            if (PG_DBG(p)) wait_b += clock64() - t0;
            const uint32_t b_lo = b_lo_const | ((b_base + (uint32_t)sb * p.b_slot_bytes) >> 4);
            if (issue) {
#pragma unroll
                for (int kw = 0; kw < KS; kw++) {
                    const uint32_t s0 = (KS == 3) ? (uint32_t)kh * pw + (uint32_t)kw : 0u;     // tap shift in strip positions == 16-byte rows
                    const uint64_t bdesc = ((uint64_t)hi << 32) | (uint64_t)(b_lo + (uint32_t)kw * b_tile16);
                    const uint32_t acc_flag = (ci | kh | kw) ? 1u : 0u;
                    if (p.fmt == 2) {
                        // tf32: K = 8 per MMA, a 16-channel chunk is two K steps; step 1 starts two 4-channel planes further into the stage / weight tile
#pragma unroll
                        for (int kstep = 0; kstep < 2; kstep++) {
                            const uint64_t bd = bdesc + (uint64_t)((uint32_t)kstep * 2u * bn);
#pragma unroll
                            for (int a = 0; a < NACC; a++) {
                                const uint64_t adesc = ((uint64_t)hi << 32) | (uint64_t)(a_lo + s0 + (uint32_t)a * 128u + (uint32_t)kstep * 2u * p.a_lbo16);
                                umma_tf32(tmem_base + (uint32_t)a * bn, adesc, bd, idesc, acc_flag | (uint32_t)kstep);
                            }
                        }
                        continue;
                    }
#pragma unroll
                    for (int a = 0; a < NACC; a++) {
                        const uint64_t adesc = ((uint64_t)hi << 32) | (uint64_t)(a_lo + s0 + (uint32_t)a * 128u);
                        umma_f16(tmem_base + (uint32_t)a * bn, adesc, bdesc, idesc, acc_flag);
                    }
                }
                umma_commit(b_empty + 8u * (uint32_t)sb);
                if (kh == KS - 1) umma_commit(a_empty + 8u * (uint32_t)sa);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; pb ^= 1; }
        }
        if (++sa == SA) { sa = 0; pa ^= 1; }
    }
    if (issue) { PG_TS(3); PG_PUT(8, wait_a); PG_PUT(9, wait_b); }
    if (issue) umma_commit(acc_full);
    ring.sa = sa; ring.sb = sb; ring.pa = pa; ring.pb = pb;
}

// Converter task stream of one warp: slots (chunk, idx), idx < tpw, in register batches of BATCH tasks of TREGS floats.  pipelined: the loads of
// batch k + 1 are issued before batch k is converted and stored (two batches of registers); otherwise one batch per memory round trip.
template <int TREGS, int BATCH, class Load, class Store>
__device__ __forceinline__ void stream_tasks(Load&& load_task, Store&& store_task, const int tpw, const int nchunks, const bool pipelined,
                                             const bool timed, long long& t_issue, long long& t_store) {
    float va[BATCH][TREGS], vb[BATCH][TREGS];
    int l_ci = 0, l_idx = 0, c_ci = 0, c_idx = 0;          // load cursor, convert/store cursor
    auto advance = [&](int& ci, int& idx) { if (++idx == tpw) { idx = 0; ci++; } };
#pragma unroll
    for (int u = 0; u < BATCH; u++) { load_task(va[u], l_ci, l_idx); advance(l_ci, l_idx); }
    while (!pipelined && c_ci < nchunks) {
        const long long t0 = timed ? clock64() : 0;
#pragma unroll
        for (int u = 0; u < BATCH; u++) { store_task(va[u], c_ci, c_idx); advance(c_ci, c_idx); }
        const long long t1 = timed ? clock64() : 0;
#pragma unroll
        for (int u = 0; u < BATCH; u++) { load_task(va[u], l_ci, l_idx); advance(l_ci, l_idx); }
        if (timed) { t_store += t1 - t0; t_issue += clock64() - t1; }
    }
    while (c_ci < nchunks) {
#pragma unroll
        for (int u = 0; u < BATCH; u++) { load_task(vb[u], l_ci, l_idx); advance(l_ci, l_idx); }
#pragma unroll
        for (int u = 0; u < BATCH; u++) { store_task(va[u], c_ci, c_idx); advance(c_ci, c_idx); }
        if (c_ci >= nchunks) break;
#pragma unroll
        for (int u = 0; u < BATCH; u++) { load_task(va[u], l_ci, l_idx); advance(l_ci, l_idx); }
#pragma unroll
        for (int u = 0; u < BATCH; u++) { store_task(vb[u], c_ci, c_idx); advance(c_ci, c_idx); }
    }
}

// Lean A converter for the aligned 8-byte loader.  Everything that depends only on (warp, lane, task slot) -- the element offset of the lane's
// pair inside a channel plane and the two shared-memory slots it lands on -- is computed once per CTA; per chunk only the channel base moves.
// A warp owns task slots idx = r * PER + u (task = cw + 8 * idx: 32 pairs of one 8-channel plane; all of a warp's tasks share the plane cw & 1),
// processed in ROUNDS register batches of PER tasks: the loads of a batch are issued before the stage is waited for.
template <bool SCALE, bool IN_HALF, int PER, int ROUNDS>
__device__ __forceinline__ void convert_vec2(const ConvParams& p, const float* xn, const float* xn2, const int HW, const int cw, const int lane,
                                             const int g_lo, const int g_hi, const int ntasks, const int q0, uint8_t* a_base, const float* s_style,
                                             uint64_t* a_full, uint64_t* a_empty, long long& wait_e,
                                             const int nwarps = kConvWarps, const int chunk_base = 0, const bool zero_pads = false, const int band = 0) {
    // nwarps: converter warps of the CTA; chunk_base: chunks converted by this CTA before this tile (persistent CTAs: the stage ring continues);
    // zero_pads: the padding slots of the strip move from tile to tile, so the first use of every stage in a tile re-zeroes them
    constexpr int NS = PER * ROUNDS;
    const int wpg = nwarps / p.cgroups;         // warps per group; group g converts chunks g, g + cgroups, ...
    const int grp = cw / wpg, gw = cw - grp * wpg;
    const int plane = cw & 1;
    const bool swap = (lane >> 2) & 1;          // lanes sit 32 B apart in the stage: lanes 4..7 of each group of 8 write their second slot first
    const int dpitch = p.PW - p.W;
    int goff[NS]; uint32_t sA[NS], sB[NS];
#pragma unroll
    for (int idx = 0; idx < NS; idx++) {
        const int tt = gw + idx * wpg;                               // wpg is even: every task of a warp has the plane cw & 1
        const int g = g_lo + (tt >> 1) * 32 + lane;
        bool ok = tt < ntasks && g < g_hi;
        const int e = 2 * g;
        const int h = (int)__umulhi((uint32_t)e, p.w_magic);
        const int s0 = e + h * dpitch - q0;                          // staged slot of the first element; the second is s0 + 1 (same row)
        const int s1st = swap ? s0 + 1 : s0, s2nd = swap ? s0 : s0 + 1;
        int eoff = e;                                                // element offset inside the channel plane
        if (p.band_tw) {
            // band mode: e indexes the band's own strip (rows of W = band_tw + 4 columns); column 0 of the band is image column band * band_tw - 2
            // (even, so pairs are whole); pairs outside the image stay zero
            const int wimg = band * p.band_tw - 2 + (e - h * p.W);
            ok = ok && wimg >= 0 && wimg < p.Wimg;
            eoff = h * p.Wimg + wimg;
        }
        goff[idx] = (ok && !(PG_DBGMODE(p) & 1)) ? eoff : -1;
        const bool st_ok = ok && !(PG_DBGMODE(p) & 2);
        sA[idx] = (st_ok && s1st >= 0 && s1st < p.PA) ? (uint32_t)((plane * p.PA + s1st) * 16) : 0xffffffffu;
        sB[idx] = (st_ok && s2nd >= 0 && s2nd < p.PA) ? (uint32_t)((plane * p.PA + s2nd) * 16) : 0xffffffffu;
    }
    const bool has_in_act = p.in_act != PG_ACT_LINEAR;
    const float in_slope = (p.in_act == PG_ACT_RELU) ? 0.f : p.in_alpha;
    const int nchunks = p.nchunks;
    for (int ci = grp; ci < nchunks; ci += p.cgroups) {
        const int gci = chunk_base + ci;
        const int st = gci % p.SA; const uint32_t ph = (uint32_t)(gci / p.SA) & 1u;
        const int c0 = ci * kKC + plane * 8;
        const int nval = p.Cin - c0;                                  // channels of this group that exist (>= 8: all)
        const float* cb = (c0 < p.cin1 ? xn : xn2) + (size_t)c0 * HW;
        const __half* cbh = reinterpret_cast<const __half*>(xn) + (size_t)c0 * HW;      // IN_HALF: x is fp16, one sample's planes start at xn
        uint8_t* stage = a_base + (size_t)st * p.a_stage_bytes;
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            float v[PER][IN_HALF ? 8 : 16];                      // IN_HALF: v[u][i] holds the half2 (positions e, e + 1) of channel i as raw bits
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int off = goff[r * PER + u];
#pragma unroll
                for (int i = 0; i < (IN_HALF ? 8 : 16); i++) v[u][i] = 0.f;
                if (off >= 0) {
                    if (IN_HALF) {
                        const __half* src = cbh + off;
#pragma unroll
                        for (int i = 0; i < 8; i++)
                            if (i < nval) v[u][i] = __uint_as_float(__ldg(reinterpret_cast<const unsigned int*>(src + (size_t)i * HW)));
                    } else {
                        const float* src = cb + off;
#pragma unroll
                        if (PG_LDMODE(p) == 1) {
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                if (i < nval) asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v[u][i]), "=f"(v[u][8 + i]) : "l"(src + (size_t)i * HW));
                        } else if (PG_LDMODE(p) == 2) {
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                if (i < nval) asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v[u][i]), "=f"(v[u][8 + i]) : "l"(src + (size_t)i * HW));
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                if (i < nval) { const float2 t = __ldg(reinterpret_cast<const float2*>(src + (size_t)i * HW)); v[u][i] = t.x; v[u][8 + i] = t.y; }
                        }
                    }
                }
            }
            if (r == 0) {
                const long long t0 = PG_DBG(p) ? clock64() : 0;
                mbar_wait(smem_u32(&a_empty[st]), ph ^ 1);
                if (PG_DBG(p)) wait_e += clock64() - t0;
                if (zero_pads && ci < p.SA) {
                    for (int sl = gw * 32 + lane; sl < p.PA; sl += wpg * 32) {
                        const int q = q0 + sl;
                        bool pad = q < 0 || q >= p.Lp;
                        if (!pad) { const int hq = (int)__umulhi((uint32_t)q, p.pw_magic); pad = q - hq * p.PW >= p.W; }
                        if (pad) {
                            *reinterpret_cast<uint4*>(stage + (size_t)sl * 16) = make_uint4(0u, 0u, 0u, 0u);
                            *reinterpret_cast<uint4*>(stage + ((size_t)p.PA + sl) * 16) = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t a1 = sA[r * PER + u], a2 = sB[r * PER + u];
                if ((a1 & a2) == 0xffffffffu) continue;
                uint4 lo, hi;
                if (IN_HALF) {
                    // already fp16: gather the low (position e) / high (position e + 1) halves of channel pairs into the two 16-byte rows
                    const unsigned int b0 = __float_as_uint(v[u][0]), b1 = __float_as_uint(v[u][1]), b2 = __float_as_uint(v[u][2]), b3 = __float_as_uint(v[u][3]),
                                       b4 = __float_as_uint(v[u][4]), b5 = __float_as_uint(v[u][5]), b6 = __float_as_uint(v[u][6]), b7 = __float_as_uint(v[u][7]);
                    lo.x = __byte_perm(b0, b1, 0x5410); lo.y = __byte_perm(b2, b3, 0x5410); lo.z = __byte_perm(b4, b5, 0x5410); lo.w = __byte_perm(b6, b7, 0x5410);
                    hi.x = __byte_perm(b0, b1, 0x7632); hi.y = __byte_perm(b2, b3, 0x7632); hi.z = __byte_perm(b4, b5, 0x7632); hi.w = __byte_perm(b6, b7, 0x7632);
                } else {
                    if (SCALE) {
                        const float* sc = s_style + c0;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            float a = v[u][i], b = v[u][(IN_HALF ? 0 : 8) + i];
                            if (has_in_act) { a = fmaxf(a, 0.f) + in_slope * fminf(a, 0.f); b = fmaxf(b, 0.f) + in_slope * fminf(b, 0.f); }
                            v[u][i] = a * sc[i]; v[u][(IN_HALF ? 0 : 8) + i] = b * sc[i];
                        }
                    }
                    constexpr int H8 = IN_HALF ? 0 : 8;
                    lo.x = pack2(v[u][0], v[u][1], p.fmt); lo.y = pack2(v[u][2], v[u][3], p.fmt); lo.z = pack2(v[u][4], v[u][5], p.fmt); lo.w = pack2(v[u][6], v[u][7], p.fmt);
                    hi.x = pack2(v[u][H8 + 0], v[u][H8 + 1], p.fmt); hi.y = pack2(v[u][H8 + 2], v[u][H8 + 3], p.fmt);
                    hi.z = pack2(v[u][H8 + 4], v[u][H8 + 5], p.fmt); hi.w = pack2(v[u][H8 + 6], v[u][H8 + 7], p.fmt);
                }
                const uint4 d1 = swap ? hi : lo, d2 = swap ? lo : hi;
                if (a1 != 0xffffffffu) *reinterpret_cast<uint4*>(stage + a1) = d1;
                if (a2 != 0xffffffffu) *reinterpret_cast<uint4*>(stage + a2) = d2;
            }
        }
        if (!(PG_DBGMODE(p) & 16)) fence_proxy_async();        // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&a_full[st]));
    }
}

// Lean A converter for the down-2 form (space-to-depth view, virtual channels ordered (row parity a, real channel, column parity b)): a task is
// 32 strip positions x 8 virtual channels = 4 real channels x both column parities of input row 2h + a, read as four aligned 8-byte loads per lane.
// As in convert_vec2 the per-(warp, lane, slot) geometry -- input element offset of the lane's output position, its shared-memory slot -- is
// computed once per CTA; per chunk only the channel / row-parity base moves.  Every staged slot is written (zeros where out of range).
template <bool SCALE, int PER, int ROUNDS>
__device__ __forceinline__ void convert_down2(const ConvParams& p, const float* xn, const int cw, const int lane, const int ntasks, const int q0, const int band,
                                              uint8_t* a_base, const float* s_style, uint64_t* a_full, uint64_t* a_empty, long long& wait_e) {
    constexpr int NS = PER * ROUNDS;
    const int plane = cw & 1;                   // kConvWarps is even: every task cw + 8 * idx of a warp has the same plane
    int soff[NS]; uint32_t slot[NS];
#pragma unroll
    for (int idx = 0; idx < NS; idx++) {
        const int tt = cw + idx * kConvWarps;
        const int spos = (tt >> 1) * 32 + lane;
        const int q = q0 + spos;
        bool ok = tt < ntasks && q >= 0 && q < p.Lp && !(PG_DBGMODE(p) & 1);
        int h = 0, w = 0;
        if (ok) { h = (int)__umulhi((uint32_t)q, p.pw_magic); w = q - h * p.PW; ok = w < p.W; }
        if (p.band_tw) { w = band * p.band_tw - 2 + w; ok = ok && w >= 0 && w < p.Wimg; }
        soff[idx] = ok ? 2 * h * p.win + 2 * w : -1;                                  // input element of (row 2h, column 2w); + a * win per chunk
        slot[idx] = (tt < ntasks && !(PG_DBGMODE(p) & 2)) ? (uint32_t)((plane * p.PA + spos) * 16) : 0xffffffffu;
    }
    const bool has_in_act = p.in_act != PG_ACT_LINEAR;
    const float in_slope = (p.in_act == PG_ACT_RELU) ? 0.f : p.in_alpha;
    const size_t cs = (size_t)p.hin * p.win;
    const int nchunks = p.nchunks;
    int st = 0; uint32_t ph = 0;
    for (int ci = 0; ci < nchunks; ci++) {
        const int c0 = ci * kKC + plane * 8;
        const int a = c0 / (2 * p.cin_real), cr = (c0 - a * 2 * p.cin_real) >> 1;
        const float* cb = xn + (size_t)cr * cs + (size_t)a * p.win;
        uint8_t* stage = a_base + (size_t)st * p.a_stage_bytes;
#pragma unroll
        for (int r = 0; r < ROUNDS; r++) {
            float v[PER][8];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int off = soff[r * PER + u];
#pragma unroll
                for (int i = 0; i < 8; i++) v[u][i] = 0.f;
                if (off >= 0) {
                    const float* src = cb + off;
#pragma unroll
                    for (int i = 0; i < 4; i++) { const float2 t = __ldg(reinterpret_cast<const float2*>(src + (size_t)i * cs)); v[u][2 * i] = t.x; v[u][2 * i + 1] = t.y; }
                }
            }
            if (r == 0) {
                const long long t0 = PG_DBG(p) ? clock64() : 0;
                mbar_wait(smem_u32(&a_empty[st]), ph ^ 1);
                if (PG_DBG(p)) wait_e += clock64() - t0;
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const uint32_t sl = slot[r * PER + u];
                if (sl == 0xffffffffu) continue;
                if (SCALE) {
                    const float* sc = s_style + c0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        float x = v[u][i];
                        if (has_in_act) x = fmaxf(x, 0.f) + in_slope * fminf(x, 0.f);
                        v[u][i] = x * sc[i];
                    }
                }
                uint4 pk;
                pk.x = pack2(v[u][0], v[u][1], p.fmt); pk.y = pack2(v[u][2], v[u][3], p.fmt);
                pk.z = pack2(v[u][4], v[u][5], p.fmt); pk.w = pack2(v[u][6], v[u][7], p.fmt);
                *reinterpret_cast<uint4*>(stage + sl) = pk;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&a_full[st]));
        if (++st == p.SA) { st = 0; ph ^= 1; }
    }
}

// Epilogue of one output tile for one warp: TMEM lane quarter `quarter`, 16-column chunks part, part + step, ... of every accumulator.
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, const int n, const int jn, const int m0, const int HW, const uint32_t tmem_base,
                                              const float* s_scale, const float* s_shift, const int quarter, const int part, const int step, const int lane,
                                              const int band = 0) {
    const int ncol_chunks = p.BN / 16;
    const float slope = (p.act == PG_ACT_LINEAR) ? 1.f : (p.act == PG_ACT_RELU ? 0.f : p.alpha);   // act(v) = max(v,0) + slope*min(v,0)
    const float cl = p.clamp >= 0.f ? p.clamp : __int_as_float(0x7f800000);
    const bool do_act = p.act != PG_ACT_LINEAR, do_clamp = p.clamp >= 0.f;
    const int W2 = 2 * p.Wimg;
    for (int a = 0; a < p.NACC; a++) {
        const int q = m0 + a * 128 + quarter * 32 + lane;
        const int h = (int)__umulhi((uint32_t)q, p.pw_magic), ws = q - h * p.PW;      // row, column inside the strip
        const int w = p.band_tw ? band * p.band_tw + ws - 2 : ws;                      // image column
        const bool ok = q < p.Lp && (p.band_tw ? (ws >= 2 && ws < p.band_tw + 2 && w < p.Wimg) : ws < p.W);
        if (p.spade) {
            // N tile jn holds gamma of channels [jn * Ct, (jn + 1) * Ct) in columns [0, Ct) and their beta in columns [Ct, 2 Ct) of the same accumulator
            // row (Ct = BN / 2; one tile when 2C <= BN); normalise x with the staged statistics
            const int C = p.cout_real >> 1, Ct = p.BN >> 1, cbase = jn * Ct;
            const float* xp = p.sp_x + ((size_t)n * C + cbase) * HW + (size_t)h * p.Wimg + w;
            const size_t yoff = ((size_t)n * C + cbase) * HW + (size_t)h * p.Wimg + w;
            for (int cc = part; cc < Ct / 16; cc += step) {
                uint32_t rg[16], rb[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * p.BN + cc * 16), rg);
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * p.BN + Ct + cc * 16), rb);
                if (!ok) continue;
                float xv[16];
#pragma unroll
                for (int i = 0; i < 16; i++) xv[i] = __ldg(xp + (size_t)(cc * 16 + i) * HW);
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float xn_ = fmaf(xv[i], s_scale[cc * 16 + i], s_shift[cc * 16 + i]);
                    v[i] = fmaf(xn_, 1.f + __uint_as_float(rg[i]), __uint_as_float(rb[i]));
                    v[i] = (fmaxf(v[i], 0.f) + slope * fminf(v[i], 0.f)) * p.gain;
                }
                if (p.y_c8) {
                    uint4* yb = reinterpret_cast<uint4*>(p.y) + ((size_t)n * p.cb_out + (size_t)(cbase + cc * 16) / 8) * HW + (size_t)h * p.Wimg + w;
                    yb[0] = pack_half8(v); yb[HW] = pack_half8(v + 8);
                } else if (p.out_half) {
#pragma unroll
                    for (int i = 0; i < 16; i++) store_half(p.y, yoff + (size_t)(cc * 16 + i) * HW, v[i]);
                } else {
                    float* yp = p.y + yoff;
#pragma unroll
                    for (int i = 0; i < 16; i++) yp[(size_t)(cc * 16 + i) * HW] = v[i];
                }
            }
            continue;
        }
        float nz0 = 0.f;
        if (ok && p.noise && !p.up2) nz0 = __ldg(p.noise + (size_t)n * p.noise_bstride + (size_t)h * p.Wimg + w) * p.gain;
        // output element offset / channel stride / valid channels of column chunk cc for this thread's position
        auto chunk_out = [&](int cc, size_t& off, size_t& ystride, int& oy, int& ox) {
            const int v0 = jn * p.BN + cc * 16;                  // first (virtual) output channel of this chunk
            int nvalid = p.Cout - v0; nvalid = nvalid > 16 ? 16 : nvalid;
            if (!p.up2) {
                off = ((size_t)n * p.Cout + v0) * HW + (size_t)h * p.Wimg + w;
                ystride = (size_t)HW; oy = h; ox = w;
            } else {
                // polyphase up-2: virtual channel = phase * Cout + o (Cout % 16 == 0, so a chunk has one phase); output is 2H x 2W
                const int phase = v0 / p.cout_real, o0 = v0 - phase * p.cout_real;
                oy = 2 * h + (phase >> 1); ox = 2 * w + (phase & 1);
                ystride = (size_t)4 * HW;
                off = ((size_t)n * p.cout_real + o0) * ystride + (size_t)oy * W2 + ox;
            }
            return nvalid;
        };
        // the residual of chunk cc + step is fetched while chunk cc is read from TMEM, transformed and stored (one memory round trip hidden).  It stays
        // in its raw form (16 floats, or 8 words of packed halves) until it is added: converting at fetch time would wait for the load on the spot
        uint32_t res[16];
#pragma unroll
        for (int i = 0; i < 16; i++) res[i] = 0u;
        auto fetch_res = [&](int cc, uint32_t (&dst)[16]) {
            if (!p.residual || !ok || cc >= ncol_chunks) return;
            if (p.res_c8) {
                // channel-blocked fp16 residual: the chunk's 16 channels are two 16-byte rows at this pixel
                const uint4* rb = reinterpret_cast<const uint4*>(p.residual) + ((size_t)n * p.cb_out + (size_t)(jn * p.BN + cc * 16) / 8) * HW + (size_t)h * p.Wimg + w;
                const uint4 q0 = __ldg(rb), q1 = __ldg(rb + HW);
                dst[0] = q0.x; dst[1] = q0.y; dst[2] = q0.z; dst[3] = q0.w; dst[4] = q1.x; dst[5] = q1.y; dst[6] = q1.z; dst[7] = q1.w;
                return;
            }
            size_t off, ystride; int oy, ox;
            const int nvalid = chunk_out(cc, off, ystride, oy, ox);
#pragma unroll
            for (int i = 0; i < 16; i++) if (i < nvalid) dst[i] = __float_as_uint(__ldg(p.residual + off + (size_t)i * ystride));
        };
        fetch_res(part, res);
        for (int cc = part; cc < ncol_chunks; cc += step) {
            uint32_t res_next[16];
#pragma unroll
            for (int i = 0; i < 16; i++) res_next[i] = 0u;
            fetch_res(cc + step, res_next);
            uint32_t r[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * p.BN + cc * 16), r);
            if (!ok) continue;
            size_t off, ystride; int oy, ox;
            const int nvalid = chunk_out(cc, off, ystride, oy, ox);
            if (nvalid <= 0) continue;
            float nz = nz0;
            if (p.up2 && p.noise) nz = __ldg(p.noise + (size_t)n * p.noise_bstride + (size_t)oy * W2 + ox) * p.gain;
            float v[16];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float4 a4 = reinterpret_cast<const float4*>(s_scale + cc * 16)[i];
                const float4 b4 = reinterpret_cast<const float4*>(s_shift + cc * 16)[i];
                v[4 * i]     = fmaf(__uint_as_float(r[4 * i]),     a4.x, b4.x + nz);
                v[4 * i + 1] = fmaf(__uint_as_float(r[4 * i + 1]), a4.y, b4.y + nz);
                v[4 * i + 2] = fmaf(__uint_as_float(r[4 * i + 2]), a4.z, b4.z + nz);
                v[4 * i + 3] = fmaf(__uint_as_float(r[4 * i + 3]), a4.w, b4.w + nz);
            }
            if (do_act) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f) + slope * fminf(v[i], 0.f);
            }
            if (do_clamp) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = fminf(fmaxf(v[i], -cl), cl);
            }
            if (p.res_c8) {
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&res[e]));
                    v[2 * e] += f2.x; v[2 * e + 1] += f2.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] += __uint_as_float(res[i]);
            }
            float* yp = p.y + off;
            if (p.y_c8) {
                // channel-blocked fp16: the 16 channels of this chunk are two 16-byte rows, consecutive lanes write consecutive rows (host side
                // guarantees Cout % 16 == 0 and no up-2)
                uint4* yb = reinterpret_cast<uint4*>(p.y) + ((size_t)n * p.cb_out + (size_t)(jn * p.BN + cc * 16) / 8) * HW + (size_t)h * p.Wimg + w;
                yb[0] = pack_half8(v); yb[HW] = pack_half8(v + 8);
            } else if (p.out_half) {
#pragma unroll
                for (int i = 0; i < 16; i++) if (i < nvalid) store_half(p.y, off + (size_t)i * ystride, v[i]);
            } else if (nvalid == 16) {
                if (PG_LDMODE(p) == 3) {                              // tuning: streaming (evict-first) stores
#pragma unroll
                    for (int i = 0; i < 16; i++) __stcs(yp + (size_t)i * ystride, v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) yp[(size_t)i * ystride] = v[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) if (i < nvalid) yp[(size_t)i * ystride] = v[i];
            }
#pragma unroll
            for (int i = 0; i < 16; i++) res[i] = res_next[i];
        }
    }
}

// Residual prefetch.  The epilogue reads the residual one 16-channel chunk ahead of its use, i.e. ~1 KB in flight per warp: with one persistent CTA
// per SM that is ~1 MB in flight on the whole GPU, and the layer becomes bound by DRAM latency instead of bandwidth (128->128 @128^2 with a residual:
// 135 us against 73 us without).  The epilogue warps therefore pull the tile's residual lines into L2 BEFORE they wait for the accumulator (the main
// loop of the tile is still running); the demand loads of epilogue_tile then hit L2.  One lane per 128-byte line issues the prefetch.
__device__ __forceinline__ void prefetch_residual_tile(const ConvParams& p, const int n, const int jn, const int m0, const int HW, const int quarter,
                                                       const int part, const int step, const int lane, const int band = 0) {
    if (!p.residual || p.up2 || p.spade) return;
    const int ncol_chunks = p.BN / 16;
    for (int a = 0; a < p.NACC; a++) {
        const int q = m0 + a * 128 + quarter * 32 + lane;
        const int h = (int)__umulhi((uint32_t)q, p.pw_magic), ws = q - h * p.PW;
        const int w = p.band_tw ? band * p.band_tw + ws - 2 : ws;
        const bool ok = q < p.Lp && (p.band_tw ? (ws >= 2 && ws < p.band_tw + 2 && w < p.Wimg) : ws < p.W);
        if (!ok) continue;
        const size_t pix = (size_t)h * p.Wimg + w;
        if (p.res_c8) {
            if (lane != 0 && (pix & 7) != 0) continue;                       // 8 pixels x 16 B per line
            const uint4* rb = reinterpret_cast<const uint4*>(p.residual) + ((size_t)n * p.cb_out + (size_t)(jn * p.BN) / 8) * HW + pix;
            for (int cc = part; cc < ncol_chunks; cc += step) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rb + (size_t)(2 * cc) * HW));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rb + (size_t)(2 * cc + 1) * HW));
            }
        } else {
            if (lane != 0 && (pix & 31) != 0) continue;                      // 32 pixels x 4 B per line
            const int v0 = jn * p.BN;
            const float* rp = p.residual + ((size_t)n * p.Cout + v0) * HW + pix;
            for (int cc = part; cc < ncol_chunks; cc += step) {
                const int nv = p.Cout - v0 - cc * 16;
#pragma unroll 4
                for (int i = 0; i < 16; i++)
                    if (i < nv) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)(cc * 16 + i) * HW));
            }
        }
    }
}

// Polyphase up-2 epilogue when an N tile holds both x-phases of a row parity (BN % (2 Cout) == 0: Cout <= 128): a thread reads the two accumulator
// chunks (phase 2py, phase 2py + 1) of 16 output channels and writes the two horizontally adjacent output pixels (2h + py, 2w), (2h + py, 2w + 1)
// together -- 8 contiguous bytes per channel in fp32 NCHW, 32 contiguous bytes per 8-channel block in the channel-blocked fp16 layout -- where the
// one-phase-at-a-time path writes every other element (half-used sectors).
__device__ __forceinline__ void epilogue_tile_up2_pair(const ConvParams& p, const int n, const int jn, const int m0, const int HW, const uint32_t tmem_base,
                                                       const float* s_scale, const float* s_shift, const int quarter, const int part, const int step,
                                                       const int lane, const int band) {
    const float slope = (p.act == PG_ACT_LINEAR) ? 1.f : (p.act == PG_ACT_RELU ? 0.f : p.alpha);
    const float cl = p.clamp >= 0.f ? p.clamp : __int_as_float(0x7f800000);
    const bool do_act = p.act != PG_ACT_LINEAR, do_clamp = p.clamp >= 0.f;
    const int W2 = 2 * p.Wimg, H2W2 = 4 * HW;
    const int co = p.cout_real, ocs = co / 16;                 // 16-channel chunks per phase
    const int ph_per_tile = p.BN / co, npairs = ph_per_tile >> 1;
    for (int a = 0; a < p.NACC; a++) {
        const int q = m0 + a * 128 + quarter * 32 + lane;
        const int h = (int)__umulhi((uint32_t)q, p.pw_magic), ws = q - h * p.PW;
        const int w = p.band_tw ? band * p.band_tw + ws - 2 : ws;
        const bool ok = q < p.Lp && (p.band_tw ? (ws >= 2 && ws < p.band_tw + 2 && w < p.Wimg) : ws < p.W);
        for (int pc = part; pc < npairs * ocs; pc += step) {
            const int lp = 2 * (pc / ocs), ob = pc - (pc / ocs) * ocs;          // local phase of the pair's first member, channel chunk
            const int col0 = lp * co + ob * 16, col1 = col0 + co;
            uint32_t r0[16], r1[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * p.BN + col0), r0);
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * p.BN + col1), r1);
            if (!ok) continue;
            const int py = (jn * ph_per_tile + lp) >> 1;
            const int oy = 2 * h + py, ox = 2 * w;
            float nz0 = 0.f, nz1 = 0.f;
            if (p.noise) {
                const float2 t = __ldg(reinterpret_cast<const float2*>(p.noise + (size_t)n * p.noise_bstride + (size_t)oy * W2 + ox));
                nz0 = t.x * p.gain; nz1 = t.y * p.gain;
            }
            float v0[16], v1[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                v0[i] = fmaf(__uint_as_float(r0[i]), s_scale[col0 + i], s_shift[col0 + i] + nz0);
                v1[i] = fmaf(__uint_as_float(r1[i]), s_scale[col1 + i], s_shift[col1 + i] + nz1);
            }
            if (do_act) {
#pragma unroll
                for (int i = 0; i < 16; i++) { v0[i] = fmaxf(v0[i], 0.f) + slope * fminf(v0[i], 0.f); v1[i] = fmaxf(v1[i], 0.f) + slope * fminf(v1[i], 0.f); }
            }
            if (do_clamp) {
#pragma unroll
                for (int i = 0; i < 16; i++) { v0[i] = fminf(fmaxf(v0[i], -cl), cl); v1[i] = fminf(fmaxf(v1[i], -cl), cl); }
            }
            const int o0 = ob * 16;
            if (p.y_c8) {
                uint4* yb = reinterpret_cast<uint4*>(p.y) + ((size_t)n * p.cb_out + (size_t)o0 / 8) * H2W2 + (size_t)oy * W2 + ox;
                yb[0] = pack_half8(v0); yb[1] = pack_half8(v1);
                yb[H2W2] = pack_half8(v0 + 8); yb[H2W2 + 1] = pack_half8(v1 + 8);
            } else if (p.out_half) {
                __half2* yh = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p.y) + ((size_t)n * co + o0) * H2W2 + (size_t)oy * W2 + ox);
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    uint32_t hh;
                    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(v1[i]), "f"(v0[i]));
                    *reinterpret_cast<uint32_t*>(yh + (size_t)i * (H2W2 / 2)) = hh;
                }
            } else {
                float2* yp = reinterpret_cast<float2*>(p.y + ((size_t)n * co + o0) * H2W2 + (size_t)oy * W2 + ox);
#pragma unroll
                for (int i = 0; i < 16; i++) yp[(size_t)i * (H2W2 / 2)] = make_float2(v0[i], v1[i]);
            }
        }
    }
}

// per-sample input scale: style * in_gain (1 * in_gain for plain convs); zero for padded channels
__device__ __forceinline__ void stage_styles(const ConvParams& p, const int n, float* s_style, const int cin_pad, const int tid, const int nthreads) {
    for (int c = tid; c < cin_pad; c += nthreads)
        s_style[c] = (c < p.Cin) ? (p.styles ? (p.down2 ? p.styles[(size_t)n * p.cin_real + ((c % (2 * p.cin_real)) >> 1)] : p.im2col ? p.styles[(size_t)n * p.cin_real + c / (p.im2col * p.im2col)] : p.styles[(size_t)n * p.Cin + c]) : 1.f) * p.in_gain : 0.f;
}

// epilogue constants with the output gain folded in (relu / lrelu / linear are positively homogeneous, gain > 0)
__device__ __forceinline__ void stage_epilogue_constants(const ConvParams& p, const int n, const int jn, float* s_scale, float* s_shift, const int tid, const int nthreads) {
    for (int j = tid; j < p.BN; j += nthreads) {
        const int v = jn * p.BN + j;
        const int o = p.up2 ? v % p.cout_real : v;
        const bool live = v < p.Cout;
        if (p.spade) {
            const int C = p.cout_real >> 1, Ct = p.BN >> 1, ch = jn * Ct + j;
            const float r = j < Ct ? p.sp_rstd[(size_t)n * C + ch] : 0.f;
            s_scale[j] = r;
            s_shift[j] = j < Ct ? -p.sp_mean[(size_t)n * C + ch] * r : 0.f;
            continue;
        }
        s_scale[j] = live ? (p.dcoefs ? p.dcoefs[(size_t)n * p.cout_real + o] : 1.f) * p.gain : 0.f;
        s_shift[j] = live && p.bias ? p.bias[o] * p.gain : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------- main kernel
// SCALE: the A operand needs a per-channel scale and/or an input activation (modulated / SPADE layers); plain layers skip both.
template <bool SCALE>
__global__ void __launch_bounds__(kConvThreads, 2) conv_igemm_kernel(const __grid_constant__ ConvParams p, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int tile = blockIdx.x % p.tiles_per_img;
    const int nb_  = blockIdx.x / p.tiles_per_img;
    const int band = p.band_tw ? nb_ % p.nbands : 0;
    const int n    = p.band_tw ? nb_ / p.nbands : nb_;
    const int jn   = blockIdx.y;
    const int BM   = 128 * p.NACC;
    const int m0   = tile * BM;
    const int HW   = p.H * p.Wimg;
    if (threadIdx.x == 0 && PG_DBG(p)) {
        PG_TS(0);
        uint32_t smid; unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        PG_PUT(11, (long long)smid); PG_PUT(12, (long long)gt);
    }

    uint8_t* a_base = smem;
    uint8_t* b_base = a_base + (size_t)p.SA * p.a_stage_bytes;
    float*   s_style = reinterpret_cast<float*>(b_base + (size_t)p.SB * p.b_slot_bytes);
    const int cin_pad = p.nchunks * kKC;
    float* s_scale = s_style + cin_pad;                     // per output column: dcoef * gain
    float* s_shift = s_scale + p.BN;                        //                    bias * gain
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + p.BN);
    uint64_t* a_full = bars, *a_empty = bars + p.SA, *b_full = bars + 2 * p.SA, *b_empty = bars + 2 * p.SA + p.SB;
    uint64_t* acc_full = bars + 2 * p.SA + 2 * p.SB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    int2* s_tab = reinterpret_cast<int2*>(acc_full + 2);    // folded-tap mode: per virtual channel (element offset, (dh << 16) | (dw & 0xffff))

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.SA; i++) { mbar_init(smem_u32(&a_full[i]), p.tma_a ? 1u : (uint32_t)(kConvWarps / p.cgroups)); mbar_init(smem_u32(&a_empty[i]), 1); }
        for (int i = 0; i < p.SB; i++) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (p.tma_a) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        if (p.tma_a && p.tma_cb2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    stage_styles(p, n, s_style, cin_pad, threadIdx.x, kConvThreads);
    if (p.im2col) {
        const int ks = p.im2col, kk = ks * ks, pad = ks >> 1;
        for (int c = threadIdx.x; c < cin_pad; c += kConvThreads) {
            const int cr = c / kk, r2 = c - cr * kk, kh = r2 / ks, kw = r2 - kh * ks;
            s_tab[c] = c < p.Cin ? make_int2(cr * HW + (kh - pad) * p.W + (kw - pad), ((kh - pad) << 16) | ((kw - pad) & 0xffff))
                                 : make_int2(0, 0x7fff0000);                                    // padding channel: row offset out of range
        }
    }
    stage_epilogue_constants(p, n, jn, s_scale, s_shift, threadIdx.x, kConvThreads);
    if (p.vec2 && !p.tma_a) {     // the vec2 loader never writes the padding slots of the strip: zero every A stage once
        uint4* z = reinterpret_cast<uint4*>(a_base);
        const int nz = (int)((size_t)p.SA * p.a_stage_bytes / 16);
        for (int i = threadIdx.x; i < nz; i += kConvThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) PG_TS(1);

    // TMA A operand: the staged box starts at image row r0 = floor(q0 / PW) (q0 = first staged strip position of this tile, may be negative);
    // the tile's first GEMM row sits a_off rows into the stage
    const int tma_halo = (p.ks == 3) ? p.PW + 1 : 0;
    const int tma_q0 = m0 - tma_halo;
    const int tma_r0 = tma_q0 >= 0 ? tma_q0 / p.PW : -((-tma_q0 + p.PW - 1) / p.PW);
    const uint32_t a_off16 = p.tma_a ? (uint32_t)(tma_q0 - tma_r0 * p.PW) : 0u;
    if (warp == 0) {
        // ===================== B producer: a ring of small bulk copies (`tps` taps each) keeps many copies in flight =====================
        // (TMA mode: the same thread also issues the A boxes, one per 16-channel chunk, in chunk order ahead of that chunk's weight slots)
        if (lane == 0) {
            const uint8_t* src = (const uint8_t*)p.wpack + (size_t)n * p.wpack_sample_stride + (size_t)jn * p.nchunks * p.ntaps * p.b_tile_bytes;
            const int nslots = p.nchunks * (p.ntaps / p.tps);
            const int spc = p.ntaps / p.tps;                   // weight slots per chunk
            const int col0 = p.band_tw ? 2 * (band * p.band_tw - 2) : 0;       // box start along the row, in 8-byte units (two per position)
            int st = 0; uint32_t ph = 0;
            int sa = 0; uint32_t pa = 0;
            for (int g = 0; g < nslots; g++) {
                if (p.tma_a && g % spc == 0) {
                    mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1);
                    mbar_expect_tx(smem_u32(&a_full[sa]), p.a_tx_bytes);
                    const int cb = 2 * (g / spc);                // first channel block of this chunk; blocks >= tma_cb come from the second input (fused concat)
                    if (p.tma_down2) {
                        const int cpp = p.tma_cb >> 1, q = (g / spc) / cpp, cbk = 2 * ((g / spc) - q * cpp);       // chunks per parity, parity (a, b), first block
                        tma_load_4d(smem_u32(a_base + (size_t)sa * p.a_stage_bytes), &tmap_a, 0, col0 + (q & 1), 2 * tma_r0 + (q >> 1), n * p.tma_cb + cbk, smem_u32(&a_full[sa]));
                    } else
                    if (cb < p.tma_cb) tma_load_3d(smem_u32(a_base + (size_t)sa * p.a_stage_bytes), &tmap_a, col0, tma_r0, n * p.tma_cb + cb, smem_u32(&a_full[sa]));
                    else               tma_load_3d(smem_u32(a_base + (size_t)sa * p.a_stage_bytes), &tmap_a2, col0, tma_r0, n * p.tma_cb2 + cb - p.tma_cb, smem_u32(&a_full[sa]));
                    if (++sa == p.SA) { sa = 0; pa ^= 1; }
                }
                mbar_wait(smem_u32(&b_empty[st]), ph ^ 1);
                if (PG_DBGMODE(p) & 4) { mbar_arrive(smem_u32(&b_full[st])); if (++st == p.SB) { st = 0; ph ^= 1; } continue; }
                mbar_expect_tx(smem_u32(&b_full[st]), p.b_slot_bytes);
                bulk_g2s(smem_u32(b_base + (size_t)st * p.b_slot_bytes), src + (size_t)g * p.b_slot_bytes, p.b_slot_bytes, smem_u32(&b_full[st]));
                if (++st == p.SB) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The issue loop must cost less than one MMA (64 cycles at N = 128): descriptors are a constant high word plus a
        // 14-bit start-address field, so each MMA is two integer adds and the tcgen05.mma itself; taps are fully unrolled.
        const uint32_t issue = elect_one();
        const uint32_t ab = smem_u32(a_base), bb = smem_u32(b_base), af = smem_u32(a_full), ae = smem_u32(a_empty), bf = smem_u32(b_full),
                       be = smem_u32(b_empty), accf = smem_u32(acc_full);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        MmaRing ring = {0, 0, 0u, 0u};
#define PG_ISSUE(KS_, NACC_) mma_issue_loop<KS_, NACC_>(p, ab, bb, af, ae, bf, be, accf, tb, issue, ring, a_off16)
        if (p.ks == 3) { if (p.NACC == 4) PG_ISSUE(3, 4); else if (p.NACC == 3) PG_ISSUE(3, 3); else if (p.NACC == 2) PG_ISSUE(3, 2); else PG_ISSUE(3, 1); }
        else           { if (p.NACC == 4) PG_ISSUE(1, 4); else if (p.NACC == 3) PG_ISSUE(1, 3); else if (p.NACC == 2) PG_ISSUE(1, 2); else PG_ISSUE(1, 1); }
#undef PG_ISSUE
    } else {
        // ===================== A converters (none in TMA mode: these warps are the epilogue only) =====================
        // Each warp owns the tasks t = cw, cw + 8, ... of every chunk and walks them as ONE stream that runs across chunk boundaries, in register
        // batches (stream_tasks): the global loads of a batch never wait for the shared-memory stage, only the stores wait on a_empty.
        //   * vec2 loader (W even, 8-byte aligned planes, not down-2): a task is 64 consecutive floats (one aligned 256-byte run) of 8 channels of
        //     the flat NCHW plane, read with LDG.64: 2 L1 wavefronts per 256 B where the scalar loader (32 strip positions per task, unaligned
        //     because the strip pitch is W + 1) needs 2 per 128 B.  Element e of row h lands on staged strip slot e + h * (PW - W) - q0; the
        //     padding slots (zero column, rows outside the image) are never written and stay zero from the one-time fill in the prologue.
        //   * scalar loader: any W, down-2 space-to-depth reads; writes every staged slot (zeros where out of range).
        const int cw = warp - 2;
        const int halo = (p.ks == 3) ? p.PW + 1 : 0;
        const int n_in = (PG_DBGMODE(p) & 8) ? 0 : n;      // tuning: every sample reads sample 0 (input stays L2-resident)
        const float* xn = p.down2 ? p.x + (size_t)n_in * p.cin_real * p.hin * p.win
                        : p.in_half ? reinterpret_cast<const float*>(reinterpret_cast<const __half*>(p.x) + (size_t)n_in * p.cin1 * HW)
                                    : p.x + (size_t)n_in * p.cin1 * HW;
        const float* xn2 = p.x2 ? p.x2 + (size_t)n * (p.Cin - p.cin1) * HW - (size_t)p.cin1 * HW : xn;   // indexed with the global channel number
        auto chan_base = [&](int c0) { return (c0 < p.cin1 ? xn : xn2) + (size_t)c0 * HW; };
        const bool has_in_act = p.in_act != PG_ACT_LINEAR;
        const float in_slope = (p.in_act == PG_ACT_RELU) ? 0.f : p.in_alpha;
        const int nchunks = p.nchunks;
        const int q0 = m0 - halo;                              // strip position of staged slot 0
        int s_st = 0; uint32_t s_ph = 0;                       // store cursor: stage / parity of the chunk being written
        long long wait_e = 0, t_issue = 0, t_store = 0;
        auto stage_begin = [&]() {
            const long long t0 = PG_DBG(p) ? clock64() : 0;
            mbar_wait(smem_u32(&a_empty[s_st]), s_ph ^ 1);
            if (PG_DBG(p)) wait_e += clock64() - t0;
        };
        auto stage_end = [&]() {
            fence_proxy_async();                               // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&a_full[s_st]));
            if (++s_st == p.SA) { s_st = 0; s_ph ^= 1; }
        };
        auto scale8 = [&](float (&v)[8], int ci, int plane) {
            const float* sc = s_style + ci * kKC + plane * 8;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float a = v[i];
                if (has_in_act) a = fmaxf(a, 0.f) + in_slope * fminf(a, 0.f);
                v[i] = a * sc[i];
            }
        };
        auto pack8 = [&](const float (&v)[8]) {
            uint4 pk;
            pk.x = pack2(v[0], v[1], p.fmt); pk.y = pack2(v[2], v[3], p.fmt);
            pk.z = pack2(v[4], v[5], p.fmt); pk.w = pack2(v[6], v[7], p.fmt);
            return pk;
        };

        if (p.tma_a) {
            // nothing to stage: the A boxes arrive by TMA (warp 0)
        } else if (p.vec2) {
            // flat element range [e_lo, e_hi) of the plane that the staged strip slots [q0, q0 + PA) cover
            const int qa = q0 < 0 ? 0 : q0, qb = (q0 + p.PA < p.Lp ? q0 + p.PA : p.Lp) - 1;     // first / last strip position inside the image
            int ha = (int)__umulhi((uint32_t)qa, p.pw_magic), wa = qa - ha * p.PW;
            int hb = (int)__umulhi((uint32_t)qb, p.pw_magic), wb = qb - hb * p.PW;
            const int e_lo = wa >= p.W ? (ha + 1) * p.W : ha * p.W + wa;
            const int e_hi = wb >= p.W ? (hb + 1) * p.W : hb * p.W + wb + 1;
            const int g_lo = e_lo >> 1, g_hi = (e_hi + 1) >> 1;                 // pairs of floats
            const int nseg = g_hi > g_lo ? (g_hi - g_lo + 31) >> 5 : 0;
            const int ntasks = nseg * 2;                                        // (segment of 32 pairs, plane)
            const int wpg_ = p.lean || p.in_half ? kConvWarps / p.cgroups : kConvWarps;
            const int tpw = ntasks ? (ntasks + wpg_ - 1) / wpg_ : 1;
            const int dpitch = p.PW - p.W;
            auto load_task = [&](float (&v)[16], int ci, int idx) {
                const int tt = cw + idx * kConvWarps;
                const int g = g_lo + (tt >> 1) * 32 + lane;
                const bool ok = ci < nchunks && tt < ntasks && g < g_hi && !(PG_DBGMODE(p) & 1);
                const int c0 = ci * kKC + (tt & 1) * 8;
                const float* src = chan_base(c0) + 2 * g;
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = 0.f;
                if (ok) {
                    if (c0 + 8 <= p.Cin) {
#pragma unroll
                        for (int i = 0; i < 8; i++) { const float2 t = __ldg(reinterpret_cast<const float2*>(src + (size_t)i * HW)); v[i] = t.x; v[8 + i] = t.y; }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; i++) if (c0 + i < p.Cin) { const float2 t = __ldg(reinterpret_cast<const float2*>(src + (size_t)i * HW)); v[i] = t.x; v[8 + i] = t.y; }
                    }
                }
            };
            auto store_task = [&](float (&v)[16], int ci, int idx) {
                if (ci >= nchunks) return;
                if (idx == 0) stage_begin();
                const int tt = cw + idx * kConvWarps;
                const int g = g_lo + (tt >> 1) * 32 + lane;
                if (tt < ntasks && g < g_hi && !(PG_DBGMODE(p) & 2)) {
                    const int plane = tt & 1;
                    const int e = 2 * g;
                    const int h = (int)__umulhi((uint32_t)e, p.w_magic);
                    const int s0 = e + h * dpitch - q0;                          // staged slot of the first element; the second is s0 + 1 (same row)
                    float lo[8], hi[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) { lo[i] = v[i]; hi[i] = v[8 + i]; }
                    if (SCALE) { scale8(lo, ci, plane); scale8(hi, ci, plane); }
                    const uint4 pl = pack8(lo), ph = pack8(hi);
                    // lanes sit 32 B apart in the stage: quarter-warps store conflict-free when lanes 4..7 of each group of 8 write their second slot first
                    const bool swap = (lane >> 2) & 1;
                    uint8_t* base = a_base + (size_t)s_st * p.a_stage_bytes + (size_t)plane * p.PA * 16;
                    const int sa_ = swap ? s0 + 1 : s0, sb_ = swap ? s0 : s0 + 1;
                    const uint4 pa_ = swap ? ph : pl, pb_ = swap ? pl : ph;
                    if (sa_ >= 0 && sa_ < p.PA) *reinterpret_cast<uint4*>(base + (size_t)sa_ * 16) = pa_;
                    if (sb_ >= 0 && sb_ < p.PA) *reinterpret_cast<uint4*>(base + (size_t)sb_ * 16) = pb_;
                }
                if (idx == tpw - 1) stage_end();
            };
#define PG_CONVERT(PER_, ROUNDS_) convert_vec2<SCALE, false, PER_, ROUNDS_>(p, xn, xn2, HW, cw, lane, g_lo, g_hi, ntasks, q0, a_base, s_style, a_full, a_empty, wait_e, kConvWarps, 0, false, band)
#define PG_CONVERT_H(PER_, ROUNDS_) convert_vec2<false, true, PER_, ROUNDS_>(p, xn, xn2, HW, cw, lane, g_lo, g_hi, ntasks, q0, a_base, s_style, a_full, a_empty, wait_e, kConvWarps, 0, false, band)
            if (p.in_half) {                              // fp16 input (host side guarantees tpw <= 6, no input scale / activation)
                if (tpw <= 2) PG_CONVERT_H(2, 1); else if (tpw <= 4) PG_CONVERT_H(2, 2); else PG_CONVERT_H(3, 2);
            } else if (p.lean && tpw <= 6) {
                if (tpw <= 1) PG_CONVERT(1, 1); else if (tpw == 2) PG_CONVERT(2, 1); else if (tpw == 3) PG_CONVERT(3, 1);
                else if (tpw == 4) PG_CONVERT(2, 2); else PG_CONVERT(3, 2);
            } else {
                stream_tasks<16, 2>(load_task, store_task, tpw, nchunks, PG_PIPE(p) != 0, PG_DBG(p), t_issue, t_store);
            }
#undef PG_CONVERT
#undef PG_CONVERT_H
        } else {
            const int ntasks = (p.PA / 32) * 2;                       // (group of 32 strip positions, plane)
            const int tpw = (ntasks + kConvWarps - 1) / kConvWarps;   // stream slots per warp per chunk (trailing ones may be void)
            auto load_task = [&](float (&v)[8], int ci, int idx) {
                const int tt = cw + idx * kConvWarps;
                const int q = q0 + (tt >> 1) * 32 + lane;                 // strip position of this staged row
                int h = 0, w = 0;
                bool ok = ci < nchunks && tt < ntasks && q >= 0 && q < p.Lp && !(PG_DBGMODE(p) & 1);
                if (ok) { h = (int)__umulhi((uint32_t)q, p.pw_magic); w = q - h * p.PW; ok = w < p.W; }
                if (p.band_tw) { w = band * p.band_tw - 2 + w; ok = ok && w >= 0 && w < p.Wimg; }     // band mode (down-2 layers): image column of this slot
                const int c0 = ci * kKC + (tt & 1) * 8;
#pragma unroll
                for (int i = 0; i < 8; i++) v[i] = 0.f;
                if (p.im2col) {
                    if (ok) {
                        const int2* tab = s_tab + c0;
                        const float* src = xn + h * p.W + w;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const int2 t = tab[i];
                            const int hh = h + (t.y >> 16), ww = w + (int)(short)(t.y & 0xffff);
                            if ((unsigned)hh < (unsigned)p.H && (unsigned)ww < (unsigned)p.W) v[i] = __ldg(src + t.x);
                        }
                    }
                } else if (!p.down2) {
                    const float* src = chan_base(c0) + h * p.W + w;
                    if (ok) {
                        if (c0 + 8 <= p.Cin) {
#pragma unroll
                            for (int i = 0; i < 8; i++) v[i] = __ldg(src + (size_t)i * HW);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; i++) if (c0 + i < p.Cin) v[i] = __ldg(src + (size_t)i * HW);
                        }
                    }
                } else if (ok) {
                    // down-2: the 8 virtual channels of this group are 4 real channels x both column parities of input row 2h + a: four aligned
                    // 8-byte reads per lane, consecutive lanes read consecutive pairs (fully used sectors)
                    const int a = c0 / (2 * p.cin_real), cr = (c0 - a * 2 * p.cin_real) >> 1;
                    const size_t cs = (size_t)p.hin * p.win;
                    const float* src = xn + (size_t)cr * cs + (size_t)(2 * h + a) * p.win + 2 * w;
#pragma unroll
                    for (int i = 0; i < 4; i++) { const float2 t = __ldg(reinterpret_cast<const float2*>(src + (size_t)i * cs)); v[2 * i] = t.x; v[2 * i + 1] = t.y; }
                }
            };
            auto store_task = [&](float (&v)[8], int ci, int idx) {
                if (ci >= nchunks) return;
                if (idx == 0) stage_begin();
                const int tt = cw + idx * kConvWarps;
                if (tt < ntasks && !(PG_DBGMODE(p) & 2)) {
                    const int plane = tt & 1, spos = (tt >> 1) * 32 + lane;
                    if (SCALE) scale8(v, ci, plane);
                    if (p.fmt == 2) {       // tf32: the 8 channels of this task are two 4-channel planes
                        uint8_t* dst = a_base + (size_t)s_st * p.a_stage_bytes + (size_t)(2 * plane) * p.PA * 16 + (size_t)spos * 16;
                        *reinterpret_cast<uint4*>(dst) = make_uint4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
                        *reinterpret_cast<uint4*>(dst + (size_t)p.PA * 16) = make_uint4(to_tf32(v[4]), to_tf32(v[5]), to_tf32(v[6]), to_tf32(v[7]));
                    } else
                    *reinterpret_cast<uint4*>(a_base + (size_t)s_st * p.a_stage_bytes + (size_t)plane * p.PA * 16 + (size_t)spos * 16) = pack8(v);
                }
                if (idx == tpw - 1) stage_end();
            };
#define PG_CONVERT_D(PER_, ROUNDS_) convert_down2<SCALE, PER_, ROUNDS_>(p, xn, cw, lane, ntasks, q0, band, a_base, s_style, a_full, a_empty, wait_e)
            if (p.down2 && p.lean && tpw <= 6) {
                if (tpw <= 2) PG_CONVERT_D(2, 1); else if (tpw <= 4) PG_CONVERT_D(4, 1); else PG_CONVERT_D(3, 2);
            } else {
                stream_tasks<8, 4>(load_task, store_task, tpw, nchunks, PG_PIPE(p) != 0, PG_DBG(p), t_issue, t_store);
            }
#undef PG_CONVERT_D
        }
        if (cw == 0 && lane == 0) { PG_TS(6); PG_PUT(10, wait_e); PG_PUT(14, t_issue); PG_PUT(15, t_store); }
        // ===================== epilogue (same warps) =====================
        prefetch_residual_tile(p, n, jn, m0, HW, warp & 3, cw >> 2, kConvWarps / 4, lane, band);
        mbar_wait(smem_u32(acc_full), 0);
        tc_fence_after();
        if (cw == 0 && lane == 0) PG_TS(4);
        if (p.up2_pair) epilogue_tile_up2_pair(p, n, jn, m0, HW, tmem_base, s_scale, s_shift, warp & 3, cw >> 2, kConvWarps / 4, lane, band);
        else epilogue_tile(p, n, jn, m0, HW, tmem_base, s_scale, s_shift, warp & 3, cw >> 2, kConvWarps / 4, lane, band);   // 2 warps per TMEM lane quarter
    }
    if (threadIdx.x == 64 && PG_DBG(p)) {
        PG_TS(5);
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        PG_PUT(13, (long long)gt);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- persistent kernel
// One CTA per SM walks a static list of output tiles (tile = blockIdx.x, + gridDim.x, ...).  Roles: warp 0 streams the weights, warp 1 issues the
// MMAs, 12 converter warps (two groups on alternate chunks) stage the A operand, 4 epilogue warps drain TMEM.  TMEM holds TWO accumulator
// buffers of NACC x BN <= 256 columns: while the epilogue warps read buffer i & 1 the MMA warp is already filling the other one, and the
// converters / weight stream never stop at a tile boundary (their stage rings run on), so the prologue, pipeline fill and epilogue that cost a
// quarter of a one-tile CTA's life are off the critical path.  Covers the aligned pair loader (plain / modulated / split input / residual /
// fp16 in / fp16 out), 1x1 and 3x3, stride 1; everything else keeps conv_igemm_kernel.
constexpr int kPConvWarps = 12;
constexpr int kPEpiWarps  = 4;
constexpr int kPThreads   = 64 + 32 * (kPConvWarps + kPEpiWarps);

template <bool SCALE>
__global__ void __launch_bounds__(kPThreads, 1) conv_igemm_persistent_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int BM = 128 * p.NACC;
    const int HW = p.H * p.Wimg;
    const int tiles_n = p.N * p.tiles_per_img;               // tiles of one n-tile (jn)
    const int total = tiles_n * p.ntiles_n;
    const int cin_pad = p.nchunks * kKC;

    uint8_t* a_base = smem;
    uint8_t* b_base = a_base + (size_t)p.SA * p.a_stage_bytes;
    float*   s_style = reinterpret_cast<float*>(b_base + (size_t)p.SB * p.b_slot_bytes);      // [2][cin_pad]
    float*   s_scale = s_style + 2 * cin_pad;                                                  // [2][BN]
    float*   s_shift = s_scale + 2 * p.BN;                                                     // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 2 * p.BN);
    uint64_t* a_full = bars, *a_empty = bars + p.SA, *b_full = bars + 2 * p.SA, *b_empty = bars + 2 * p.SA + p.SB;
    uint64_t* acc_full = bars + 2 * p.SA + 2 * p.SB, *acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.SA; i++) { mbar_init(smem_u32(&a_full[i]), (uint32_t)(kPConvWarps / p.cgroups)); mbar_init(smem_u32(&a_empty[i]), 1); }
        for (int i = 0; i < p.SB; i++) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), kPEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // padding slots are re-zeroed per tile by the converters; start from a clean slate anyway
        uint4* z = reinterpret_cast<uint4*>(a_base);
        const int nz = (int)((size_t)p.SA * p.a_stage_bytes / 16);
        for (int i = threadIdx.x; i < nz; i += kPThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== B producer =====================
        if (lane == 0) {
            const int nslots = p.nchunks * (p.ntaps / p.tps);
            int st = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int jn = t / tiles_n, n_t = (t - jn * tiles_n) / p.tiles_per_img;
                const uint8_t* src = (const uint8_t*)p.wpack + (size_t)n_t * p.wpack_sample_stride + (size_t)jn * p.nchunks * p.ntaps * p.b_tile_bytes;
                for (int g = 0; g < nslots; g++) {
                    mbar_wait(smem_u32(&b_empty[st]), ph ^ 1);
                    mbar_expect_tx(smem_u32(&b_full[st]), p.b_slot_bytes);
                    bulk_g2s(smem_u32(b_base + (size_t)st * p.b_slot_bytes), src + (size_t)g * p.b_slot_bytes, p.b_slot_bytes, smem_u32(&b_full[st]));
                    if (++st == p.SB) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t issue = elect_one();
        const uint32_t ab = smem_u32(a_base), bb = smem_u32(b_base), af = smem_u32(a_full), ae = smem_u32(a_empty), bf = smem_u32(b_full),
                       be = smem_u32(b_empty);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        MmaRing ring = {0, 0, 0u, 0u};
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, it++) {
            const int buf = it & 1;
            mbar_wait(smem_u32(&acc_empty[buf]), (((uint32_t)it >> 1) & 1u) ^ 1u);       // the epilogue has drained this buffer (tile it - 2)
            tc_fence_after();
            const uint32_t accf = smem_u32(&acc_full[buf]), tacc = tb + (uint32_t)buf * 256u;
#define PG_ISSUE(KS_, NACC_) mma_issue_loop<KS_, NACC_>(p, ab, bb, af, ae, bf, be, accf, tacc, issue, ring)
            if (p.ks == 3) { if (p.NACC == 4) PG_ISSUE(3, 4); else if (p.NACC == 3) PG_ISSUE(3, 3); else if (p.NACC == 2) PG_ISSUE(3, 2); else PG_ISSUE(3, 1); }
            else           { if (p.NACC == 4) PG_ISSUE(1, 4); else if (p.NACC == 3) PG_ISSUE(1, 3); else if (p.NACC == 2) PG_ISSUE(1, 2); else PG_ISSUE(1, 1); }
#undef PG_ISSUE
        }
    } else if (warp < 2 + kPConvWarps) {
        // ===================== A converters =====================
        const int cw = warp - 2;
        const int halo = (p.ks == 3) ? p.PW + 1 : 0;
        long long wait_e = 0;
        int it = 0, chunk_base = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, it++, chunk_base += p.nchunks) {
            const int r = t % tiles_n, n = r / p.tiles_per_img, tile = r - n * p.tiles_per_img;
            const int m0 = tile * BM, q0 = m0 - halo;
            float* sty = s_style + (it & 1) * cin_pad;
            if (SCALE) {
                // this tile's per-sample input scales; buffer it & 1 was last read for tile it - 2, which every converter warp finished before it
                // passed the named barrier of tile it - 1
                stage_styles(p, n, sty, cin_pad, (int)threadIdx.x - 64, 32 * kPConvWarps);
                asm volatile("bar.sync 1, %0;" ::"r"(32 * kPConvWarps) : "memory");
            }
            const float* xn = p.in_half ? reinterpret_cast<const float*>(reinterpret_cast<const __half*>(p.x) + (size_t)n * p.cin1 * HW) : p.x + (size_t)n * p.cin1 * HW;
            const float* xn2 = p.x2 ? p.x2 + (size_t)n * (p.Cin - p.cin1) * HW - (size_t)p.cin1 * HW : xn;
            const int qa = q0 < 0 ? 0 : q0, qb = (q0 + p.PA < p.Lp ? q0 + p.PA : p.Lp) - 1;
            const int ha = (int)__umulhi((uint32_t)qa, p.pw_magic), wa = qa - ha * p.PW;
            const int hb = (int)__umulhi((uint32_t)qb, p.pw_magic), wb = qb - hb * p.PW;
            const int e_lo = wa >= p.W ? (ha + 1) * p.W : ha * p.W + wa;
            const int e_hi = wb >= p.W ? (hb + 1) * p.W : hb * p.W + wb + 1;
            const int g_lo = e_lo >> 1, g_hi = (e_hi + 1) >> 1;
            const int nseg = g_hi > g_lo ? (g_hi - g_lo + 31) >> 5 : 0;
            const int ntasks = nseg * 2;
            const int wpg = kPConvWarps / p.cgroups;
            const int tpw = ntasks ? (ntasks + wpg - 1) / wpg : 1;
#define PG_CONVERT(PER_, ROUNDS_) convert_vec2<SCALE, false, PER_, ROUNDS_>(p, xn, xn2, HW, cw, lane, g_lo, g_hi, ntasks, q0, a_base, sty, a_full, a_empty, wait_e, kPConvWarps, chunk_base, true)
#define PG_CONVERT_H(PER_, ROUNDS_) convert_vec2<false, true, PER_, ROUNDS_>(p, xn, xn2, HW, cw, lane, g_lo, g_hi, ntasks, q0, a_base, sty, a_full, a_empty, wait_e, kPConvWarps, chunk_base, true)
            if (p.in_half) { if (tpw <= 2) PG_CONVERT_H(2, 1); else if (tpw <= 4) PG_CONVERT_H(2, 2); else PG_CONVERT_H(3, 2); }
            else if (tpw <= 1) PG_CONVERT(1, 1); else if (tpw == 2) PG_CONVERT(2, 1); else if (tpw == 3) PG_CONVERT(3, 1);
            else if (tpw == 4) PG_CONVERT(2, 2); else PG_CONVERT(3, 2);                   // host side guarantees tpw <= 6
#undef PG_CONVERT
#undef PG_CONVERT_H
        }
    } else {
        // ===================== epilogue warps =====================
        const int et = (int)threadIdx.x - 32 * (2 + kPConvWarps);     // 0..127
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, it++) {
            const int jn = t / tiles_n, r = t - jn * tiles_n, n = r / p.tiles_per_img, tile = r - n * p.tiles_per_img;
            const int buf = it & 1;
            float* sc = s_scale + buf * p.BN; float* sh = s_shift + buf * p.BN;
            stage_epilogue_constants(p, n, jn, sc, sh, et, 32 * kPEpiWarps);
            asm volatile("bar.sync 2, %0;" ::"r"(32 * kPEpiWarps) : "memory");
            mbar_wait(smem_u32(&acc_full[buf]), ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            epilogue_tile(p, n, jn, tile * BM, HW, tmem_base + (uint32_t)buf * 256u, sc, sh, warp & 3, 0, 1, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------- persistent TMA kernel
// Channel-blocked fp16 input (TMA operand path): no converter warps, so the whole CTA is three pipelines with nothing else in the way.
//   warp 0      producer: per 16-channel chunk one A box (cp.async.bulk.tensor) and the chunk's weight slots (cp.async.bulk), rings run on across tiles
//   warp 1      MMA issuer: tcgen05.mma into TMEM accumulator buffer (tile & 1); commits acc_full[buf] and goes straight on to the next tile
//   warps 2..9  epilogue: wait acc_full[buf], drain TMEM (2 warps per lane quarter; 16 epilogue warps were measured: no gain), arrive acc_empty[buf]
// One CTA per SM walks tiles t = blockIdx.x, + gridDim.x, ... of the list (jn, n, band, tile); TMEM holds two accumulator buffers of NACC * BN <= 256
// columns, so the epilogue of tile i overlaps the main loop of tile i + 1 and the per-CTA prologue (TMEM allocation, barrier setup, pipeline fill)
// is paid once per SM instead of once per tile.
constexpr int kTEpiWarps = 8;
constexpr int kTThreads  = 64 + 32 * kTEpiWarps;

__global__ void __launch_bounds__(kTThreads, 1) conv_igemm_tma_persistent_kernel(const __grid_constant__ ConvParams p, const __grid_constant__ CUtensorMap tmap_a,
                                                                                  const __grid_constant__ CUtensorMap tmap_a2) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int BM = 128 * p.NACC;
    const int HW = p.H * p.Wimg;
    const int tiles_s = p.nbands * p.tiles_per_img;          // tiles of one (n-tile, sample)
    const int tiles_n = p.N * tiles_s;                       // tiles of one n-tile
    const int total = tiles_n * p.ntiles_n;
    const int halo = (p.ks == 3) ? p.PW + 1 : 0;

    uint8_t* a_base = smem;
    uint8_t* b_base = a_base + (size_t)p.SA * p.a_stage_bytes;
    float*   s_scale = reinterpret_cast<float*>(b_base + (size_t)p.SB * p.b_slot_bytes);      // [2][BN]
    float*   s_shift = s_scale + 2 * p.BN;                                                     // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 2 * p.BN);
    uint64_t* a_full = bars, *a_empty = bars + p.SA, *b_full = bars + 2 * p.SA, *b_empty = bars + 2 * p.SA + p.SB;
    uint64_t* acc_full = bars + 2 * p.SA + 2 * p.SB, *acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.SA; i++) { mbar_init(smem_u32(&a_full[i]), 1); mbar_init(smem_u32(&a_empty[i]), 1); }
        for (int i = 0; i < p.SB; i++) { mbar_init(smem_u32(&b_full[i]), 1); mbar_init(smem_u32(&b_empty[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), kTEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        if (p.tma_cb2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile list decode: t -> (jn, n, band, tile); first staged strip position q0 = m0 - halo, box row r0 = floor(q0 / PW)
    auto decode = [&](int t, int& jn, int& n, int& band, int& m0) {
        jn = t / tiles_n;
        int r = t - jn * tiles_n;
        n = r / tiles_s; r -= n * tiles_s;
        band = r / p.tiles_per_img;
        m0 = (r - band * p.tiles_per_img) * BM;
    };

    if (warp == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            const int spc = p.ntaps / p.tps;                   // weight slots per chunk
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                int jn, n, band, m0;
                decode(t, jn, n, band, m0);
                const int q0 = m0 - halo;
                const int r0 = q0 >= 0 ? q0 / p.PW : -((-q0 + p.PW - 1) / p.PW);
                const int col0 = p.band_tw ? 2 * (band * p.band_tw - 2) : 0;
                const uint8_t* wsrc = (const uint8_t*)p.wpack + (size_t)n * p.wpack_sample_stride + (size_t)jn * p.nchunks * p.ntaps * p.b_tile_bytes;
                for (int ci = 0; ci < p.nchunks; ci++) {
                    mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1);
                    mbar_expect_tx(smem_u32(&a_full[sa]), p.a_tx_bytes);
                    const int cb = 2 * ci;
                    if (p.tma_down2) {
                        const int cpp = p.tma_cb >> 1, q = ci / cpp, cbk = 2 * (ci - q * cpp);
                        tma_load_4d(smem_u32(a_base + (size_t)sa * p.a_stage_bytes), &tmap_a, 0, col0 + (q & 1), 2 * r0 + (q >> 1), n * p.tma_cb + cbk, smem_u32(&a_full[sa]));
                    } else
                    if (cb < p.tma_cb) tma_load_3d(smem_u32(a_base + (size_t)sa * p.a_stage_bytes), &tmap_a, col0, r0, n * p.tma_cb + cb, smem_u32(&a_full[sa]));
                    else               tma_load_3d(smem_u32(a_base + (size_t)sa * p.a_stage_bytes), &tmap_a2, col0, r0, n * p.tma_cb2 + cb - p.tma_cb, smem_u32(&a_full[sa]));
                    if (++sa == p.SA) { sa = 0; pa ^= 1; }
                    for (int g = 0; g < spc; g++) {
                        mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1);
                        mbar_expect_tx(smem_u32(&b_full[sb]), p.b_slot_bytes);
                        bulk_g2s(smem_u32(b_base + (size_t)sb * p.b_slot_bytes), wsrc + (size_t)(ci * spc + g) * p.b_slot_bytes, p.b_slot_bytes, smem_u32(&b_full[sb]));
                        if (++sb == p.SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t issue = elect_one();
        const uint32_t ab = smem_u32(a_base), bb = smem_u32(b_base), af = smem_u32(a_full), ae = smem_u32(a_empty), bf = smem_u32(b_full),
                       be = smem_u32(b_empty);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        MmaRing ring = {0, 0, 0u, 0u};
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, it++) {
            int jn, n, band, m0;
            decode(t, jn, n, band, m0);
            const int q0 = m0 - halo;
            const int r0 = q0 >= 0 ? q0 / p.PW : -((-q0 + p.PW - 1) / p.PW);
            const uint32_t a_off16 = (uint32_t)(q0 - r0 * p.PW);
            const int buf = it & 1;
            mbar_wait(smem_u32(&acc_empty[buf]), (((uint32_t)it >> 1) & 1u) ^ 1u);       // the epilogue has drained this buffer (tile it - 2)
            tc_fence_after();
            const uint32_t accf = smem_u32(&acc_full[buf]), tacc = tb + (uint32_t)buf * 256u;
#define PG_ISSUE(KS_, NACC_) mma_issue_loop<KS_, NACC_>(p, ab, bb, af, ae, bf, be, accf, tacc, issue, ring, a_off16)
            if (p.ks == 3) { if (p.NACC == 2) PG_ISSUE(3, 2); else PG_ISSUE(3, 1); }
            else           { if (p.NACC == 2) PG_ISSUE(1, 2); else PG_ISSUE(1, 1); }
#undef PG_ISSUE
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 2;                                      // 0..7: TMEM lane quarter warp & 3, column part ew >> 2
        const int et = (int)threadIdx.x - 64;
        int it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, it++) {
            int jn, n, band, m0;
            decode(t, jn, n, band, m0);
            const int buf = it & 1;
            float* sc = s_scale + buf * p.BN; float* sh = s_shift + buf * p.BN;
            // buffer `buf` of the constants was last read for tile it - 2; every epilogue warp finished that tile before it arrived on acc_empty and
            // passed the named barrier of tile it - 1, so it is free to overwrite
            stage_epilogue_constants(p, n, jn, sc, sh, et, 32 * kTEpiWarps);
            asm volatile("bar.sync 2, %0;" ::"r"(32 * kTEpiWarps) : "memory");
            prefetch_residual_tile(p, n, jn, m0, HW, warp & 3, ew >> 2, kTEpiWarps / 4, lane, band);
            mbar_wait(smem_u32(&acc_full[buf]), ((uint32_t)it >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)buf * 256u;
            if (p.up2_pair) epilogue_tile_up2_pair(p, n, jn, m0, HW, tacc, sc, sh, warp & 3, ew >> 2, kTEpiWarps / 4, lane, band);
            else            epilogue_tile(p, n, jn, m0, HW, tacc, sc, sh, warp & 3, ew >> 2, kTEpiWarps / 4, lane, band);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- row-folded kernel (first layers)
// Small-Cin layers on wide images (the 3 -> 64 channel 7x7 / 3x3 first layers of the encoders, W % 128 == 0).  The folded-tap form above gathers
// Cin * k * k scalar values per output pixel through L1 (160 loads per pixel for 3 x 7 x 7: 465 us for one launch of the garment encoder).  Here only
// the kernel ROWS are folded into K: GEMM channel r = c * k + kh holds input row h + kh - pad of channel c, so a staged A row is
// "column w, the (c, kh) pairs" -- 21 coalesced row reads per column instead of 147 gathers -- and the k horizontal taps are k descriptor start
// addresses 16 bytes apart into the same staged tile, exactly like the taps of the main kernel.  K = 32 (4 planes of 8 pairs, zero padded), a tile
// is 128 consecutive pixels of one image row, k x (K / 16) MMAs of M = 128, N = Cout.  One accumulator, one A stage: several CTAs per SM overlap
// each other's build / MMA / epilogue phases; CTAs walk the tile list with stride gridDim.x, so concurrently processed tiles are vertical
// neighbours and the k-fold re-read of every input row is served by L2.
constexpr int kRfThreads = 256;
constexpr int kRfPlanes  = 4;       // K = 32 = 4 planes of 8 (channel, kernel row) pairs
constexpr int kRfPA      = 136;     // staged columns per tile: 128 outputs + k - 1 <= 6 halo columns, padded to a multiple of 8

struct RowFoldParams {
    const float* x; const void* wpack; const float* bias; void* y;
    int N, Cin, Cout, H, W, ks, BN, nrows, nchunks;       // nrows = Cin * ks real pairs, nchunks = K / 16 in use (1 or 2)
    int act; float alpha, gain, clamp;
    int y_c8, cb_out;
    uint32_t idesc, tmem_cols;
    int tiles_total, wsegs;
};

__global__ void conv_rowfold_prepack_kernel(PackParams p, long long out_sample_stride_halves) {
    // w [Cout, Cin, ks, ks] -> [tap kw][plane (4)][BN rows][8] fp16; GEMM channel r = plane * 8 + e = c * ks + kh
    const size_t total = (size_t)p.ks * kRfPlanes * p.BN * 8;
    const int sample = blockIdx.y;
    const float* w = p.w + (size_t)sample * p.w_bstride;
    __half* out = (__half*)p.out + (size_t)sample * out_sample_stride_halves;
    const float* sty = p.styles ? p.styles + (size_t)sample * p.Cin : nullptr;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        size_t q = idx;
        const int e = q % 8; q /= 8;
        const int o = q % p.BN; q /= p.BN;
        const int j = q % kRfPlanes; q /= kRfPlanes;
        const int kw = (int)q;
        const int r = j * 8 + e;
        float val = 0.f;
        if (o < p.Cout && r < p.Cin * p.ks) {
            const int c = r / p.ks, kh = r - c * p.ks;
            const int wh = p.flip_weight ? kh : p.ks - 1 - kh, ww = p.flip_weight ? kw : p.ks - 1 - kw;
            val = w[((size_t)(o * p.Cin + c) * p.ks + wh) * p.ks + ww] * p.w_scale;
            if (sty) val *= sty[c];
        }
        out[idx] = __float2half_rn(fminf(fmaxf(val, -65504.f), 65504.f));
    }
}

__global__ void __launch_bounds__(kRfThreads) conv_rowfold_kernel(const __grid_constant__ RowFoldParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    uint8_t* a_base = smem;                                              // [plane][column][16 B]
    uint8_t* b_base = a_base + kRfPlanes * kRfPA * 16;                   // [tap][plane][BN][16 B]
    const uint32_t b_tap_bytes = (uint32_t)(kRfPlanes * p.BN * 16);
    float* s_shift = reinterpret_cast<float*>(b_base + (size_t)p.ks * b_tap_bytes);
    int2* s_tab = reinterpret_cast<int2*>(s_shift + p.BN);               // per GEMM channel: (plane offset of its input channel, kernel row - pad)
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_tab + 8 * kRfPlanes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int HW = p.H * p.W, pad = p.ks >> 1;

    if (warp == 0 && lane == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // one-time staging: zero A (padding planes / columns stay zero for the life of the CTA), the packed weights, epilogue constants, row table
        uint4* z = reinterpret_cast<uint4*>(a_base);
        for (int i = threadIdx.x; i < kRfPlanes * kRfPA; i += kRfThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        const uint4* wsrc = reinterpret_cast<const uint4*>(p.wpack);
        uint4* wdst = reinterpret_cast<uint4*>(b_base);
        const int nw = (int)((size_t)p.ks * b_tap_bytes / 16);
        for (int i = threadIdx.x; i < nw; i += kRfThreads) wdst[i] = __ldg(wsrc + i);
        for (int j = threadIdx.x; j < p.BN; j += kRfThreads) s_shift[j] = (j < p.Cout && p.bias) ? p.bias[j] * p.gain : 0.f;
        for (int r = threadIdx.x; r < 8 * kRfPlanes; r += kRfThreads) {
            const int c = r / p.ks, kh = r - c * p.ks;
            s_tab[r] = r < p.nrows ? make_int2(c * HW + (kh - pad) * p.W, kh - pad) : make_int2(0, 1 << 20);      // padding channel: row out of range
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t issue = elect_one();
    const int nplanes = (p.nrows + 7) >> 3;
    const int ncols = 128 + p.ks - 1;                                   // staged columns that carry data
    const float slope = (p.act == PG_ACT_LINEAR) ? 1.f : (p.act == PG_ACT_RELU ? 0.f : p.alpha);
    const bool do_act = p.act != PG_ACT_LINEAR, do_clamp = p.clamp >= 0.f;
    uint32_t phase = 0;
    // tile -> (segment of 128 columns, row, sample), advanced by gridDim.x per iteration with carries: the divisions happen once per CTA, not per tile
    int wseg = (int)blockIdx.x % p.wsegs, hn0 = (int)blockIdx.x / p.wsegs, h = hn0 % p.H, n = hn0 / p.H;
    const int step_w = (int)gridDim.x % p.wsegs, step_hn = (int)gridDim.x / p.wsegs, step_h = step_hn % p.H, step_n = step_hn / p.H;
    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const int w0 = wseg * 128;
        const float* xn = p.x + (size_t)n * p.Cin * HW + h * p.W;
        // ---- build: A[plane j][column t] = the 8 (channel, kernel row) pairs of plane j at image column w0 - pad + t (zeros outside the image)
        for (int task = threadIdx.x; task < nplanes * kRfPA; task += kRfThreads) {
            const int j = task / kRfPA, t = task - j * kRfPA;
            const int col = w0 - pad + t;
            const bool col_ok = t < ncols && col >= 0 && col < p.W;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int2 tb = s_tab[j * 8 + i];                          // (element offset of (channel, kernel row) relative to row h, kernel row - pad)
                v[i] = (col_ok && (unsigned)(h + tb.y) < (unsigned)p.H) ? __ldg(xn + tb.x + col) : 0.f;
            }
            *reinterpret_cast<uint4*>(a_base + (size_t)task * 16) = pack_half8(v);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- MMA: k horizontal taps x K / 16 chunks into one accumulator
        if (warp == 0) {
            tc_fence_after();
            if (issue) {
                const uint32_t hi = (128u >> 4) | (1u << 14);
                const uint32_t a_lo = ((uint32_t)kRfPA << 16) | (smem_u32(a_base) >> 4);
                const uint32_t b_lo = (((uint32_t)p.BN & 0x3FFF) << 16) | (smem_u32(b_base) >> 4);
                for (int kw = 0; kw < p.ks; kw++)
                    for (int c = 0; c < p.nchunks; c++) {
                        const uint64_t adesc = ((uint64_t)hi << 32) | (uint64_t)(a_lo + (uint32_t)(c * 2 * kRfPA + kw));
                        const uint64_t bdesc = ((uint64_t)hi << 32) | (uint64_t)(b_lo + (((uint32_t)kw * b_tap_bytes + (uint32_t)(c * 2 * p.BN * 16)) >> 4));
                        umma_f16(tmem_base, adesc, bdesc, p.idesc, (kw | c) ? 1u : 0u);
                    }
                umma_commit(smem_u32(bar));
            }
            __syncwarp();
        }
        // ---- epilogue: warps (quarter = TMEM lane group, half = alternate 16-column chunks)
        mbar_wait(smem_u32(bar), phase);
        phase ^= 1u;
        tc_fence_after();
        {
            const int quarter = warp & 3, half = warp >> 2;
            const int w = w0 + quarter * 32 + lane;
            for (int cc = half; cc < p.BN / 16; cc += kRfThreads / 128) {
                uint32_t rg[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cc * 16), rg);
                float v[16];
                if (do_act && slope == 0.f && !do_clamp) {                // relu, no clamp (the encoder stems): two instructions per value
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = fmaxf(fmaf(__uint_as_float(rg[i]), p.gain, s_shift[cc * 16 + i]), 0.f);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        float t = fmaf(__uint_as_float(rg[i]), p.gain, s_shift[cc * 16 + i]);
                        if (do_act) t = fmaxf(t, 0.f) + slope * fminf(t, 0.f);
                        if (do_clamp) t = fminf(fmaxf(t, -p.clamp), p.clamp);
                        v[i] = t;
                    }
                }
                if (p.y_c8) {
                    uint4* yb = reinterpret_cast<uint4*>(p.y) + ((size_t)n * p.cb_out + (size_t)cc * 2) * HW + (size_t)h * p.W + w;
                    yb[0] = pack_half8(v); yb[HW] = pack_half8(v + 8);
                } else {
                    float* yp = reinterpret_cast<float*>(p.y) + ((size_t)n * p.Cout + (size_t)cc * 16) * HW + (size_t)h * p.W + w;
#pragma unroll
                    for (int i = 0; i < 16; i++) if (cc * 16 + i < p.Cout) yp[(size_t)i * HW] = v[i];
                }
            }
        }
        tc_fence_before();
        __syncthreads();                 // every warp is done with the accumulator and the MMAs are done with the A stage
        wseg += step_w; h += step_h; n += step_n;
        if (wseg >= p.wsegs) { wseg -= p.wsegs; h++; }
        if (h >= p.H) { h -= p.H; n++; }
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- host side
struct ConvPlan {
    int BN, ntiles_n, nchunks, ntaps, NACC, PW, Lp, tiles_per_img, PA, SA, SB, tps;
    uint32_t a_stage, b_stage, b_slot, b_tile; size_t smem; uint32_t tmem_cols; int nvirt;
    int tma_rows; uint32_t a_lbo16;
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Plan / loader choices that were measured against each other.  Defaults are the measured best; the environment is read ONCE, when
// the library is first used, and pg_set_tuning() changes a value for the rest of the process (tests and tools/ compare variants through it).
struct Tuning {
    int bands = 1, band_tw = 64, band_minw = 128, band_ratio10 = 0, persist = 0, nacc = 0, pair = 1, vec2 = 1, lean = 1, cgroups = 1, tma = 1, tma_persist = 1, rowfold = 1;
#ifdef PG_DEBUG
    int pipe = 0, ldmode = 0, dbgmode = 0;
#endif
};
struct TuningKey { const char* key; const char* env; int Tuning::*field; };
static const TuningKey kTuningKeys[] = {
    {"conv_bands", "PASTA_B200_CONV_BANDS", &Tuning::bands}, {"conv_band_tw", "PASTA_B200_CONV_BAND_TW", &Tuning::band_tw},
    {"conv_band_minw", "PASTA_B200_CONV_BAND_MINW", &Tuning::band_minw}, {"conv_band_ratio10", "PASTA_B200_CONV_BAND_RATIO10", &Tuning::band_ratio10},
    {"conv_persist", "PASTA_B200_CONV_PERSIST", &Tuning::persist}, {"conv_nacc", "PASTA_B200_CONV_NACC", &Tuning::nacc},
    {"conv_pair", "PASTA_B200_CONV_PAIR", &Tuning::pair}, {"conv_vec2", "PASTA_B200_CONV_VEC2", &Tuning::vec2},
    {"conv_lean", "PASTA_B200_CONV_LEAN", &Tuning::lean}, {"conv_cgroups", "PASTA_B200_CONV_CGROUPS", &Tuning::cgroups},
    {"conv_tma", "PASTA_B200_CONV_TMA", &Tuning::tma}, {"conv_tma_persist", "PASTA_B200_CONV_TMA_PERSIST", &Tuning::tma_persist},
    {"conv_rowfold", "PASTA_B200_CONV_ROWFOLD", &Tuning::rowfold},
#ifdef PG_DEBUG
    {"conv_pipe", "PASTA_B200_CONV_PIPE", &Tuning::pipe}, {"conv_ldmode", "PASTA_B200_CONV_LDMODE", &Tuning::ldmode},
    {"conv_dbgmode", "PASTA_B200_CONV_DBGMODE", &Tuning::dbgmode},
#endif
};
static Tuning& tuning() {
    static Tuning t = [] {
        Tuning v;
        for (const TuningKey& k : kTuningKeys) { const char* e = getenv(k.env); if (e && *e) v.*(k.field) = atoi(e); }
        return v;
    }();
    return t;
}

static int make_plan(ConvPlan& pl, int N, int Cin, int Cout, int H, int W, int ks, int up2, bool band = false, int max_nacc = 4, bool tma = false, int n_tile = 0,
                     int forced_nacc = 0, int wide = 0) {
    // wide: tf32 operands (4-byte elements): a 16-channel chunk is four planes of four channels, stages and weight tiles are twice as large
    const Tuning& tn = tuning();
    pl.nvirt = up2 ? 4 * Cout : Cout;
    pl.ntaps = ks * ks;
    pl.nchunks = (Cin + kKC - 1) / kKC;
    int bn = round_up(pl.nvirt, 16);
    if (bn > 256) bn = 256;
    if (up2 && pl.nvirt > 256) bn = (2 * Cout <= 256 && (2 * Cout) % 16 == 0) ? 2 * Cout : 256;   // keep both x-phases of a row parity together
    if (n_tile > 0 && n_tile % 16 == 0 && n_tile < bn && !up2) bn = n_tile;                        // caller-chosen N tile (must match the packed weights)
    pl.BN = bn;
    pl.ntiles_n = (pl.nvirt + bn - 1) / bn;
    pl.PW = (ks == 3 && !band) ? W + 1 : W;          // band mode: the halo columns of the band are its own padding
    pl.Lp = H * pl.PW;
    // Two co-resident CTAs per SM when the N tile is narrow (BN <= 128): each gets half of TMEM (256 columns) and ~100 KB of shared
    // memory, so one CTA's prologue / pipeline fill / epilogue overlaps the other's main loop.  Wide tiles (BN = 256) keep the SM alone.
    bool pair = bn <= 128;
    const int force_nacc = forced_nacc ? forced_nacc : tn.nacc;
    if (tn.pair == 0 || force_nacc * bn > 256) pair = false;
    const int tmem_budget = pair ? 256 : 512;
    const int max_acc = tmem_budget / bn < max_nacc ? tmem_budget / bn : max_nacc;
    int nacc = 1;
    for (int cand = (max_acc < 4 ? max_acc : 4); cand >= 1; cand >>= 1) {
        const long long ctas = (long long)N * ((pl.Lp + 128 * cand - 1) / (128 * cand)) * pl.ntiles_n;
        if (ctas >= 2 * kNumSMs * (pair ? 2 : 1) || cand == 1) { nacc = cand; break; }
    }
    if (!pair && max_acc >= 2 && ks == 3) {
        // wide tiles (one CTA per SM): pick the strip length by (fill of the last wave) x (useful fraction of the staged strip); measured on
        // 256->256 @64^2 (2 accumulators: 80 us vs 103), 512->512 @32^2 (1: 84 vs 125) and 512->256 up-2 @32^2 (1: 189 vs 230)
        double best = -1.0;
        for (int cand = 1; cand <= (max_acc < 4 ? max_acc : 4); cand <<= 1) {
            const long long ctas = (long long)N * ((pl.Lp + 128 * cand - 1) / (128 * cand)) * pl.ntiles_n;
            const double wave_fill = (double)ctas / (double)(((ctas + kNumSMs - 1) / kNumSMs) * kNumSMs);
            const double useful = 128.0 * cand / (128.0 * cand + 2.0 * pl.PW + 2.0);
            if (wave_fill * useful > best) { best = wave_fill * useful; nacc = cand; }
        }
    }
    if (force_nacc && force_nacc * bn <= 512) nacc = force_nacc;
    while (nacc > 1 && 128 * (nacc - 1) >= pl.Lp) nacc--;
    pl.NACC = nacc;
    const int BM = 128 * nacc;
    pl.tiles_per_img = (pl.Lp + BM - 1) / BM;
    const int halo = (ks == 3) ? 2 * pl.PW + 2 : 0;
    pl.PA = round_up(BM + halo, 32);
    pl.a_stage = (uint32_t)pl.PA * 32u * (wide ? 2u : 1u);     // 2 planes x 16 B per position (tf32: 4 planes)
    pl.tma_rows = 0; pl.a_lbo16 = (uint32_t)pl.PA;
    if (tma) {
        // the box starts at the image row that holds the tile's first staged position: up to PW - 1 positions before it, BM + halo after
        pl.tma_rows = (pl.PW - 1 + BM + halo + pl.PW - 1) / pl.PW;
        pl.a_lbo16 = (uint32_t)(pl.tma_rows * pl.PW);
        pl.a_stage = (uint32_t)round_up(2 * pl.tma_rows * pl.PW * 16, 128);
    }
    pl.b_tile = (uint32_t)bn * 32u * (wide ? 2u : 1u);
    pl.b_stage = pl.b_tile * pl.ntaps;
    const size_t fixed = (size_t)pl.nchunks * kKC * 4 + (size_t)bn * 8 + 96 * 8 + ((ks == 1 && Cin <= 160) ? 160 * 8 : 0);   // last term: folded-tap offset table
    pl.tps = (ks == 3) ? 3 : 1;
    pl.b_slot = pl.b_tile * pl.tps;
    size_t budget = pair ? 100 * 1024 : 200 * 1024;
    pl.SA = 3;
    // small strips (4^2 .. 16^2 images): the layer is a latency chain of nchunks (load chunk -> convert -> MMA) round trips, so run the ring deep
    if (!tma && (size_t)8 * pl.a_stage <= budget / 4) pl.SA = 8;
    if ((size_t)pl.SA * pl.a_stage > budget / 2) pl.SA = 2;
    if ((size_t)pl.SA * pl.a_stage + 2 * (size_t)pl.b_slot + fixed + 128 > budget) budget = 200 * 1024;   // large strips: one CTA per SM after all
    if (pl.nchunks < pl.SA) pl.SA = pl.nchunks < 2 ? 2 : pl.nchunks;
    const size_t left = budget > (size_t)pl.SA * pl.a_stage + fixed + 128 ? budget - (size_t)pl.SA * pl.a_stage - fixed - 128 : 0;
    pl.SB = (int)(left / pl.b_slot);
    const int nslots = pl.nchunks * (pl.ntaps / pl.tps);
    if (pl.SB > 24) pl.SB = 24;
    if (pl.SB > nslots) pl.SB = nslots < 2 ? 2 : nslots;
    if (pl.SB < 2) return fail(PG_ERR_UNSUPPORTED, "conv2d_igemm: tile does not fit shared memory (W=%d, BN=%d)", W, bn);
    auto total = [&]() { return (size_t)pl.SA * pl.a_stage + (size_t)pl.SB * pl.b_slot + fixed + 128; };
    pl.smem = total();
    uint32_t cols = 32;
    while (cols < (uint32_t)(nacc * bn)) cols <<= 1;
    pl.tmem_cols = cols;
    return PG_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// Tensor map of a channel-blocked fp16 activation [N][CB][H][W][8] for the A operand: a row of one channel block is W x 16 contiguous bytes, described as
// 2 W eight-byte elements so that a box row of up to 128 strip positions is ONE contiguous run; box = {2 PW, rows, 2 channel blocks}.  Positions
// outside the image (the strip's zero column, halo rows above / below, band halos beyond the edges) are the map's zero fill.
static int make_a_tensor_map(CUtensorMap& tm, const void* x, int N, int CB, int H, int W, int PW, int rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail(PG_ERR_CUDA, "conv2d_igemm: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)2 * W, (cuuint64_t)H, (cuuint64_t)N * CB};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    const cuuint32_t box[3] = {(cuuint32_t)2 * PW, (cuuint32_t)rows, 2u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PG_ERR_CUDA, "conv2d_igemm: cuTensorMapEncodeTiled failed (%d) for W=%d H=%d PW=%d rows=%d", (int)r, W, H, PW, rows);
    return PG_OK;
}

// The same tensor seen through the space-to-depth view of the down-2 form: a box samples every second column and every second row (element strides
// 2), so one (row parity, column parity) plane of two channel blocks arrives as a dense [2 blocks][rows][PW positions][16 B] tile.  Box sizes count
// the elements traversed before striding (2 * PW columns, 2 * rows rows; both <= 256); out-of-range coordinates (the pad of the 'same' 3x3 over
// the planes, negative or past the image) are zero-filled.
static int make_a_tensor_map_down2(CUtensorMap& tm, const void* x, int N, int CB, int Hin, int Win, int PW, int rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail(PG_ERR_CUDA, "conv2d_igemm: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[4] = {2u, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N * CB};
    const cuuint64_t strides[3] = {16u, (cuuint64_t)Win * 16, (cuuint64_t)Hin * Win * 16};
    const cuuint32_t box[4] = {2u, (cuuint32_t)2 * PW, (cuuint32_t)2 * rows, 2u};
    const cuuint32_t estr[4] = {1u, 2u, 2u, 1u};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PG_ERR_CUDA, "conv2d_igemm: cuTensorMapEncodeTiled (down-2 view) failed (%d) for W=%d H=%d PW=%d rows=%d", (int)r, Win, Hin, PW, rows);
    return PG_OK;
}

}  // namespace pg

extern "C" int pg_set_tuning(const char* key, int32_t value) {
    using namespace pg;
    PG_REQUIRE(key != nullptr, "pg_set_tuning: key must not be NULL");
    for (const TuningKey& k : kTuningKeys)
        if (strcmp(k.key, key) == 0) { tuning().*(k.field) = value; return PG_OK; }
    return fail(PG_ERR_INVALID_ARGUMENT, "pg_set_tuning: unknown key '%s'", key);
}

#ifdef PG_DEBUG
static long long* g_conv_dbg = nullptr;
extern "C" void pg_debug_set_buffer(void* buf) { g_conv_dbg = (long long*)buf; }
#endif

// Small-K layers (first layers: 3 -> 64 channels, 7x7 or 3x3): the k x k taps are folded into the GEMM K dimension ("im2col" virtual channels
// v = c * k * k + kh * k + kw, i.e. the weight tensor's own memory order) and the layer runs as a 1x1 convolution over Cin * k * k channels,
// instead of k * k shifted MMAs over a 16-channel chunk that is mostly padding.
static bool use_im2col(int Cin, int ksize, int up) { return up == 1 && ksize > 1 && Cin * ksize * ksize <= 160; }

// Row-folded form of the same layers (conv_rowfold_kernel): its packed weights follow the folded-tap pack in the same workspace, so the choice
// between the two can be made per launch (the row-folded kernel needs W % 128 == 0 and a plain epilogue).
static bool rowfold_shape_ok(int Cin, int Cout, int ksize, int up) {
    // (k = 1 with a handful of input channels -- the 6-channel pose stem -- is the degenerate case: one tap, Cin "rows")
    return (use_im2col(Cin, ksize, up) || (up == 1 && ksize == 1 && Cin < 16)) && Cin * ksize <= 8 * pg::kRfPlanes && ksize <= 7 && Cout >= 16 && Cout <= 128;
}
static int64_t rowfold_pack_bytes(int Cin, int Cout, int ksize, int up) {
    return rowfold_shape_ok(Cin, Cout, ksize, up) ? (int64_t)ksize * pg::kRfPlanes * ((Cout + 15) / 16 * 16) * 16 : 0;
}

extern "C" int64_t pg_conv2d_igemm_workspace_bytes_fmt(int32_t Cin, int32_t Cout, int32_t ksize, int32_t up, int32_t operand_format) {
    pg::ConvPlan pl;
    const bool im2col = use_im2col(Cin, ksize, up);
    if (pg::make_plan(pl, 1, (up == PG_CONV_DOWN2 || up == PG_CONV_DOWN2_C8) ? 4 * Cin : (im2col ? Cin * ksize * ksize : Cin), Cout, 8, 8, im2col ? 1 : ksize, up == 2,
                      false, 4, false, 0, 0, operand_format == 2) != PG_OK) return -1;
    return (int64_t)pl.ntiles_n * pl.nchunks * pl.b_stage + rowfold_pack_bytes(Cin, Cout, ksize, up);
}
extern "C" int64_t pg_conv2d_igemm_workspace_bytes(int32_t Cin, int32_t Cout, int32_t ksize, int32_t up) {
    return pg_conv2d_igemm_workspace_bytes_fmt(Cin, Cout, ksize, up, 0);
}

static int conv_validate(int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up, int32_t operand_format) {
    using namespace pg;
    PG_REQUIRE(ksize == 1 || ksize == 3 || (ksize % 2 == 1 && ksize <= 7 && use_im2col(Cin, ksize, up)),
               "conv2d_igemm: kernel size must be 1 or 3, or odd <= 7 with Cin * k * k <= 160 (got k = %d, Cin = %d)", ksize, Cin);
    PG_REQUIRE(up == 1 || ((up == 2 || up == PG_CONV_DOWN2 || up == PG_CONV_DOWN2_C8) && ksize == 3), "conv2d_igemm: resample must be 1, 2 (up), PG_CONV_DOWN2 or PG_CONV_DOWN2_C8, the latter three with a 3x3 kernel");
    PG_REQUIRE((up != PG_CONV_DOWN2 && up != PG_CONV_DOWN2_C8) || (Cin % 16 == 0 && H % 2 == 0 && W % 2 == 0), "conv2d_igemm: down-2 needs Cin %% 16 == 0 and even H, W");
    PG_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1, "conv2d_igemm: bad sizes");
    PG_REQUIRE(operand_format >= 0 && operand_format <= 2, "conv2d_igemm: operand_format must be 0 (fp16), 1 (bf16) or 2 (tf32)");
    PG_REQUIRE(up == 1 || Cout % 16 == 0, "conv2d_igemm: up=2 needs Cout to be a multiple of 16");
    return PG_OK;
}

extern "C" int pg_conv2d_igemm_prepack_batched(const float* w, int64_t w_batch_stride, const float* styles, int32_t batch, const float* fir, float w_scale,
                                               int32_t Cin, int32_t Cout, int32_t ksize, int32_t up, int32_t flip_weight, int32_t operand_format, int32_t n_tile,
                                               void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace pg;
    int rc = conv_validate(1, Cin, Cout, 8, 8, ksize, up, operand_format);
    if (rc != PG_OK) return rc;
    PG_REQUIRE(up == 1 || fir != nullptr, "conv2d_igemm: resampling needs the 4x4 FIR");
    PG_REQUIRE(w && workspace, "conv2d_igemm: w and workspace must be device pointers");
    PG_REQUIRE(batch >= 1 && batch <= 65535 && w_batch_stride >= 0, "conv2d_igemm: batch must be in [1, 65535]");
    const bool down2 = up == PG_CONV_DOWN2 || up == PG_CONV_DOWN2_C8;
    ConvPlan pl;
    const bool im2col = use_im2col(Cin, ksize, up);
    rc = make_plan(pl, 1, down2 ? 4 * Cin : (im2col ? Cin * ksize * ksize : Cin), Cout, 8, 8, im2col ? 1 : ksize, up == 2, false, 4, false, n_tile, 0, operand_format == 2);
    if (rc != PG_OK) return rc;
    PG_REQUIRE(n_tile == 0 || (n_tile % 16 == 0 && Cout % n_tile == 0 && up != 2), "conv2d_igemm: n_tile must be a multiple of 16 that divides Cout (no up-2)");
    const int64_t need_main = (int64_t)pl.ntiles_n * pl.nchunks * pl.b_stage;
    const int64_t need = need_main + rowfold_pack_bytes(Cin, Cout, ksize, up);                          // per-sample stride of the workspace
    PG_REQUIRE(workspace_bytes >= need * batch, "conv2d_igemm: workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)(need * batch));
    PackParams pp;
    pp.w = w; pp.fir = fir; pp.out = workspace; pp.Cout = Cout; pp.Cin = Cin; pp.ks = ksize; pp.BN = pl.BN; pp.nchunks = pl.nchunks;
    pp.ntaps = pl.ntaps; pp.ntiles = pl.ntiles_n; pp.flip_weight = flip_weight ? 1 : 0; pp.fmt = operand_format; pp.up2 = up == 2; pp.down2 = up == PG_CONV_DOWN2_C8 ? 2 : (down2 ? 1 : 0); pp.w_scale = w_scale; pp.im2col = im2col;
    pp.w_bstride = w_batch_stride; pp.styles = styles;
    const size_t pack_total = (size_t)need_main / (operand_format == 2 ? 4 : 2);
    int pblocks = (int)((pack_total + 255) / 256);
    const int cap = kNumSMs * 16 / (batch < 16 ? batch : 16);
    if (pblocks > cap) pblocks = cap < 1 ? 1 : cap;
    if (up == 2 && pl.ntiles_n * pl.BN == 4 * Cout) {
        // (every row of every N tile is a real virtual channel: nothing left for the element-wise kernel to zero-fill)
        const long long work = (long long)Cout * pl.nchunks * kKC;
        long long blocks = (work + 255) / 256;
        const long long bcap = (long long)kNumSMs * 32 / (batch < 32 ? batch : 32);
        if (blocks > bcap) blocks = bcap < 1 ? 1 : bcap;
        conv_prepack_up2_kernel<<<dim3((unsigned)blocks, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(pp, (long long)(need / 2));
    } else if (up == 1 && !im2col && ksize <= 3) {
        const unsigned blocks = (unsigned)(pl.ntiles_n * pl.nchunks * ((pl.BN + kPackRows - 1) / kPackRows));
        const size_t smem = (size_t)kPackRows * (16 * ksize * ksize + 1) * sizeof(float);
        conv_prepack_tiled_kernel<<<dim3(blocks, (unsigned)batch), 256, smem, (cudaStream_t)stream>>>(pp, (long long)(need / 2));
    } else
    conv_prepack_kernel<<<dim3((unsigned)pblocks, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(pp, (long long)(need / 2));
    if (need > need_main) {
        // the same weights in the row-folded layout, behind the folded-tap pack (conv_rowfold_kernel)
        PackParams pr = pp;
        pr.out = (uint8_t*)workspace + need_main; pr.BN = round_up(Cout, 16);
        const int rblocks = (int)(((size_t)(need - need_main) / 2 + 255) / 256);
        conv_rowfold_prepack_kernel<<<dim3((unsigned)rblocks, (unsigned)batch), 256, 0, (cudaStream_t)stream>>>(pr, (long long)(need / 2));
        return launch_status("conv2d_igemm_prepack", 2);
    }
    return launch_status("conv2d_igemm_prepack", 1);
}

extern "C" int pg_conv2d_igemm_prepack(const float* w, const float* fir, float w_scale, int32_t Cin, int32_t Cout, int32_t ksize, int32_t up,
                                       int32_t flip_weight, int32_t operand_format, void* workspace, int64_t workspace_bytes, void* stream) {
    return pg_conv2d_igemm_prepack_batched(w, 0, nullptr, 1, fir, w_scale, Cin, Cout, ksize, up, flip_weight, operand_format, 0, workspace, workspace_bytes, stream);
}

// Launch of conv_rowfold_kernel (eligibility is checked by the caller).  The packed weights sit behind the folded-tap pack of the same workspace.
static int launch_rowfold(const pg_conv_args& a) {
    using namespace pg;
    RowFoldParams p;
    memset(&p, 0, sizeof(p));
    ConvPlan pl;
    int rc = make_plan(pl, 1, a.Cin * a.ksize * a.ksize, a.Cout, 8, 8, 1, false);
    if (rc != PG_OK) return rc;
    p.x = (const float*)a.x; p.wpack = (const uint8_t*)a.wpack + (size_t)pl.ntiles_n * pl.nchunks * pl.b_stage; p.bias = a.bias; p.y = a.y;
    p.N = a.N; p.Cin = a.Cin; p.Cout = a.Cout; p.H = a.H; p.W = a.W; p.ks = a.ksize; p.BN = round_up(a.Cout, 16);
    p.nrows = a.Cin * a.ksize; p.nchunks = (p.nrows + 15) / 16;
    p.act = a.act; p.alpha = a.alpha; p.gain = a.gain; p.clamp = a.clamp;
    p.y_c8 = a.y_layout == PG_LAYOUT_C8; p.cb_out = a.Cout / 8;
    p.idesc = (1u << 4) | ((uint32_t)(p.BN >> 3) << 17) | (8u << 24);
    uint32_t cols = 32;
    while (cols < (uint32_t)p.BN) cols <<= 1;
    p.tmem_cols = cols;
    p.wsegs = a.W / 128;
    const long long tiles = (long long)a.N * a.H * p.wsegs;
    PG_REQUIRE(tiles < (1ll << 31), "conv2d_igemm: too many tiles");
    p.tiles_total = (int)tiles;
    size_t smem = (size_t)kRfPlanes * kRfPA * 16 + (size_t)a.ksize * kRfPlanes * p.BN * 16 + (size_t)p.BN * 4 + 8 * kRfPlanes * 8 + 64;
    // co-resident CTAs share the SM's 512 TMEM columns: never let more CTAs fit than can allocate (a CTA spinning in tcgen05.alloc until a
    // persistent neighbour exits would run its tiles alone at the end)
    const int want = (int)(512 / cols) < 5 ? (int)(512 / cols) : 5;
    const size_t floor_smem = (size_t)(227 * 1024) / (size_t)(want + 1) + 1024;
    if (smem < floor_smem) smem = floor_smem;
    PG_CUDA(cudaFuncSetAttribute(conv_rowfold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long cap = (long long)kNumSMs * want;
    conv_rowfold_kernel<<<(unsigned)(tiles < cap ? tiles : cap), kRfThreads, smem, (cudaStream_t)a.stream>>>(p);
    return launch_status("conv2d_igemm(rowfold)", 1);
}

static int conv_run_impl(const pg_conv_args& a) {
    using namespace pg;
    const Tuning& tn = tuning();
    int32_t N = a.N, Cin = a.Cin, Cout = a.Cout, H = a.H, W = a.W, ksize = a.ksize;
    const int32_t up = a.up, operand_format = a.operand_format;
    int rc = conv_validate(N, Cin, Cout, H, W, ksize, up, operand_format);
    if (rc != PG_OK) return rc;
    PG_REQUIRE(a.act == PG_ACT_LINEAR || a.act == PG_ACT_RELU || a.act == PG_ACT_LRELU, "conv2d_igemm: epilogue act must be linear/relu/lrelu");
    PG_REQUIRE(a.in_act == PG_ACT_LINEAR || a.in_act == PG_ACT_RELU || a.in_act == PG_ACT_LRELU, "conv2d_igemm: input act must be linear/relu/lrelu");
    PG_REQUIRE((int64_t)N * Cin * H * W <= INT32_MAX && (int64_t)N * Cout * H * W * (up == 2 ? 4 : 1) <= INT32_MAX, "conv2d_igemm: tensor too large");
    PG_REQUIRE(a.gain > 0.f && a.in_gain > 0.f, "conv2d_igemm: gains must be positive (they are folded through the activation)");
    if (N == 0) return PG_OK;
    const float* x = (const float*)a.x; const float* x2 = (const float*)a.x2;
    PG_REQUIRE(x && a.wpack && a.y, "conv2d_igemm: x, packed weights and y must be device pointers");
    PG_REQUIRE(a.wpack_sample_stride >= 0 && a.wpack_sample_stride % 16 == 0, "conv2d_igemm: the per-sample weight stride must be a multiple of 16 bytes");
    const bool down2 = up == PG_CONV_DOWN2 || up == PG_CONV_DOWN2_C8;
    PG_REQUIRE(!down2 || ((uintptr_t)x & 7) == 0, "conv2d_igemm: down-2 needs an 8-byte aligned input");
    PG_REQUIRE((up == PG_CONV_DOWN2_C8) == (down2 && a.x_layout == PG_LAYOUT_C8), "conv2d_igemm: PG_CONV_DOWN2_C8 goes with a channel-blocked input (and PG_CONV_DOWN2 with a dense one): the packed weights order their channels differently");
    const int hin = H, win = W, cin_real = Cin;
    const bool scale = a.styles != nullptr || a.in_act != PG_ACT_LINEAR || a.in_gain != 1.f;
    PG_REQUIRE((a.x_dtype == PG_F32 || a.x_dtype == PG_F16) && (a.y_dtype == PG_F32 || a.y_dtype == PG_F16), "conv2d_igemm: x / y must be float32 or float16");
    PG_REQUIRE((a.x_layout == PG_LAYOUT_NCHW || a.x_layout == PG_LAYOUT_C8) && (a.y_layout == PG_LAYOUT_NCHW || a.y_layout == PG_LAYOUT_C8), "conv2d_igemm: unknown tensor layout");
    const bool tma = a.x_layout == PG_LAYOUT_C8;
    const int n_tile = a.n_tile;
    PG_REQUIRE(n_tile == 0 || (n_tile % 16 == 0 && Cout % n_tile == 0 && up != 2), "conv2d_igemm: n_tile must be a multiple of 16 that divides Cout (no up-2)");
    if (tma) {
        // channel-blocked fp16 input: taken as the operand bits by TMA, so no input scale / activation / concat, fp16 operands, whole 16-channel chunks
        PG_REQUIRE(a.x_dtype == PG_F16 && !scale && (!down2 || !x2) && operand_format == 0 && Cin % 16 == 0 && !use_im2col(Cin, ksize, up) && ((uintptr_t)x & 15) == 0,
                   "conv2d_igemm: a channel-blocked input needs float16, a plain (unmodulated, no input activation) stride-1 or up-2 layer, fp16 operands and Cin %% 16 == 0");
        PG_REQUIRE(!x2 || (a.cin1 % 16 == 0 && ((uintptr_t)x2 & 15) == 0), "conv2d_igemm: a channel-blocked split input needs Cin1 %% 16 == 0 (x2 is channel-blocked float16 too)");
    }
    if (a.y_layout == PG_LAYOUT_C8)
        PG_REQUIRE(a.y_dtype == PG_F16 && (up != 2 || Cout <= 128) && (!a.residual || a.residual_layout == PG_LAYOUT_C8) &&
                   (a.spade_x ? (Cout / 2) % 16 == 0 : Cout % 16 == 0) && ((uintptr_t)a.y & 15) == 0,
                   "conv2d_igemm: a channel-blocked output needs float16, Cout %% 16 == 0 (and Cout <= 128 with up-2); a residual added to it must be channel-blocked too");
    PG_REQUIRE(a.residual_layout == PG_LAYOUT_NCHW || (a.residual_layout == PG_LAYOUT_C8 && Cout % 16 == 0 && up != 2 && !a.spade_x && ((uintptr_t)a.residual & 15) == 0),
               "conv2d_igemm: a channel-blocked residual needs Cout %% 16 == 0 and no up-sampling");
    if (tma && ksize == 1 && W > 128 && !down2) {
        // a 1x1 convolution does not care how H * W pixels are cut into rows: view a wide image as rows of 128 pixels so that a row fits one TMA box
        PG_REQUIRE(W % 128 == 0, "conv2d_igemm: a channel-blocked input of a 1x1 layer wider than 128 columns needs W %% 128 == 0");
        H *= W / 128; W = 128;
    }
    if (down2) { H /= 2; W /= 2; Cin *= 4; }                      // the GEMM runs over the space-to-depth view at the output resolution
    const bool im2col = use_im2col(Cin, ksize, up);
    const int ks_real = ksize;
    if (tn.rowfold && rowfold_shape_ok(Cin, Cout, ksize, up) && W % 128 == 0 && !scale && !x2 && !a.spade_x && !a.residual && !a.noise && !a.dcoefs &&
        a.x_dtype == PG_F32 && a.x_layout == PG_LAYOUT_NCHW && operand_format == 0 && n_tile == 0 && a.wpack_sample_stride == 0 &&
        (a.y_layout == PG_LAYOUT_C8 || a.y_dtype == PG_F32) && ((uintptr_t)a.wpack & 15) == 0)
        return launch_rowfold(a);
    if (im2col) { Cin *= ksize * ksize; ksize = 1; }              // taps folded into K: a 1x1 convolution over Cin * k * k virtual channels
    ConvPlan pl;
    // Column bands for wide images (W >= 256): with the full-width strip a 256..512-position tile is 1..2 rows and stages 2..3x what it outputs
    // (one halo row above and below); bands of 64 columns (+ 2 halo columns each side, real data) make the same tile 4..8 rows tall: 1.3x.
    // down-2: the register-batched loader covers at most 6 tasks per warp per chunk; a 4-accumulator strip of a wide image exceeds that and would
    // fall back to the generic task stream (64->64 down-2 @512^2: 827 us vs 503 us with 2 accumulators + bands)
    int max_nacc = 4;
    if (down2) {
        ConvPlan pt;
        if (make_plan(pt, N, Cin, Cout, H, W, ksize, false) == PG_OK && (2 * (pt.PA / 32) + kConvWarps - 1) / kConvWarps > 6) max_nacc = 2;
    }
    const int kBandTW = tn.band_tw;     // even
    int band_tw = 0, nbands = 1;
    const int Wimg = W;
    // (W is the GEMM's width here: the output width for down-2, the input width for up-2.)
    // TMA boxes hold at most 128 strip positions per row, so a channel-blocked input wider than 127 columns (3x3) is always processed in bands.
    const bool tma_needs_bands = tma && ksize == 3 && W + 1 > 128;
    const int wide = operand_format == 2;     // tf32 operands: generic converter path only (no bands, no lean / pair loaders, no TMA)
    if ((tn.bands || tma_needs_bands) && !wide && ksize == 3 && !im2col && W >= (tma_needs_bands ? 1 : tn.band_minw) && W % 2 == 0 &&
        (tma || (((uintptr_t)x & 7) == 0 && ((uintptr_t)x2 & 7) == 0 && tn.vec2 && tn.lean))) {
        ConvPlan pb, pf;
        const int nb = (W + kBandTW - 1) / kBandTW;
        // only where the full-width strip stages >= 2.5x what it outputs (a tile of one row), or >= 2x with an N tile of >= 128 columns (measured:
        // 256->128 @128^2 -8 %, 128->128 @128^2 -6 %); narrow N tiles at 2x lose more to the bands' 6 % of unused MMA rows than they gain
        // (64->64 @256^2 +4 %, 64->64 down-2 @512^2 +7 %)
        const int ratio10 = tn.band_ratio10;
        bool worth = make_plan(pf, N, Cin, Cout, H, W, ksize, up == 2, false, max_nacc, false, n_tile) == PG_OK;
        if (worth) {
            const int staged10 = 10 * pf.PA / (128 * pf.NACC);
            worth = ratio10 ? staged10 >= ratio10 : (staged10 >= 25 || (staged10 >= 20 && pf.BN >= 128));
        }
        if ((worth || tma_needs_bands) && make_plan(pb, N * nb, Cin, Cout, H, kBandTW + 4, ksize, up == 2, true, max_nacc, tma, n_tile) == PG_OK) {
            const int pairs = (pb.PA + 3) / 2, nt = 2 * ((pairs + 31) / 32);
            if (tma || down2 || (nt + kConvWarps - 1) / kConvWarps <= 6) { pl = pb; band_tw = kBandTW; nbands = nb; W = kBandTW + 4; }   // down-2: generic task stream, any count
        }
    }
    PG_REQUIRE(!tma_needs_bands || band_tw, "conv2d_igemm: no band plan for a channel-blocked input of width %d", Wimg);
    if (!band_tw) {
        rc = make_plan(pl, N, Cin, Cout, H, W, ksize, up == 2, false, max_nacc, tma, n_tile, 0, wide);
        if (rc != PG_OK) return rc;
    }
    PG_REQUIRE(!tma || (2 * pl.PW <= 256 && pl.tma_rows <= 256), "conv2d_igemm: TMA box out of range (PW=%d rows=%d)", pl.PW, pl.tma_rows);
    cudaStream_t s = (cudaStream_t)a.stream;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.band_tw = band_tw; p.nbands = nbands; p.Wimg = Wimg;
    p.down2 = down2; p.cin_real = cin_real; p.hin = hin; p.win = win;
    PG_REQUIRE(!x2 || (!down2 && a.cin1 > 0 && a.cin1 < Cin && a.cin1 % 8 == 0), "conv2d_igemm: the split input needs 0 < Cin1 < Cin, Cin1 %% 8 == 0 and no down-sampling");
    PG_REQUIRE(!(im2col && x2), "conv2d_igemm: the split input is not available for folded-tap (small Cin) layers");
    p.x2 = x2; p.cin1 = x2 ? a.cin1 : (im2col ? cin_real : Cin); p.residual = a.residual;
    p.im2col = im2col ? ks_real : 0;
    p.kk_magic = (uint32_t)((0x100000000ull + (uint64_t)(ks_real * ks_real) - 1) / (uint64_t)(ks_real * ks_real));
    p.ks_magic = (uint32_t)((0x100000000ull + (uint64_t)ks_real - 1) / (uint64_t)ks_real);
    p.x = x; p.wpack = a.wpack; p.wpack_sample_stride = a.wpack_sample_stride;
    p.styles = a.styles; p.dcoefs = a.dcoefs; p.noise = a.noise; p.bias = a.bias; p.y = (float*)a.y;
    p.noise_bstride = a.noise_batch_stride;
    p.N = N; p.Cin = Cin; p.Cout = pl.nvirt; p.H = H; p.W = W; p.ks = ksize;
    p.PW = pl.PW; p.Lp = pl.Lp; p.tiles_per_img = pl.tiles_per_img; p.NACC = pl.NACC; p.BN = pl.BN; p.nchunks = pl.nchunks; p.ntaps = pl.ntaps;
    p.PA = pl.PA; p.SA = pl.SA; p.SB = pl.SB; p.tps = pl.tps; p.a_stage_bytes = pl.a_stage; p.b_slot_bytes = pl.b_slot; p.b_tile_bytes = pl.b_tile;
    p.in_act = a.in_act; p.in_alpha = a.in_alpha; p.in_gain = a.in_gain; p.act = a.act; p.alpha = a.alpha; p.gain = a.gain; p.clamp = a.clamp; p.fmt = operand_format;
    p.idesc = (1u << 4) | ((uint32_t)operand_format << 7) | ((uint32_t)operand_format << 10) | ((uint32_t)(pl.BN >> 3) << 17) | (8u << 24);
    p.tmem_cols = pl.tmem_cols;
    p.up2 = up == 2; p.cout_real = Cout;
    p.up2_pair = (up == 2 && pl.BN % (2 * Cout) == 0 && !a.residual && ((uintptr_t)a.y & 15) == 0 && ((uintptr_t)a.noise & 7) == 0 && (a.noise_batch_stride % 2) == 0) ? 1 : 0;
    PG_REQUIRE(!(p.up2 && a.y_layout == PG_LAYOUT_C8) || p.up2_pair, "conv2d_igemm: up-2 with a channel-blocked output needs the paired-phase epilogue (Cout <= 128, aligned noise)");
    p.sp_x = a.spade_x; p.sp_mean = a.spade_mean; p.sp_rstd = a.spade_rstd; p.spade = a.spade_x != nullptr;
    if (p.spade) PG_REQUIRE(pl.ntiles_n * pl.BN == Cout && (pl.BN / 2) % 16 == 0 && up == 1 && a.spade_mean && a.spade_rstd && !a.wpack_sample_stride,
                            "conv2d_igemm_spade: every N tile must hold [gamma_t | beta_t] of whole 16-channel groups (2C = %d, tile %d)", Cout, pl.BN);
    p.tma_a = tma; p.tma_rows = pl.tma_rows; p.tma_cb = (x2 ? a.cin1 : cin_real) / 8; p.tma_cb2 = (tma && x2) ? (cin_real - a.cin1) / 8 : 0; p.a_lbo16 = pl.a_lbo16; p.a_tx_bytes = (uint32_t)(2 * pl.tma_rows * pl.PW * 16);
    p.y_c8 = a.y_layout == PG_LAYOUT_C8; p.cb_out = (p.spade ? Cout / 2 : Cout) / 8;
    p.res_c8 = (a.residual && a.residual_layout == PG_LAYOUT_C8) ? 1 : 0;
#ifdef PG_DEBUG
    p.dbg = g_conv_dbg; p.pipe = tn.pipe; p.ldmode = tn.ldmode; p.dbgmode = tn.dbgmode;
#endif
    p.lean = tn.lean;
    p.cgroups = 1;
    p.pw_magic = (uint32_t)((0x100000000ull + (uint64_t)pl.PW - 1) / (uint64_t)pl.PW);
    p.w_magic = (uint32_t)((0x100000000ull + (uint64_t)W - 1) / (uint64_t)W);
    p.vec2 = (!tma && !down2 && !im2col && !wide && W % 2 == 0 && ((uintptr_t)x & 7) == 0 && ((uintptr_t)x2 & 7) == 0 && tn.vec2) ? 1 : 0;
    if (wide) p.lean = 0;
    p.in_half = a.x_dtype == PG_F16; p.out_half = a.y_dtype == PG_F16;
    if (p.in_half && !tma) {
        // fp16 NCHW input: taken as the operand bits (no conversion), so no input scale / activation, fp16 operand format, the aligned pair loader
        const int pairs = (pl.PA + 3) / 2, tpw = (2 * ((pairs + 31) / 32) + kConvWarps - 1) / kConvWarps;
        PG_REQUIRE(!scale && !x2 && !down2 && !im2col && operand_format == 0 && W % 2 == 0 && ((uintptr_t)x & 3) == 0 && tpw <= 6,
                   "conv2d_igemm: a float16 input needs a plain (unmodulated, no input activation) stride-1 layer, fp16 operands, even W");
        p.vec2 = 1;
    }
    if (!tma) {   // converter warp groups: only with the lean / fp16 loaders, and only while the doubled per-warp task count stays in the register batches
        const int cg = tn.cgroups;      // measured neutral (profiles/r1 notes in DESIGN.md): off by default
        const int pairs = (pl.PA + 3) / 2, nt = 2 * ((pairs + 31) / 32);
        if (p.vec2 && (p.lean || p.in_half) && cg == 2 && (nt + kConvWarps / 2 - 1) / (kConvWarps / 2) <= 6) p.cgroups = 2;
        // tiny strips (<= 2 tasks per chunk): four groups of two warps take alternate chunks, so four chunks' load round trips are in flight
        // (the in-kernel task count only covers real image elements: H * W / 64 tasks when the whole image is one tile)
        if (p.vec2 && p.lean && !p.in_half && !band_tw && pl.SA >= 4 && tn.cgroups != 0 && p.cgroups == 1) {
            if ((long long)H * W <= 128 && (nt + 1) / 2 <= 6) p.cgroups = 4;
            else if ((long long)H * W <= 512 && (nt + 3) / 4 <= 6) p.cgroups = 2;
        }
        if (p.vec2 && p.lean && !p.in_half && p.cgroups == 1 && (nt + kConvWarps - 1) / kConvWarps > 6) p.lean = 0;
    }
    PG_REQUIRE(!(p.out_half && a.residual && !p.y_c8), "conv2d_igemm: the residual add is not available with a dense float16 output");
    p.ntiles_n = pl.ntiles_n;
    CUtensorMap tmap, tmap2;
    memset(&tmap, 0, sizeof(tmap));
    memset(&tmap2, 0, sizeof(tmap2));
    p.tma_down2 = (tma && down2) ? 1 : 0;
    if (tma) {
        rc = down2 ? make_a_tensor_map_down2(tmap, x, N, p.tma_cb, hin, win, pl.PW, pl.tma_rows) : make_a_tensor_map(tmap, x, N, p.tma_cb, H, Wimg, pl.PW, pl.tma_rows);
        if (rc != PG_OK) return rc;
        if (x2) {
            rc = make_a_tensor_map(tmap2, x2, N, p.tma_cb2, H, Wimg, pl.PW, pl.tma_rows);
            if (rc != PG_OK) return rc;
        }
    }
    if (!tma) {   // persistent variant (one CTA per SM, double-buffered TMEM, dedicated epilogue warps) where it applies and there is more than one wave of tiles
        const long long total_tiles = (long long)N * pl.tiles_per_img * pl.ntiles_n;
        const int pairs = (pl.PA + 3) / 2, nt = 2 * ((pairs + 31) / 32), tpw_p = (nt + kPConvWarps / 2 - 1) / (kPConvWarps / 2);
        const bool ok = tn.persist != 0 && !band_tw &&      // opt-in: measured slower than two co-resident one-tile CTAs
                        p.vec2 && (p.lean || p.in_half) && up == 1 && !down2 && !im2col && !p.spade && !p.y_c8 &&
                        pl.BN <= 128 && pl.BN * pl.NACC <= 256 && tpw_p <= 6 && total_tiles > 2 * kNumSMs && total_tiles < (1ll << 30);
        if (ok) {
            const size_t fixed_p = (size_t)2 * pl.nchunks * kKC * 4 + (size_t)2 * pl.BN * 8 + (size_t)(2 * 4 + 2 * 24 + 4) * 8 + 64;
            const size_t budget = 200 * 1024;
            int SA = 4;
            while (SA > 2 && (size_t)SA * pl.a_stage + 2 * (size_t)pl.b_slot + fixed_p + 128 > budget) SA--;
            const size_t used = (size_t)SA * pl.a_stage + fixed_p + 128;
            int SB = used < budget ? (int)((budget - used) / pl.b_slot) : 0;
            if (SB > 24) SB = 24;
            if (SB >= 2) {
                p.SA = SA; p.SB = SB; p.cgroups = 2;
                const size_t smem_p = (size_t)SA * pl.a_stage + (size_t)SB * pl.b_slot + fixed_p + 128;
                auto pk = scale ? conv_igemm_persistent_kernel<true> : conv_igemm_persistent_kernel<false>;
                PG_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
                const unsigned gridp = (unsigned)(total_tiles < kNumSMs ? total_tiles : kNumSMs);
                pk<<<gridp, kPThreads, smem_p, s>>>(p);
                return launch_status("conv2d_igemm(persistent)", 1);
            }
        }
    }
    if (tma && tn.tma_persist) {
        // persistent TMA kernel: one CTA per SM, two TMEM accumulator buffers of <= 256 columns.  Re-plan the strip length for that budget.
        ConvPlan pp;
        const int wband = band_tw ? kBandTW + 4 : W;
        const long long nimg = (long long)N * nbands;
        // accumulators per buffer: as many as fit 256 columns (<= 2 keeps the staged box small), fewer if the image is shorter
        int nacc = 256 / pl.BN; if (nacc > 2) nacc = 2; if (nacc < 1) nacc = 1;
        const int prc = make_plan(pp, (int)nimg, Cin, Cout, H, wband, ksize, up == 2, band_tw != 0, 2, true, n_tile, nacc);
        if (prc == PG_OK && pp.NACC * pp.BN <= 256 && pp.BN == pl.BN) {
            const long long total_tiles = nimg * pp.tiles_per_img * pp.ntiles_n;
            const size_t fixed_p = (size_t)4 * pp.BN * 4 + (size_t)(2 * 8 + 2 * 24 + 4) * 8 + 64;
            const size_t budget = 200 * 1024;
            int SA = 6;
            while (SA > 2 && (size_t)SA * pp.a_stage + 3 * (size_t)pp.b_slot + fixed_p + 128 > budget) SA--;
            const size_t used = (size_t)SA * pp.a_stage + fixed_p + 128;
            int SB = used < budget ? (int)((budget - used) / pp.b_slot) : 0;
            if (SB > 24) SB = 24;
            if (SB >= 2 && total_tiles < (1ll << 30) && 2 * pp.PW <= 256 && pp.tma_rows <= 256) {
                ConvParams q = p;
                q.PW = pp.PW; q.Lp = pp.Lp; q.tiles_per_img = pp.tiles_per_img; q.NACC = pp.NACC; q.PA = pp.PA;
                q.a_stage_bytes = pp.a_stage; q.SA = SA; q.SB = SB;
                q.tma_rows = pp.tma_rows; q.a_lbo16 = pp.a_lbo16; q.a_tx_bytes = (uint32_t)(2 * pp.tma_rows * pp.PW * 16);
                q.pw_magic = (uint32_t)((0x100000000ull + (uint64_t)pp.PW - 1) / (uint64_t)pp.PW);
                CUtensorMap m1, m2;
                memset(&m1, 0, sizeof(m1)); memset(&m2, 0, sizeof(m2));
                rc = down2 ? make_a_tensor_map_down2(m1, x, N, q.tma_cb, hin, win, pp.PW, pp.tma_rows) : make_a_tensor_map(m1, x, N, q.tma_cb, H, Wimg, pp.PW, pp.tma_rows);
                if (rc != PG_OK) return rc;
                if (x2) { rc = make_a_tensor_map(m2, x2, N, q.tma_cb2, H, Wimg, pp.PW, pp.tma_rows); if (rc != PG_OK) return rc; }
                const size_t smem_p = (size_t)SA * pp.a_stage + (size_t)SB * pp.b_slot + fixed_p + 128;
                PG_CUDA(cudaFuncSetAttribute(conv_igemm_tma_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
                const unsigned gridp = (unsigned)(total_tiles < kNumSMs ? total_tiles : kNumSMs);
                conv_igemm_tma_persistent_kernel<<<gridp, kTThreads, smem_p, s>>>(q, m1, m2);
                return launch_status("conv2d_igemm(tma persistent)", 1);
            }
        }
    }
    auto kern = scale ? conv_igemm_kernel<true> : conv_igemm_kernel<false>;
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    dim3 grid((unsigned)(N * nbands * pl.tiles_per_img), (unsigned)pl.ntiles_n);
    kern<<<grid, kConvThreads, pl.smem, s>>>(p, tmap, tmap2);
    return launch_status("conv2d_igemm", 1);
}

extern "C" int pg_conv2d_igemm_launch(const pg_conv_args* a) {
    using namespace pg;
    PG_REQUIRE(a != nullptr && a->struct_bytes == sizeof(pg_conv_args), "conv2d_igemm_launch: pg_conv_args.struct_bytes does not match this library (%u vs %u)",
               a ? a->struct_bytes : 0u, (unsigned)sizeof(pg_conv_args));
    return conv_run_impl(*a);
}

static pg_conv_args base_args(const float* x, const void* wpack, const float* styles, const float* dcoefs, const float* noise, int64_t noise_batch_stride,
                              const float* bias, float* y, int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up,
                              int32_t in_act, float in_alpha, float in_gain, int32_t act, float alpha, float gain, float clamp, int32_t operand_format, void* stream) {
    pg_conv_args a;
    memset(&a, 0, sizeof(a));
    a.struct_bytes = sizeof(a);
    a.N = N; a.Cin = Cin; a.Cout = Cout; a.H = H; a.W = W; a.ksize = ksize; a.up = up;
    a.x = x; a.wpack = wpack; a.styles = styles; a.dcoefs = dcoefs; a.noise = noise; a.noise_batch_stride = noise_batch_stride; a.bias = bias; a.y = y;
    a.in_act = in_act; a.in_alpha = in_alpha; a.in_gain = in_gain; a.act = act; a.alpha = alpha; a.gain = gain; a.clamp = clamp;
    a.operand_format = operand_format; a.stream = stream;
    return a;
}

extern "C" int pg_conv2d_igemm_run(const float* x, const void* wpack, const float* styles, const float* dcoefs,
                                   const float* noise, int64_t noise_batch_stride, const float* bias, float* y,
                                   int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up,
                                   int32_t in_act, float in_alpha, float in_gain,
                                   int32_t act, float alpha, float gain, float clamp, int32_t operand_format, void* stream) {
    return conv_run_impl(base_args(x, wpack, styles, dcoefs, noise, noise_batch_stride, bias, y, N, Cin, Cout, H, W, ksize, up, in_act, in_alpha, in_gain,
                                   act, alpha, gain, clamp, operand_format, stream));
}

extern "C" int pg_conv2d_igemm_run2(const float* x, const float* x2, int32_t Cin1, const void* wpack, const float* styles, const float* dcoefs,
                                    const float* noise, int64_t noise_batch_stride, const float* bias, const float* residual, float* y,
                                    int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up,
                                    int32_t in_act, float in_alpha, float in_gain,
                                    int32_t act, float alpha, float gain, float clamp, int32_t operand_format, int32_t x_dtype, int32_t y_dtype, void* stream) {
    pg_conv_args a = base_args(x, wpack, styles, dcoefs, noise, noise_batch_stride, bias, y, N, Cin, Cout, H, W, ksize, up, in_act, in_alpha, in_gain,
                               act, alpha, gain, clamp, operand_format, stream);
    a.x2 = x2; a.cin1 = Cin1; a.residual = residual; a.x_dtype = x_dtype; a.y_dtype = y_dtype;
    return conv_run_impl(a);
}

extern "C" int pg_conv2d_igemm_spade_run(const float* feat, const void* wpack_gamma_beta, const float* x, const float* mean, const float* rstd,
                                         float* y, int32_t N, int32_t Cin, int32_t C, int32_t H, int32_t W, int32_t ksize,
                                         int32_t act, float alpha, float gain, int32_t operand_format, int32_t feat_dtype, int32_t y_dtype, void* stream) {
    using namespace pg;
    PG_REQUIRE(x && mean && rstd, "conv2d_igemm_spade: x, mean and rstd must be device pointers");
    pg_conv_args a = base_args(feat, wpack_gamma_beta, nullptr, nullptr, nullptr, 0, nullptr, y, N, Cin, 2 * C, H, W, ksize, 1, PG_ACT_LINEAR, 0.f, 1.f,
                               act, alpha, gain, -1.f, operand_format, stream);
    a.spade_x = x; a.spade_mean = mean; a.spade_rstd = rstd; a.x_dtype = feat_dtype; a.y_dtype = y_dtype;
    return conv_run_impl(a);
}

extern "C" int pg_conv2d_igemm_fwd(const float* x, const float* w, const float* fir, const float* styles, const float* dcoefs,
                                   const float* noise, int64_t noise_batch_stride, const float* bias, float* y,
                                   int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize, int32_t up,
                                   int32_t flip_weight, int32_t in_act, float in_alpha, float in_gain,
                                   int32_t act, float alpha, float gain, float clamp, int32_t operand_format,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = pg_conv2d_igemm_prepack(w, fir, 1.f, Cin, Cout, ksize, up, flip_weight, operand_format, workspace, workspace_bytes, stream);
    if (rc != PG_OK) return rc;
    return pg_conv2d_igemm_run(x, workspace, styles, dcoefs, noise, noise_batch_stride, bias, y, N, Cin, Cout, H, W, ksize, up,
                               in_act, in_alpha, in_gain, act, alpha, gain, clamp, operand_format, stream);
}
