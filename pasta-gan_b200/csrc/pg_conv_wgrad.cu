// Weight gradient of the stride-1 'same' convolution on tcgen05 / TMEM.
//
// Replaces the cuDNN weight-gradient call behind the reference's conv2d_gradfix (torch_utils/ops/conv2d_gradfix.py:140-148,
// aten::cudnn_convolution_backward_weight; on current torch the aten.convolution_backward it degrades to):
//
//     dW[o, c, kh, kw] = sum_{n, h, w} dy[n, o, h, w] * x[n, c, h + kh - p, w + kw - p]            k in {1, 3}, p = k / 2
//
// GEMM view: M = output channels (tile of 128), N = input channels (tile of BN <= 48), K = pixels of all samples, one accumulator per filter tap:
// D_tap[o, c] += sum_pos dy[o, pos] * x[c, pos + shift(tap)].
//   * Both operands are staged exactly as the forward kernel stages its activations: strip positions with pitch W + 1 (one shared zero column), as
//     [plane = 8 channels][position][8 x 2 B].  For the forward GEMM that is a K-major operand (K = channels); here the SAME bytes are an MN-major
//     operand (MN = channels, K = positions): 8 consecutive positions x 16 B form a core matrix, LBO = 128 B (next 8 positions), SBO = plane stride
//     (next 8 channels).  A filter tap is again nothing but a start-address offset of the x operand: (kh * PW + kw) rows of 16 B.
//   * bf16 operands (gradients need the exponent range; 8-bit mantissa), fp32 accumulation in TMEM: taps x BN columns (9 x 48 = 432 of 512).
//   * The K dimension is split over CTAs (a CTA owns a contiguous range of 128-position chunks over all samples); each CTA writes its partial
//     [tap][o][c] block to a workspace with coalesced stores and a second kernel sums the splits into dW (deterministic, no atomics).
//   * Warp roles: 16 converter warps (fp32 NCHW -> bf16 strip stage, four tasks of loads in flight each; later the epilogue), one MMA warp; a 2..3 stage mbarrier ring.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>
#include "pg_common.cuh"

namespace pg {
namespace wg {

constexpr int kWarps = 16;
constexpr int kThreads = 32 + 32 * kWarps;       // warp 0: MMA issuer, warps 1..16: converters / epilogue
constexpr int kBatch = 4;                        // converter tasks whose loads are issued together (32 independent 128-byte requests per warp in flight)
constexpr int kChunk = 128;                      // strip positions per pipeline stage (8 MMA K steps of 16)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WG_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WG_DONE;\n\t"
        "bra WG_WAIT;\n\t"
        "WG_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

struct Params {
    const float* x; const float* dy; float* partial;
    int N, Cin, Cout, H, W, ks, ntaps;
    int PW, Lp, cps;                // strip pitch, strip length, 128-position chunks per sample
    int BN, c_tiles, o_tiles, splits, total_chunks, chunks_per_split;
    int PAx, halo;                  // staged x positions per chunk (multiple of 32), PW + 1 (k = 3) or 0
    int S;                          // pipeline stages
    uint32_t dy_stage_bytes, x_stage_bytes, pw_magic, idesc;
};

// One task = 32 strip positions x 8 channels of `src_n` (fp32 NCHW planes of one sample) -> one 16-byte bf16 row per position.  Loads and stores are
// separate so that a warp can put the loads of kBatch tasks in flight before it converts the first one.
__device__ __forceinline__ void task_load(const Params& p, const float* src_n, const int C, const int c0, const int q0, const int lane, float (&v)[8]) {
    const int HW = p.H * p.W;
    const int q = q0 + lane;
    bool ok = q >= 0 && q < p.Lp;
    int h = 0, w = 0;
    if (ok) { h = (int)__umulhi((uint32_t)q, p.pw_magic); w = q - h * p.PW; ok = w < p.W; }
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.f;
    if (ok) {
        const float* s = src_n + (size_t)c0 * HW + (size_t)h * p.W + w;
#pragma unroll
        for (int i = 0; i < 8; i++) if (c0 + i < C) v[i] = __ldg(s + (size_t)i * HW);
    }
}
__device__ __forceinline__ void task_store(const float (&v)[8], uint8_t* dst_plane, const int slot) {
    uint4 pk;
    pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]); pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst_plane + (size_t)slot * 16) = pk;
}

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int split = blockIdx.x, ct = blockIdx.y, ot = blockIdx.z;
    const int stage_bytes = (int)(p.dy_stage_bytes + p.x_stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.S * stage_bytes);
    uint64_t* full = bars, *empty = bars + p.S, *acc_full = bars + 2 * p.S;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.S; i++) { mbar_init(smem_u32(&full[i]), kWarps); mbar_init(smem_u32(&empty[i]), 1); }
        mbar_init(smem_u32(acc_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ch_lo = split * p.chunks_per_split;
    const int ch_hi = min(ch_lo + p.chunks_per_split, p.total_chunks);
    const int nch = ch_hi > ch_lo ? ch_hi - ch_lo : 0;
    const int o0 = ot * 128, c0 = ct * p.BN;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        const uint32_t issue = elect_one();
        // MN-major, SWIZZLE_NONE: LBO = 128 B (next group of 8 positions), SBO = plane stride (next 8 channels); descriptor version 1
        const uint32_t lbo = (128u >> 4) << 16;
        const uint32_t hi_a = ((uint32_t)(kChunk * 16) >> 4) | (1u << 14);                 // dy planes are kChunk positions long
        const uint32_t hi_b = ((uint32_t)(p.PAx * 16) >> 4) | (1u << 14);
        const uint32_t base = smem_u32(smem);
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < nch; i++) {
            mbar_wait(smem_u32(&full[st]), ph);
            tc_fence_after();
            const uint32_t a0 = (base + (uint32_t)st * stage_bytes) >> 4;
            const uint32_t b0 = (base + (uint32_t)st * stage_bytes + p.dy_stage_bytes) >> 4;
            if (issue) {
                for (int tap = 0; tap < p.ntaps; tap++) {
                    const uint32_t shift = p.ks == 3 ? (uint32_t)((tap / 3) * p.PW + (tap % 3)) : 0u;
#pragma unroll
                    for (int j = 0; j < kChunk / 16; j++) {
                        const uint64_t adesc = ((uint64_t)hi_a << 32) | (uint64_t)(lbo | (a0 + (uint32_t)j * 16u));
                        const uint64_t bdesc = ((uint64_t)hi_b << 32) | (uint64_t)(lbo | (b0 + shift + (uint32_t)j * 16u));
                        umma_f16(tmem_base + (uint32_t)(tap * p.BN), adesc, bdesc, p.idesc, (i | j) ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&empty[st]));
            }
            __syncwarp();
            if (++st == p.S) { st = 0; ph ^= 1; }
        }
        if (issue) umma_commit(smem_u32(acc_full));
    } else {
        // ===================== converters =====================
        const int cw = warp - 1;
        const int dy_planes = 16, x_planes = p.BN / 8;
        const int dy_tasks = dy_planes * (kChunk / 32), x_tasks = x_planes * (p.PAx / 32);
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < nch; i++) {
            const int chunk = ch_lo + i;
            const int n = chunk / p.cps, m0 = (chunk - n * p.cps) * kChunk;
            uint8_t* sdy = smem + (size_t)st * stage_bytes;
            uint8_t* sx = sdy + p.dy_stage_bytes;
            mbar_wait(smem_u32(&empty[st]), ph ^ 1);
            const float* dyn = p.dy + (size_t)n * p.Cout * p.H * p.W;
            const float* xn = p.x + (size_t)n * p.Cin * p.H * p.W;
            const int ntasks = dy_tasks + x_tasks;
            for (int tb = cw; tb < ntasks; tb += kWarps * kBatch) {
                float v[kBatch][8];
                uint8_t* dst[kBatch]; int slot[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; u++) {
                    const int t = tb + u * kWarps;
                    dst[u] = nullptr; slot[u] = 0;
                    if (t >= ntasks) continue;
                    if (t < dy_tasks) {
                        const int plane = t / (kChunk / 32), g = t - plane * (kChunk / 32);
                        task_load(p, dyn, p.Cout, o0 + plane * 8, m0 + g * 32, lane, v[u]);
                        dst[u] = sdy + (size_t)plane * kChunk * 16; slot[u] = g * 32 + lane;
                    } else {
                        const int tt = t - dy_tasks;
                        const int plane = tt / (p.PAx / 32), g = tt - plane * (p.PAx / 32);
                        task_load(p, xn, p.Cin, c0 + plane * 8, m0 - p.halo + g * 32, lane, v[u]);
                        dst[u] = sx + (size_t)plane * p.PAx * 16; slot[u] = g * 32 + lane;
                    }
                }
#pragma unroll
                for (int u = 0; u < kBatch; u++) if (dst[u]) task_store(v[u], dst[u], slot[u]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full[st]));
            if (++st == p.S) { st = 0; ph ^= 1; }
        }
        // ===================== epilogue: partial[split][tap][o][c] (coalesced along c) =====================
        if (nch > 0) {
            mbar_wait(smem_u32(acc_full), 0);
            tc_fence_after();
        }
        const int quarter = warp & 3;                              // TMEM lane quarter this warp may read
        const int o = o0 + quarter * 32 + lane;
        const int cin_pad = p.c_tiles * p.BN, cout_pad = p.o_tiles * 128;
        const int part = (cw >> 2);                                // two warps per quarter split the (tap, column chunk) list
        const int nitems = p.ntaps * (p.BN / 16);
        for (int it = part; it < nitems; it += kWarps / 4) {
            const int tap = it / (p.BN / 16), cc = it - tap * (p.BN / 16);
            uint32_t r[16];
            if (nch > 0) tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tap * p.BN + cc * 16), r);
            else {
#pragma unroll
                for (int k = 0; k < 16; k++) r[k] = 0u;
            }
            float4* dst = reinterpret_cast<float4*>(p.partial + (((size_t)split * p.ntaps + tap) * cout_pad + o) * cin_pad + c0 + cc * 16);
#pragma unroll
            for (int k = 0; k < 4; k++)
                dst[k] = make_float4(__uint_as_float(r[4 * k]), __uint_as_float(r[4 * k + 1]), __uint_as_float(r[4 * k + 2]), __uint_as_float(r[4 * k + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// dW[o][c][tap] = sum_split partial[split][tap][o][c]   (taps in cross-correlation order: the gradient of F.conv2d's weight)
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int splits, int ntaps, int Cout, int Cin,
                                                                int cout_pad, int cin_pad, float scale, int accumulate) {
    // thread order (tap, o, c) with c fastest: the reads of every split are coalesced; the (small) write to dw[o][c][tap] is strided
    const int total = Cout * Cin * ntaps;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c = idx % Cin;
        const int o = (idx / Cin) % Cout;
        const int tap = idx / (Cin * Cout);
        const float* src = partial + ((size_t)tap * cout_pad + o) * cin_pad + c;
        const size_t sstride = (size_t)ntaps * cout_pad * cin_pad;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int sp = 0;
        for (; sp + 4 <= splits; sp += 4) {
            s0 += __ldg(src + (size_t)sp * sstride); s1 += __ldg(src + (size_t)(sp + 1) * sstride);
            s2 += __ldg(src + (size_t)(sp + 2) * sstride); s3 += __ldg(src + (size_t)(sp + 3) * sstride);
        }
        for (; sp < splits; sp++) s0 += __ldg(src + (size_t)sp * sstride);
        const float s = ((s0 + s1) + (s2 + s3)) * scale;
        float* d = dw + ((size_t)o * Cin + c) * ntaps + tap;
        *d = accumulate ? *d + s : s;
    }
}

struct Plan { int BN, c_tiles, o_tiles, splits, total_chunks, chunks_per_split, PW, Lp, cps, PAx, halo, S; uint32_t dy_stage, x_stage; size_t smem; };

static int make_plan(Plan& pl, int N, int Cin, int Cout, int H, int W, int ks) {
    const int ntaps = ks * ks;
    pl.BN = ntaps == 1 ? (Cin >= 128 ? 128 : ((Cin + 15) / 16) * 16) : (Cin % 48 == 0 || Cin > 96 ? 48 : (Cin > 16 ? 32 : 16));
    if (ntaps * pl.BN > 512) pl.BN = 48;
    pl.c_tiles = (Cin + pl.BN - 1) / pl.BN;
    pl.o_tiles = (Cout + 127) / 128;
    pl.PW = ks == 3 ? W + 1 : W;
    pl.Lp = H * pl.PW;
    pl.cps = (pl.Lp + kChunk - 1) / kChunk;
    pl.total_chunks = N * pl.cps;
    pl.halo = ks == 3 ? pl.PW + 1 : 0;
    pl.PAx = ((kChunk + 2 * pl.halo + 31) / 32) * 32;
    pl.dy_stage = 16u * kChunk * 16u;
    pl.x_stage = (uint32_t)(pl.BN / 8) * (uint32_t)pl.PAx * 16u;
    const size_t stage = (size_t)pl.dy_stage + pl.x_stage;
    pl.S = 3;
    while (pl.S > 2 && pl.S * stage + 256 > 200 * 1024) pl.S--;
    if (pl.S * stage + 256 > 220 * 1024) return fail(PG_ERR_UNSUPPORTED, "conv2d_wgrad: image too wide for the staged strip (W = %d)", W);
    pl.smem = pl.S * stage + 256;
    // split K so that the grid is one wave of CTAs (one CTA per SM: 512 TMEM columns), but keep at least 4 chunks per CTA
    const int tiles = pl.c_tiles * pl.o_tiles;
    int splits = kNumSMs / tiles;
    if (splits > pl.total_chunks / 4) splits = pl.total_chunks / 4;
    if (splits < 1) splits = 1;
    pl.chunks_per_split = (pl.total_chunks + splits - 1) / splits;
    pl.splits = (pl.total_chunks + pl.chunks_per_split - 1) / pl.chunks_per_split;
    return PG_OK;
}

}  // namespace wg
}  // namespace pg

extern "C" int64_t pg_conv2d_wgrad_workspace_bytes(int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize) {
    pg::wg::Plan pl;
    if (ksize != 1 && ksize != 3) return -1;
    if (pg::wg::make_plan(pl, N, Cin, Cout, H, W, ksize) != PG_OK) return -1;
    return (int64_t)pl.splits * ksize * ksize * (pl.o_tiles * 128) * (pl.c_tiles * pl.BN) * 4;
}

extern "C" int pg_conv2d_wgrad(const float* x, const float* dy, float* dw, int32_t N, int32_t Cin, int32_t Cout, int32_t H, int32_t W, int32_t ksize,
                               float scale, int32_t accumulate, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace pg;
    using namespace pg::wg;
    PG_REQUIRE(ksize == 1 || ksize == 3, "conv2d_wgrad: kernel size must be 1 or 3 (got %d)", ksize);
    PG_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1, "conv2d_wgrad: bad sizes");
    PG_REQUIRE((int64_t)N * Cin * H * W <= INT32_MAX && (int64_t)N * Cout * H * W <= INT32_MAX, "conv2d_wgrad: tensor too large");
    PG_REQUIRE(dw != nullptr, "conv2d_wgrad: dw must be a device pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) {
        if (!accumulate) PG_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cout * Cin * ksize * ksize * 4, s));
        return PG_OK;
    }
    PG_REQUIRE(x && dy && workspace, "conv2d_wgrad: x, dy and workspace must be device pointers");
    Plan pl;
    int rc = make_plan(pl, N, Cin, Cout, H, W, ksize);
    if (rc != PG_OK) return rc;
    const int64_t need = (int64_t)pl.splits * ksize * ksize * (pl.o_tiles * 128) * (pl.c_tiles * pl.BN) * 4;
    PG_REQUIRE(workspace_bytes >= need && ((uintptr_t)workspace & 15) == 0, "conv2d_wgrad: workspace too small or unaligned (%lld < %lld)", (long long)workspace_bytes, (long long)need);
    Params p;
    memset(&p, 0, sizeof(p));
    p.x = x; p.dy = dy; p.partial = (float*)workspace;
    p.N = N; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.ks = ksize; p.ntaps = ksize * ksize;
    p.PW = pl.PW; p.Lp = pl.Lp; p.cps = pl.cps; p.BN = pl.BN; p.c_tiles = pl.c_tiles; p.o_tiles = pl.o_tiles; p.splits = pl.splits;
    p.total_chunks = pl.total_chunks; p.chunks_per_split = pl.chunks_per_split; p.PAx = pl.PAx; p.halo = pl.halo; p.S = pl.S;
    p.dy_stage_bytes = pl.dy_stage; p.x_stage_bytes = pl.x_stage;
    p.pw_magic = (uint32_t)((0x100000000ull + (uint64_t)pl.PW - 1) / (uint64_t)pl.PW);
    // D fp32, A / B bf16, both MN-major (bits 15 / 16), N = BN, M = 128
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(pl.BN >> 3) << 17) | (8u << 24);
    PG_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    dim3 grid((unsigned)pl.splits, (unsigned)pl.c_tiles, (unsigned)pl.o_tiles);
    conv_wgrad_kernel<<<grid, kThreads, pl.smem, s>>>(p);
    int st = launch_status("conv2d_wgrad", 1);
    if (st != PG_OK) return st;
    const int total = Cout * Cin * p.ntaps;
    int blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    conv_wgrad_reduce_kernel<<<blocks, 256, 0, s>>>(p.partial, dw, pl.splits, p.ntaps, Cout, Cin, pl.o_tiles * 128, pl.c_tiles * pl.BN, scale, accumulate ? 1 : 0);
    return launch_status("conv2d_wgrad(reduce)", 1);
}
