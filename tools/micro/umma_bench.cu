// Micro-benchmark: tcgen05.mma throughput on B200 for the shared-memory operand layouts the conv kernel can use.
// One CTA per SM, one thread issues R MMAs (M=128, N, K=16, kind::f16) back to back on resident smem operands and
// waits for the commit; cycles per MMA = (t1 - t0) / R.  Layout variants: SWIZZLE_NONE K-major with rows 16 B apart
// (the conv kernel's "shiftable" layout) vs SWIZZLE_128B K-major.  Accumulator rotation over NACC TMEM tiles.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) bench(int N, int R, int nacc, int mode, int shift_rows, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_slot;
    const int nacc_c = nacc;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
        uint64_t ad0, bd0;
        if (mode == 0) {       // SWIZZLE_NONE, rows 16 B apart, K chunks one plane (16 KB) apart
            ad0 = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)((16384u >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
            bd0 = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)N * 16u >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
        } else {               // SWIZZLE_128B K-major: rows of 128 B, 8-row atoms 1024 B apart
            ad0 = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
            bd0 = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
        }
        const uint64_t astep = (mode == 0) ? (uint64_t)shift_rows : 2ull;   // in 16-byte units
        long long t0 = clock64();
        for (int r = 0; r < R; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint64_t ad = ad0 + astep * (uint64_t)(u & 3);
                const uint32_t d = tmem + (uint32_t)((u % nacc_c) * N);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad), "l"(bd0), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int R = 2048;
    struct Cfg { int N, nacc, mode, shift; const char* name; } cfgs[] = {
        {128, 4, 0, 0, "none  N=128 nacc=4 shift=0"}, {128, 4, 0, 131, "none  N=128 nacc=4 shift=131 rows"}, {128, 1, 0, 0, "none  N=128 nacc=1"},
        {256, 2, 0, 0, "none  N=256 nacc=2"}, {64, 4, 0, 0, "none  N=64  nacc=4"},
        {128, 4, 1, 0, "sw128 N=128 nacc=4"}, {256, 2, 1, 0, "sw128 N=256 nacc=2"}, {64, 4, 1, 0, "sw128 N=64 nacc=4"},
    };
    for (auto& c : cfgs) {
        for (int grid : {1, 148}) {
            bench<<<grid, 128, 160 * 1024>>>(c.N, R, c.nacc, c.mode, c.shift, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("%-36s grid=%3d  %7.1f cycles/MMA  (ideal %d)  %s\n", c.name, grid, (double)h / R, c.N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
