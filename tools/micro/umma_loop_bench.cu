// Micro-benchmark: the conv kernel's MMA issue loop in isolation (operands resident in shared memory, no producers).
// chunks x 3 rows x 3 taps x NACC accumulators of tcgen05.mma M=128, N, K=16 (kind::f16, SWIZZLE_NONE rows 16 B apart, tap = start-address
// shift), optional tcgen05.commit after every tap row / chunk as in the kernel.  Reports cycles per MMA for 1 and 2 co-resident CTAs per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 2) bench(int N, int chunks, int nacc, int commits, int PW, int order, uint32_t cols, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, sink[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 90 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 4; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&sink[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t PA = 128 * nacc + 2 * PW + 2;
        const uint32_t hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo = ((PA & 0x3FFF) << 16) | (smem_u32(smem) >> 4);
        const uint32_t b_lo = (((uint32_t)N & 0x3FFF) << 16) | (smem_u32(smem + 48 * 1024) >> 4);
        long long t0 = clock64();
        for (int ci = 0; ci < chunks; ci++) {
            if (order == 0) {
#pragma unroll
                for (int kh = 0; kh < 3; kh++) {
#pragma unroll
                    for (int kw = 0; kw < 3; kw++) {
                        const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(b_lo + (uint32_t)(kh * 3 + kw) * (uint32_t)(N * 2));
                        for (int a = 0; a < nacc; a++) {
                            const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + (uint32_t)(kh * PW + kw) + (uint32_t)a * 128u);
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(tmem + (uint32_t)(a * N)), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                        }
                    }
                    if (commits) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sink[kh])) : "memory");
                }
            } else {      // accumulator-major: all 9 taps of one accumulator back to back (longest dependent chains)
                for (int a = 0; a < nacc; a++) {
#pragma unroll
                    for (int t = 0; t < 9; t++) {
                        const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(b_lo + (uint32_t)t * (uint32_t)(N * 2));
                        const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + (uint32_t)((t / 3) * PW + t % 3) + (uint32_t)a * 128u);
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem + (uint32_t)(a * N)), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                    }
                }
            }
            if (commits) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&sink[3])) : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols));
}

int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int chunks = 64;
    struct Cfg { int N, nacc, commits, order; } cfgs[] = {
        {128, 1, 0, 0}, {128, 2, 0, 0}, {128, 4, 0, 0}, {128, 2, 1, 0}, {128, 4, 1, 0}, {128, 2, 0, 1}, {128, 4, 0, 1},
        {256, 1, 0, 0}, {256, 2, 0, 0}, {256, 1, 1, 0}, {256, 2, 1, 0}, {64, 4, 1, 0}, {64, 4, 0, 0}, {64, 2, 1, 0},
    };
    for (auto& c : cfgs) {
        for (int per_sm : {1, 2}) {
            uint32_t cols = 32; while (cols < (uint32_t)(c.N * c.nacc)) cols <<= 1;
            if (per_sm == 2 && cols > 256) continue;
            const size_t smem = per_sm == 1 ? 190 * 1024 : 96 * 1024;
            for (int grid : {1, 148 * per_sm}) {
                bench<<<grid, 128, smem>>>(c.N, chunks, c.nacc, c.commits, 129, c.order, cols, d);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                const double n = (double)chunks * 9 * c.nacc;
                printf("N=%3d nacc=%d commits=%d order=%d ctas/SM=%d grid=%3d  issue %6.1f  complete %6.1f cycles/MMA (ideal %d)  %s\n", c.N, c.nacc, c.commits, c.order,
                       per_sm, grid, h[0] / n, h[1] / n, c.N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
