// mma_probe.cu -- how many cycles does ONE thread need per tcgen05.mma (kind::f16, M=128, K=16) as a function of N
// and of the number of TMEM accumulators the MMAs rotate over?  (Is a chain of MMAs into the SAME accumulator
// latency-bound for small N?)  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}
__global__ void __launch_bounds__(128) k_probe(int N, int nacc, int iters, int kk, uint32_t idesc, unsigned long long *out, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t a_addr = base, b_addr = base + 16384u;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t *>(smem + (base - smem_u32(smem)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (mode == 0 && threadIdx.x == 0) {
        const uint64_t hi = (uint64_t)(((8u * 128u) >> 4) | (1u << 14) | (2u << 29)) << 32;
        const uint64_t adesc0 = hi | (uint64_t)(((a_addr & 0x3FFFFu) >> 4) | (1u << 16));
        const uint64_t bdesc0 = hi | (uint64_t)(((b_addr & 0x3FFFFu) >> 4) | (1u << 16));
        const long long t0 = clock64();
        int acc = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem_base + (uint32_t)(acc * N);
            uint64_t ad = adesc0, bd = bdesc0;
            for (int k = 0; k < kk; ++k) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                ad += 2; bd += 2;
            }
            if (++acc == nacc) acc = 0;
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        const long long t2 = clock64();
        out[0] = (unsigned long long)(t1 - t0);
        out[1] = (unsigned long long)(t2 - t0);
    }
    if (mode == 1 && threadIdx.x < 32) {
        // the whole warp runs the loop; one elected lane issues
        const uint64_t hi = (uint64_t)(((8u * 128u) >> 4) | (1u << 14) | (2u << 29)) << 32;
        const uint64_t adesc0 = hi | (uint64_t)(((a_addr & 0x3FFFFu) >> 4) | (1u << 16));
        const uint64_t bdesc0 = hi | (uint64_t)(((b_addr & 0x3FFFFu) >> 4) | (1u << 16));
        const long long t0 = clock64();
        int acc = 0;
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem_base + (uint32_t)(acc * N);
            if (elect_one()) {
                uint64_t ad = adesc0, bd = bdesc0;
                for (int k = 0; k < kk; ++k) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                    ad += 2; bd += 2;
                }
            }
            __syncwarp();
            if (++acc == nacc) acc = 0;
        }
        const long long t1 = clock64();
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
        const long long t2 = clock64();
        if (threadIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

int main() {
    unsigned long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 512;
    printf("%5s %5s %3s | %12s %12s\n", "N", "nacc", "kk", "issue cyc/MMA", "total cyc/MMA");
    for (int mode : {0, 1})
    for (int kk : {1, 4, 8, 32})
        for (int N : {32, 64, 128, 256})
            for (int nacc : {1, 2}) {
                if (nacc * N > 512) continue;
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                for (int rep = 0; rep < 2; ++rep) {
                    k_probe<<<1, 128, 52 * 1024>>>(N, nacc, iters, kk, idesc, d, mode);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("mode %d %5d %5d %3d | %12.1f %12.1f   (floor %d)\n", mode, N, nacc, kk, (double)h[0] / (iters * kk), (double)h[1] / (iters * kk), 128 * N / 256);
            }
    return 0;
}
