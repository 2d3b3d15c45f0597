// issue_probe.cu -- how fast can tcgen05.mma (kind::f16, M = 128, K = 16) be ISSUED for the narrow tiles (N = 32 / 64 / 128)
// whose tensor time per instruction (128 * N / 256 cycles) is below the ~170-200 cycles per MMA the round-1 kernel's
// single issuing thread needed on the BN <= 128 layers (profiles/r01_tcstats_v7_graph.txt)?
//   mode 0: one thread, groups of G fully unrolled MMAs with precomputed descriptors, one tcgen05.commit per group
//           (what a pipeline stage of the convolution kernel looks like without its barrier waits)
//   mode 1: the same loop run by TWO threads in different warps at once, each on its own accumulator
//   mode 2: mode 0 plus a (never blocking) mbarrier try_wait + tcgen05.fence::after_thread_sync per group
//   mode 3: mode 0 without the per-group commit (MMAs only)
//   mode 4: FOUR threads in different warps, each on its own accumulator (N <= 128)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_probe issue_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma(uint32_t d, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accum) : "memory");
}

template <int G>
__device__ __forceinline__ long long issue_loop(uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, int groups, uint32_t bar,
                                                uint32_t ready_bar, bool waits, bool commits) {
    const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
        if (waits) {
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(ready_bar), "r"(1u) : "memory");   // parity 1 of a fresh barrier: completes at once
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t a0 = (((a_addr + (uint32_t)(g & 3) * 4096u) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b0 = ((b_addr & 0x3FFFFu) >> 4) | (1u << 16);
#pragma unroll
        for (int k = 0; k < G; ++k) mma(d, a0 + 2u * (uint32_t)(k & 3), b0 + 2u * (uint32_t)(k & 3), hi, idesc, (uint32_t)((g | k) != 0));
        if (commits) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    return clock64() - t0;
}

template <int G>
__global__ void __launch_bounds__(128) k_probe(int N, int groups, int mode, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (48 * 1024) / 4; i += 128) reinterpret_cast<uint32_t *>(smem + (base - smem_u32(smem)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long T0 = clock64();
    const int nthr = mode == 1 ? 2 : mode == 4 ? 4 : 1;
    if (lane == 0 && warp < nthr) {
        // every thread reads the same (all-ones) operand tiles; accumulators are N columns apart
        const long long dt = issue_loop<G>(tmem_base + (uint32_t)(warp * N), base, base + 16384u, idesc, groups,
                                           smem_u32(&bars[warp]), smem_u32(&bars[6]), mode == 2, mode != 3);
        out[warp] = (unsigned long long)dt;
    }
    __syncthreads();
    // wait until everything retired (bars[w] got `groups` arrivals; phases flip each time, so just poll the last commit)
    if (threadIdx.x == 0) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[7])) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bars[7])), "r"(0u) : "memory");
        out[4] = (unsigned long long)(clock64() - T0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int G>
void run(unsigned long long *d) {
    cudaFuncSetAttribute(k_probe<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int groups = 256;
    for (int mode : {0, 3, 2, 1, 4})
        for (int N : {64, 128, 256}) {
            if ((mode == 1 && N > 256) || (mode == 4 && N > 128)) continue;
            unsigned long long h[5];
            for (int rep = 0; rep < 2; ++rep) {
                k_probe<G><<<1, 128, 52 * 1024>>>(N, groups, mode, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
            }
            cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
            const double nm = (double)groups * G * (mode == 1 ? 2 : mode == 4 ? 4 : 1);
            printf("mode %d G %2d N %3d | issue cyc/MMA (thread 0) %6.1f | retire cyc/MMA (all) %6.1f | tensor floor %d\n", mode, G, N,
                   (double)h[0] / (groups * G), (double)h[4] / nm, 128 * N / 256);
        }
}

int main() {
    unsigned long long *d;
    cudaMalloc(&d, 64);
    run<1>(d); run<2>(d); run<4>(d); run<8>(d); run<16>(d);
    return 0;
}
