// tc_peak_probe.cu -- what the tensor pipe of this B200 delivers for tcgen05.mma kind::i8 (s8 x s8 -> s32) and kind::f16
// (bf16 x bf16 -> f32) when nothing but the MMAs runs: every SM issues back-to-back M = 128, N = 256 MMAs on operands that
// already sit in shared memory (no TMA, no epilogue).  This is the int8 roofline denominator SURVEY 8(d) asks for
// (MEASURED_PEAKS.json only has the bf16 cuBLAS figure); the bf16 line calibrates it against that file.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_peak_probe tc_peak_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KIND>   // 0: kind::f16 (K = 16 per MMA), 1: kind::i8 (K = 32 per MMA); both read 32 bytes per row per MMA
__global__ void __launch_bounds__(128) k_peak(int groups, uint32_t idesc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < (48 * 1024) / 4; i += 128) reinterpret_cast<uint32_t *>(smem + (base - smem_u32(smem)))[i] = KIND ? 0x01010101u : 0x3c003c00u;
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
    if (threadIdx.x == 0) {
        const uint64_t hi = (uint64_t)(((8u * 128u) >> 4) | (1u << 14) | (2u << 29)) << 32;
        const uint64_t a0 = hi | (uint64_t)(((base & 0x3FFFFu) >> 4) | (1u << 16));
        const uint64_t b0 = hi | (uint64_t)((((base + 16384u) & 0x3FFFFu) >> 4) | (1u << 16));
        for (int g = 0; g < groups; ++g) {
            const uint32_t d = tmem_base + (uint32_t)((g & 1) * 256);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint64_t ad = a0 + 2u * (uint32_t)(k & 3), bd = b0 + 2u * (uint32_t)(k & 3);
                if (KIND) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                  ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int KIND>
void run(const char *name, int sms) {
    cudaFuncSetAttribute(k_peak<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    // D f32 / s32; A, B formats; N = 256, M = 128
    const uint32_t idesc = KIND ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
    const int groups = 8192;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_peak<KIND><<<sms, 128, 52 * 1024>>>(groups, idesc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double kper = KIND ? 32.0 : 16.0;
    const double ops = 2.0 * 128 * 256 * kper * 16.0 * groups * sms;
    printf("%-28s %d SMs: %.3f ms  ->  %.1f T%s/s  (%.0f ops/clk/SM at 1.965 GHz)\n", name, sms, best, ops / best / 1e9, KIND ? "OP" : "FLOP",
           ops / (best * 1e-3) / sms / 1.965e9);
}

int main() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    run<0>("kind::f16 bf16 -> f32", sms);
    run<1>("kind::i8  s8 -> s32", sms);
    run<0>("kind::f16 bf16 -> f32 (again)", sms);
    return 0;
}
