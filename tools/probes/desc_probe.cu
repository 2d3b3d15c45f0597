// desc_probe.cu -- what does tcgen05.mma do when the A operand's shared-memory descriptor starts at a row that is NOT
// the first row of a swizzle atom, and when the 8-row groups are not a whole number of atoms apart?
//
// Background (profiles/r02_notes.md): a 3x3 convolution re-loads every activation pixel nine times through TMA (one
// box per tap).  If the tile is loaded ONCE with its halo -- [TH+2][TW+2] pixels x 128 B, 128B-swizzled by TMA on
// absolute shared-memory address bits -- tap (ky, kx) is the same tile read from row (ky*(TW+2) + kx) on, with the
// 8-row groups (TW = 8) (TW+2) rows apart: start address = base + r0*128 (any r0), SBO = (TW+2)*128.  The descriptor has
// a 3-bit "base offset" field for start addresses that are not 1024-byte aligned; this probe finds out which
// combination (base offset = 0 / = (addr >> 7) & 7) reproduces  D[m][n] = sum_k A[row(m)][k] * B[n][k],
// row(m) = r0 + (m / 8) * (SBO / rowbytes) + m % 8, for the 128B, 64B and 32B swizzles.
//
// Also checks that a TMA box whose row count is not a multiple of 8 ({64 ch, 10, 18}) lands densely with the swizzle
// a function of the absolute address only.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o desc_probe desc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Case { int swz_bytes; int r0; int sbo_rows; int use_base_off; int kchunks; };

// A_log: [rows][K] bf16 logical (K = swz_bytes / 2), B_log: [64][K].  The kernel writes them into shared memory the way TMA
// would (16-byte chunk c of row r at r*rowbytes + ((c ^ f(r)) << 4), f from the ABSOLUTE address), runs one MMA group and
// returns D [128][64] f32.
__global__ void __launch_bounds__(128) k_desc(const __nv_bfloat16 *A_log, int a_rows, const __nv_bfloat16 *B_log, Case cs, float *D) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (base - smem_u32(smem_raw));
    const int rb = cs.swz_bytes;                 // row bytes
    const int cpr = rb / 16;                     // 16-byte chunks per row
    const uint32_t a_off = 0, b_off = 48 * 1024;
    auto swz = [&](uint32_t byte_addr) {         // absolute-address swizzle: XOR bits [4, 4+log2(cpr)) with bits [7, ...)
        const uint32_t mask = (uint32_t)(cpr - 1);
        return byte_addr ^ (((byte_addr >> 7) & mask) << 4);
    };
    for (int i = threadIdx.x; i < a_rows * cpr; i += 128) {
        const int r = i / cpr, c = i % cpr;
        const uint4 v = *reinterpret_cast<const uint4 *>(A_log + (size_t)r * (rb / 2) + c * 8);
        *reinterpret_cast<uint4 *>(sm + (swz(base + a_off + (uint32_t)(r * rb + c * 16)) - base)) = v;
    }
    for (int i = threadIdx.x; i < 64 * cpr; i += 128) {
        const int r = i / cpr, c = i % cpr;
        const uint4 v = *reinterpret_cast<const uint4 *>(B_log + (size_t)r * (rb / 2) + c * 8);
        *reinterpret_cast<uint4 *>(sm + (swz(base + b_off + (uint32_t)(r * rb + c * 16)) - base)) = v;
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t layout = rb == 128 ? 2u : rb == 64 ? 4u : 6u;
        const uint32_t a_start = base + a_off + (uint32_t)(cs.r0 * rb);
        const uint32_t b_start = base + b_off;
        const uint32_t boff = cs.use_base_off ? ((a_start >> 7) & 7u) : 0u;
        const uint64_t hiA = (uint64_t)((((uint32_t)(cs.sbo_rows * rb)) >> 4) | (1u << 14) | (boff << 17) | (layout << 29)) << 32;
        const uint64_t hiB = (uint64_t)(((8u * (uint32_t)rb) >> 4) | (1u << 14) | (layout << 29)) << 32;
        // idesc: D f32, A = B = bf16, K-major both, N = 64, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < cs.kchunks; ++k) {
            const uint64_t ad = hiA | (uint64_t)((((a_start + 32u * k) & 0x3FFFFu) >> 4) | (1u << 16));
            const uint64_t bd = hiB | (uint64_t)((((b_start + 32u * k) & 0x3FFFFu) >> 4) | (1u << 16));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_base), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(k != 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t ok = 0; long long t0 = clock64();
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("desc_probe: timeout\n"); __trap(); }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) D[(size_t)threadIdx.x * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

// ---- TMA landing check: box {64 ch, 10, 18} of a [rows][cols][64] bf16 tensor -> shared memory, dumped raw ------------
__global__ void __launch_bounds__(32) k_tma(const __grid_constant__ CUtensorMap tm, uint8_t *dump, int bytes, int x0, int y0) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *sm = smem_raw + (base - smem_u32(smem_raw));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(base), "l"(&tm), "r"(smem_u32(&bar)), "r"(0), "r"(x0), "r"(y0) : "memory");
    }
    __syncwarp();
    uint32_t ok = 0; long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("k_tma: timeout\n"); __trap(); }
    }
    for (int i = threadIdx.x; i < bytes; i += 32) dump[i] = sm[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    // ---------------- part 1: descriptors -------------------------------------------------------------------------
    const int a_rows = 320;
    printf("# swizzle r0 sbo_rows base_off | max|D - expected| (0 = the hardware reads rows r0 + (m/8)*sbo_rows + m%%8)\n");
    cudaFuncSetAttribute(k_desc, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 1024);
    for (int rb : {128, 64, 32}) {
        const int K = rb / 2;
        std::vector<__nv_bfloat16> A((size_t)a_rows * K), B((size_t)64 * K);
        std::vector<float> Af(A.size()), Bf(B.size());
        uint32_t st = 12345u + rb;
        auto rnd = [&]() { st = st * 1664525u + 1013904223u; return (float)((int)((st >> 16) % 7) - 3); };
        for (size_t i = 0; i < A.size(); ++i) { Af[i] = rnd(); A[i] = __float2bfloat16(Af[i]); }
        for (size_t i = 0; i < B.size(); ++i) { Bf[i] = rnd(); B[i] = __float2bfloat16(Bf[i]); }
        __nv_bfloat16 *dA, *dB; float *dD;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
        for (int sbo_rows : {8, 10, 12, 16, 20}) {
            for (int r0 : {0, 1, 2, 3, 4, 5, 7, 8, 10, 11, 20, 21}) {
                if (r0 + 15 * sbo_rows + 8 > a_rows) continue;
                for (int ub : {0, 1}) {
                    Case cs{rb, r0, sbo_rows, ub, K / 16};
                    cudaMemset(dD, 0, 128 * 64 * 4);
                    k_desc<<<1, 128, 64 * 1024 + 1024>>>(dA, a_rows, dB, cs, dD);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("swz%d r0 %d sbo %d bo %d : CUDA error %s\n", rb, r0, sbo_rows, ub, cudaGetErrorString(e)); return 1; }
                    std::vector<float> D(128 * 64);
                    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                    double maxerr = 0;
                    for (int m = 0; m < 128; ++m) {
                        const int row = r0 + (m / 8) * sbo_rows + m % 8;
                        for (int n = 0; n < 64; ++n) {
                            float ref = 0;
                            for (int k = 0; k < K; ++k) ref += Af[(size_t)row * K + k] * Bf[(size_t)n * K + k];
                            const double d = fabs((double)D[m * 64 + n] - ref);
                            if (d > maxerr) maxerr = d;
                        }
                    }
                    printf("swz%-3d r0 %2d sbo_rows %2d base_off %d | %g %s\n", rb, r0, sbo_rows, ub, maxerr, maxerr == 0 ? "OK" : "MISMATCH");
                }
            }
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    // ---------------- part 2: TMA box with 10 x 18 rows ---------------------------------------------------------------
    {
        void *fnp = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
        EncodeTiledFn enc = (EncodeTiledFn)fnp;
        const int W = 40, H = 40, C = 64;
        std::vector<__nv_bfloat16> T((size_t)H * W * C);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c)
            T[((size_t)y * W + x) * C + c] = __float2bfloat16((float)((y * 64 + x) % 251) + (c == 0 ? 0.f : 0.f) + (float)(c % 4) * 0.25f);
        __nv_bfloat16 *dT; cudaMalloc(&dT, T.size() * 2); cudaMemcpy(dT, T.data(), T.size() * 2, cudaMemcpyHostToDevice);
        for (int bw : {10, 8}) {
            const int bh = 18;
            CUtensorMap tm;
            cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
            cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
            cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dT, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("tma box {64,%d,%d}: encode failed %d\n", bw, bh, (int)r); continue; }
            const int bytes = bw * bh * 128;
            uint8_t *dd; cudaMalloc(&dd, bytes);
            cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            const int x0 = 3, y0 = 5;
            k_tma<<<1, 32, 48 * 1024>>>(tm, dd, bytes, x0, y0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("tma box {64,%d,%d}: CUDA error %s\n", bw, bh, cudaGetErrorString(e)); return 1; }
            std::vector<uint8_t> hd(bytes); cudaMemcpy(hd.data(), dd, bytes, cudaMemcpyDeviceToHost);
            long bad = 0;
            for (int ry = 0; ry < bh; ++ry) for (int rx = 0; rx < bw; ++rx) for (int ch = 0; ch < 8; ++ch) {
                const int line = ry * bw + rx;
                const uint32_t off = (uint32_t)(line * 128 + ((ch ^ (line & 7)) << 4));   // dense lines, absolute-address swizzle
                const __nv_bfloat16 *src = &T[((size_t)(y0 + ry) * W + (x0 + rx)) * C + ch * 8];
                if (memcmp(&hd[off], src, 16) != 0) ++bad;
            }
            printf("tma box {64,%2d,%d} 128B swizzle: %ld of %d chunks differ from [dense lines, chunk ^ (line & 7)] %s\n", bw, bh, bad, bw * bh * 8, bad ? "MISMATCH" : "OK");
            cudaFree(dd);
        }
        cudaFree(dT);
    }
    return 0;
}
