// yb_conv_tc.cu -- FP32-variant convolution (reference yolov2_forward_network.c:204-261) as a
// persistent, warp-specialised implicit GEMM on the 5th-generation tensor cores of sm_100a:
//
//     D[pixel, filter] = sum_{tap, c} A[pixel + tap, c] * W[filter, (tap, c)]
//
//   * A (activations, bf16, padded NHWC) is never materialised as an im2col matrix: for every (tap, 64-channel)
//     K-block the TMA engine loads a [TH x TW pixels] x [BK channels] box straight out of the activation tensor,
//     shifted by the tap, into 128B-swizzled shared memory.  Out-of-image taps read the tensor's zero border
//     (or TMA's out-of-bounds zero fill at the ends of the batch), so there is no bounds logic anywhere.
//     Tiles are rectangles of TW x TH = 128 output pixels over (x, merged batch*row) so that the 19*2^k-wide
//     YOLO grids tile exactly.  Stride-2 convolutions use a 5-D view that splits x and y into (half, parity).
//   * W ([filters][K] bf16, K ordered (ky, kx, c)) is the K-major B operand, loaded by TMA as well.
//   * One elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> f32, M=128, N=BN<=256, K=16) with the
//     accumulator in TMEM; tcgen05.commit releases shared-memory stages / publishes the accumulator through
//     mbarriers.  Two TMEM accumulators let the epilogue of tile i overlap the main loop of tile i+1.
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b.x32), add bias (folded batch-norm), apply leaky-ReLU, add the
//     shortcut residual when fused (reference :443-449), and store bf16 (or f32 for detection heads) NHWC.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "yb_conv_tc.cuh"

namespace yb {

namespace {

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 192;
constexpr int TC_ACC = 2;   // TMEM accumulator stages

struct TcParams {
    int N;                    // images
    int TW, TWlog2, TH;       // tile = TW x TH output pixels (TW*TH == 128)
    int xt, jt, nt;           // #tiles along x, merged rows, filters
    int num_tiles;
    int PR, row_off;          // merged-row pitch per image; output row = (J % PR) - row_off
    int OH, OW, OHp, OWp;
    int size, cblocks, kblocks;
    int BK, BN;
    int stride2;
    int xoff, yoff;
    int stages;
    uint32_t stage_bytes, a_bytes, b_bytes;   // per stage (all sub-blocks); per K-block A tile; per K-block B tile
    int sps;                                  // K-blocks per pipeline stage (amortises the per-stage barrier round trip)
    uint32_t idesc, desc_hi;  // UMMA instruction descriptor; high word of the smem descriptors
    char *out; long out_ldc; int out_bf16; int n, n_store;
    const char *res; long res_ldc; int res_bf16;
    const float *bias; int act, act2;
    uint32_t tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// Bounded wait: a protocol bug must surface as a trap (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int what) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
            printf("yb k_conv_tc: mbarrier timeout (what=%d block=%d thread=%d parity=%u)\n", what, blockIdx.x,
                   threadIdx.x, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B swizzle atoms are 1024B aligned
    const uint32_t bars = smem0 + (uint32_t)p.stages * p.stage_bytes;
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(p.stages + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + TC_ACC + a); };
    const uint32_t tmem_slot = bars + 8u * (uint32_t)(2 * p.stages + 2 * TC_ACC);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < TC_ACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // bias (folded batch-norm) for all filter tiles -> shared memory, once per CTA
    float *bias_s = reinterpret_cast<float *>(smem_raw + (tmem_slot + 16u - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < p.nt * p.BN; i += TC_THREADS) bias_s[i] = (i < p.n) ? __ldg(p.bias + i) : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int n_idx = t % p.nt;
                const int m = t / p.nt;
                const int x0 = (m % p.xt) * p.TW;
                const int J0 = (m / p.xt) * p.TH;
                const int n0 = n_idx * p.BN;
                for (int kb0 = 0; kb0 < p.kblocks; kb0 += p.sps) {
                    const int nsub = min(p.sps, p.kblocks - kb0);
                    mbar_wait(empty_bar(stage), phase ^ 1u, 0);
                    const uint32_t a_dst = smem0 + (uint32_t)stage * p.stage_bytes;
                    const uint32_t b_dst = a_dst + (uint32_t)p.sps * p.a_bytes;
                    mbar_arrive_expect_tx(full_bar(stage), (uint32_t)nsub * (p.a_bytes + p.b_bytes));
                    for (int j = 0; j < nsub; ++j) {
                        const int kb = kb0 + j;
                        const int tap = kb / p.cblocks;
                        const int c0 = (kb - tap * p.cblocks) * p.BK;
                        const int ky = tap / p.size, kx = tap - ky * p.size;
                        const uint32_t ad = a_dst + (uint32_t)j * p.a_bytes, bd = b_dst + (uint32_t)j * p.b_bytes;
                        if (p.stride2) tma_load_5d(ad, &tmA, full_bar(stage), c0, kx & 1, x0 + (kx >> 1), ky & 1, J0 + (ky >> 1));
                        else tma_load_3d(ad, &tmA, full_bar(stage), c0, x0 + kx + p.xoff, J0 + ky + p.yoff);
                        tma_load_2d(bd, &tmB, full_bar(stage), kb * p.BK, n0);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const int kk = p.BK / 16;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 1);   // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                for (int kb0 = 0; kb0 < p.kblocks; kb0 += p.sps) {
                    const int nsub = min(p.sps, p.kblocks - kb0);
                    mbar_wait(full_bar(stage), phase, 2);        // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t a_base = smem0 + (uint32_t)stage * p.stage_bytes;
                    const uint32_t b_base = a_base + (uint32_t)p.sps * p.a_bytes;
                    const uint64_t hi = (uint64_t)p.desc_hi << 32;
                    for (int j = 0; j < nsub; ++j) {
                        const uint32_t a_addr = a_base + (uint32_t)j * p.a_bytes, b_addr = b_base + (uint32_t)j * p.b_bytes;
                        for (int k = 0; k < kk; ++k) {
                            // K-major operand, K advance of 16 bf16 = 32 bytes inside the swizzle row
                            const uint64_t adesc = hi | (uint64_t)((((a_addr + 32u * k) & 0x3FFFFu) >> 4) | (1u << 16));
                            const uint64_t bdesc = hi | (uint64_t)((((b_addr + 32u * k) & 0x3FFFFu) >> 4) | (1u << 16));
                            umma_bf16(d_tmem, adesc, bdesc, p.idesc, (uint32_t)((kb0 | j | k) != 0));
                        }
                    }
                    umma_commit(empty_bar(stage));               // frees the smem stage when these MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull_bar(acc));                     // accumulator complete -> epilogue
                if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ======================= epilogue (warps 2..5) =======================
        // Per 64-column slab: issue both TMEM loads and the residual loads first, wait once, then do the math and
        // the stores -- global-load latency is paid once per slab instead of once per value (the first version
        // was epilogue-bound on exactly that, profiles/r01_notes.md).
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;              // accumulator row == pixel within the tile
        const int tx = r & (p.TW - 1), ty = r >> p.TWlog2;
        const bool leaky = p.act == ACT_LEAKY, leaky2 = p.act2 == ACT_LEAKY;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            const int n_idx = t % p.nt;
            const int m = t / p.nt;
            const int ox = (m % p.xt) * p.TW + tx;
            const int J = (m / p.xt) * p.TH + ty;
            const int n0 = n_idx * p.BN;
            const int img = J / p.PR;
            const int oy = J - img * p.PR - p.row_off;
            const bool valid = (img < p.N) && (oy >= 0) && (oy < p.OH) && (ox < p.OW);
            const long pix = ((long)(img * p.OHp + oy + 1) * p.OWp + ox + 1);
            char *orow = p.out + pix * p.out_ldc * (p.out_bf16 ? 2 : 4);
            const char *rrow = (p.res && valid) ? p.res + pix * p.res_ldc * 2 : nullptr;
            const float *bs = bias_s + n0;

            // residual for the first slab can be fetched before the accumulator is ready
            uint4 rv[2][4];
            auto load_res = [&](int f0, uint4 (&dst)[4]) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    dst[g] = make_uint4(0u, 0u, 0u, 0u);
                    if (rrow && (n0 + f0 + g * 8) < p.n_store)
                        dst[g] = __ldg(reinterpret_cast<const uint4 *>(rrow + (size_t)(n0 + f0) * 2) + g);
                }
            };
            load_res(0, rv[0]);
            if (p.BN > 32) load_res(32, rv[1]);

            mbar_wait(tfull_bar(acc), acc_phase, 3);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);

            auto finish = [&](const uint32_t (&v)[32], const uint4 (&rr)[4], int f0) {
                if (!valid || (n0 + f0) >= p.n_store) return;
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float a = __uint_as_float(v[j]) + bs[f0 + j];
                    x[j] = leaky ? ((a > 0.f) ? a : 0.1f * a) : a;
                }
                if (p.res) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t w[4] = {rr[g].x, rr[g].y, rr[g].z, rr[g].w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            x[g * 8 + 2 * h] += __uint_as_float(w[h] << 16);
                            x[g * 8 + 2 * h + 1] += __uint_as_float(w[h] & 0xffff0000u);
                        }
                    }
                    if (leaky2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = (x[j] > 0.f) ? x[j] : 0.1f * x[j];
                    }
                }
                if (p.out_bf16) {
                    uint4 *op = reinterpret_cast<uint4 *>(orow + (size_t)(n0 + f0) * 2);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (n0 + f0 + g * 8 >= p.n_store) break;
                        uint4 o;
                        o.x = pack_bf16x2(x[g * 8 + 0], x[g * 8 + 1]);
                        o.y = pack_bf16x2(x[g * 8 + 2], x[g * 8 + 3]);
                        o.z = pack_bf16x2(x[g * 8 + 4], x[g * 8 + 5]);
                        o.w = pack_bf16x2(x[g * 8 + 6], x[g * 8 + 7]);
                        op[g] = o;
                    }
                } else {
                    float4 *op = reinterpret_cast<float4 *>(orow + (size_t)(n0 + f0) * 4);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (n0 + f0 + g * 4 >= p.n_store) break;
                        op[g] = make_float4(x[g * 4 + 0], x[g * 4 + 1], x[g * 4 + 2], x[g * 4 + 3]);
                    }
                }
            };

            if (p.BN == 32) {
                uint32_t v0[32];
                tmem_ld32(taddr, v0);
                tmem_ld_wait();
                finish(v0, rv[0], 0);
            } else {
                for (int f0 = 0; f0 < p.BN; f0 += 64) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + (uint32_t)f0, v0);
                    tmem_ld32(taddr + (uint32_t)f0 + 32u, v1);
                    tmem_ld_wait();
                    uint4 r0[4], r1[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) { r0[g] = rv[0][g]; r1[g] = rv[1][g]; }
                    if (f0 + 64 < p.BN) { load_res(f0 + 64, rv[0]); load_res(f0 + 96, rv[1]); }   // next slab in flight
                    finish(v0, r0, f0);
                    finish(v1, r1, f0 + 32);
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));   // 128 arrivals hand the accumulator back to the MMA warp
            if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
            fatal_throw("cuTensorMapEncodeTiled not available from the driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct TcPlan {
    CUtensorMap tmA, tmB;
    TcParams p;
    int grid;
    size_t smem;
};

int pick_bk(int C) { return (C % 64 == 0) ? 64 : (C % 32 == 0) ? 32 : (C % 16 == 0) ? 16 : 0; }
int pick_bn(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }

}  // namespace

int tc_conv_supported(const Layer &l, const TV &in, const TV &out, bool out_bf16) {
    if (!out.base || (reinterpret_cast<uintptr_t>(out.base) & 15) != 0) return 0;
    if (out_bf16 ? (out.ldc % 8 != 0) : (out.ldc % 4 != 0)) return 0;
    if (l.activation != YB_LEAKY && l.activation != YB_LINEAR) return 0;
    if (pick_bk(l.c) == 0) return 0;
    if (in.ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(in.base) & 15) != 0 || in.P != 1) return 0;
    const bool s1 = l.stride == 1 && ((l.size == 3 && l.pad == 1) || (l.size == 1 && l.pad == 0));
    const bool s2 = l.stride == 2 && l.size == 3 && l.pad == 1 && (l.h % 2 == 0) && (l.w % 2 == 0);
    if (!s1 && !s2) return 0;
    if (out_bf16 && l.n % 8 != 0) return 0;
    if (l.n < 8) return 0;
    return 1;
}

void *tc_make_plan(const Layer &l, const TV &in, const TV &out, bool out_bf16, const TV &res, bool res_bf16,
                   int act2, const void *d_weights_bf16, int ldn, const float *d_bias) {
    TcPlan *plan = new TcPlan();
    memset(plan, 0, sizeof(*plan));
    TcParams &p = plan->p;
    const int BK = pick_bk(l.c), BN = pick_bn(l.n);
    const bool s2 = l.stride == 2;
    p.N = in.N;
    p.OH = l.out_h; p.OW = l.out_w; p.OHp = out.Hp; p.OWp = out.Wp;
    p.size = l.size; p.BK = BK; p.BN = BN;
    p.cblocks = l.c / BK; p.kblocks = l.size * l.size * p.cblocks;
    p.stride2 = s2 ? 1 : 0;
    p.xoff = 1 - l.pad; p.yoff = -l.pad;
    p.PR = s2 ? (in.Hp / 2) : in.Hp;
    p.row_off = s2 ? 0 : 1;
    // tile width: power of two minimising padded work
    const long rows = (long)in.N * p.PR;
    double best = 1e30; int bestTW = 1;
    for (int tw = 1; tw <= 128; tw *= 2) {
        const int th = 128 / tw;
        const double cost = (double)((p.OW + tw - 1) / tw) * tw * (double)((rows + th - 1) / th) * th;
        if (cost < best - 0.5) { best = cost; bestTW = tw; }
    }
    p.TW = bestTW; p.TH = 128 / bestTW;
    p.TWlog2 = 0; while ((1 << p.TWlog2) < p.TW) ++p.TWlog2;
    p.xt = (p.OW + p.TW - 1) / p.TW;
    p.jt = (int)((rows + p.TH - 1) / p.TH);
    p.nt = (l.n + BN - 1) / BN;
    p.num_tiles = p.xt * p.jt * p.nt;
    p.a_bytes = (uint32_t)(TC_BM * BK * 2);
    p.b_bytes = (uint32_t)(BN * BK * 2);
    p.nt = (l.n + BN - 1) / BN;
    // several K-blocks per stage when they are small: the single MMA-issuing thread pays a fixed barrier round
    // trip per stage, which dominated the C<=64 layers (profiles/r01_notes.md)
    p.sps = (int)std::max<uint32_t>(1, std::min<uint32_t>(4, (64u * 1024u) / (p.a_bytes + p.b_bytes)));
    p.sps = std::min(p.sps, p.kblocks);
    p.stage_bytes = (uint32_t)p.sps * (p.a_bytes + p.b_bytes);
    p.stages = (int)std::min<size_t>(8, (200 * 1024 - sizeof(float) * (size_t)p.nt * BN) / p.stage_bytes);
    if (p.stages < 2) fatal_throw("tc plan: tile does not fit shared memory");
    // UMMA instruction descriptor (kind::f16): D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    // smem descriptor high word: SBO (8 rows * row bytes) >> 4 at bits 32..45, version 1 at bit 46, swizzle at 61..63
    const uint32_t row_bytes = (uint32_t)BK * 2;
    const uint32_t layout = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
    p.desc_hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout << 29);
    p.out = out.base; p.out_ldc = out.ldc; p.out_bf16 = out_bf16 ? 1 : 0;
    p.n = l.n;
    p.n_store = out_bf16 ? l.n : std::min<int>((l.n + 3) / 4 * 4, out.ldc);
    if (!out_bf16 && (out.ldc % 4 != 0)) fatal_throw("tc plan: f32 output rows must be 16-byte aligned");
    p.res = res.base; p.res_ldc = res.ldc; p.res_bf16 = res_bf16 ? 1 : 0;
    if (res.base && (res.H != l.out_h || res.W != l.out_w || res.C != l.n)) fatal_throw("tc plan: residual shape mismatch");
    if (res.base && !res_bf16) fatal_throw("tc plan: residual must be bf16");
    if (res.base && res_bf16 && (res.ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(res.base) & 15))) fatal_throw("tc plan: residual alignment");
    p.bias = d_bias; p.act = l.activation; p.act2 = act2;
    uint32_t cols = 32; while (cols < (uint32_t)(TC_ACC * BN)) cols *= 2;
    p.tmem_cols = cols;

    const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    EncodeTiledFn enc = encode_fn();
    CUresult r;
    if (!s2) {
        // activation view (c, x_padded, merged padded rows)
        cuuint64_t dims[3] = {(cuuint64_t)l.c, (cuuint64_t)in.Wp, (cuuint64_t)in.N * in.Hp};
        cuuint64_t strides[2] = {(cuuint64_t)in.ldc * 2, (cuuint64_t)in.Wp * in.ldc * 2};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)p.TW, (cuuint32_t)p.TH};
        cuuint32_t es[3] = {1, 1, 1};
        r = enc(&plan->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, in.base, dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // stride 2: (c, x parity, x half, y parity, merged y half)
        cuuint64_t dims[5] = {(cuuint64_t)l.c, 2, (cuuint64_t)in.Wp / 2, 2, (cuuint64_t)in.N * in.Hp / 2};
        cuuint64_t strides[4] = {(cuuint64_t)in.ldc * 2, (cuuint64_t)in.ldc * 4, (cuuint64_t)in.Wp * in.ldc * 2,
                                 (cuuint64_t)in.Wp * in.ldc * 4};
        cuuint32_t box[5] = {(cuuint32_t)BK, 1, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        r = enc(&plan->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, in.base, dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { delete plan; fatal_throw("cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r)); }
    {
        const cuuint64_t K = (cuuint64_t)l.size * l.size * l.c;
        cuuint64_t dims[2] = {K, (cuuint64_t)ldn};
        cuuint64_t strides[1] = {K * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        cuuint32_t es[2] = {1, 1};
        r = enc(&plan->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d_weights_bf16), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete plan; fatal_throw("cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r)); }
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    plan->grid = std::min(p.num_tiles, sms);
    plan->smem = (size_t)p.stages * p.stage_bytes + 1024 /*alignment slack*/ + 8 * (2 * p.stages + 2 * TC_ACC) + 16 +
                 sizeof(float) * (size_t)p.nt * BN /*bias*/;
    if (cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        fatal_throw("cudaFuncSetAttribute(k_conv_tc) failed");
    return plan;
}

void tc_launch(void *vp, cudaStream_t s) {
    TcPlan *plan = reinterpret_cast<TcPlan *>(vp);
    k_conv_tc<<<plan->grid, TC_THREADS, plan->smem, s>>>(plan->tmA, plan->tmB, plan->p);
}

void tc_free_plan(void *vp) { delete reinterpret_cast<TcPlan *>(vp); }

}  // namespace yb
