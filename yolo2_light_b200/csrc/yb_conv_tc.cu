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
//   * 8 epilogue warps read TMEM (tcgen05.ld 32x32b.x32), add bias (folded batch-norm), apply leaky-ReLU, add the
//     shortcut residual when fused (reference :443-449), and store bf16 NHWC through a swizzled staging tile (whole
//     128-byte lines), or f32 for detection heads -- optionally with the following [yolo] layer applied (:453-472).
//   * The same kernel runs the INT8 variant (kind::i8, exact requantising epilogue, yolov2_forward_network_quantized.c:
//     474-490), wide XNOR layers as +-1 bytes on kind::i8, and the float heads of the exact networks on kind::tf32.
//   * CG = 2: CTA pairs (cta_group::2) for the BN = 256 layers; KS = true: K-split of the tail wave (opt-in).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
// (two warps per TMEM lane quarter, each owning half of the accumulator columns).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "yb_conv_tc.cuh"

namespace yb {

namespace {

constexpr int TC_BM = 128;
constexpr int TC_EPI_WARPS = 8;              // two warps per TMEM lane quarter, each takes half of the columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_ACC = 2;   // TMEM accumulator stages
constexpr int TC_MAX_ASTAGES = 4;   // halo mode: A-ring depth (barriers are always reserved)

struct TcParams {
    int N;                    // images
    int TW, TWlog2, TH;       // tile = TW x TH output pixels (TW*TH == 128)
    int xt, jt, nt;           // #tiles along x, merged rows, filters
    int num_tiles;
    int num_work;             // work items of the persistent loop: num_tiles (CG=1) or pairs of m-tiles x nt (CG=2)
    int cg;                   // 1, or 2 = CTA pairs (cta_group::2)
    int kind;                 // 0: bf16 x bf16 -> f32 (kind::f16);  1: s8 x s8 -> s32 (kind::i8), exact requantising epilogue;
                              // 2: XNOR layer as +-1 s8 on kind::i8 (dot = 2*count - K exactly), reference float epilogue;
                              // 3: f32 operands read as tf32 (kind::tf32, K = 8 per MMA) -> f32: float heads of the exact nets
    int kk;                   // MMAs per K-block (BK bytes / 32)
    float alpha1;             // INT8: R_MULT / (input_mult * weights_mult)
    const float *mean;        // kind 2 (XNOR as +-1 s8): per-filter mean |w|; out = (float)dot * mean + bias
    int xK;                   // kind 2: true K (size*size*C) for the raw popcount dump: count = (dot + K) / 2
    int *acc_out;             // INT8: optional raw s32 accumulators, NCHW (tests)
    float *yolo_out;          // fused [yolo] layer (reference yolov2_forward_network.c:453-472): NCHW f32 destination, or null
    int yolo_per;             // 4 + classes + 1
    int PR, row_off;          // merged-row pitch per image; output row = (J % PR) - row_off
    int OH, OW, OHp, OWp;
    int size, cblocks, kblocks;
    int BK, BN;
    int stride2;
    int xoff, yoff;
    int stages;
    uint32_t stage_bytes, a_bytes, b_bytes;   // per stage (all sub-blocks); per K-block A tile; per K-block B tile
    int bstat;                                // 1: the whole filter matrix (nt == 1, <= 72 KB) is loaded once per CTA and stays in
                                              //    shared memory; the ring then streams activations only
    uint32_t bstat_bytes;
    // Halo mode (3x3 / stride 1 / pad 1): the activation tile is loaded ONCE per channel block with its 1-pixel halo -- a
    // [TH+2][TW+2] x BK box, TW = 8 -- and the nine taps are the same shared-memory tile read from line ky*(TW+2)+kx on, with
    // the 8-row groups (TW+2) lines apart (descriptor SBO).  tcgen05 applies the swizzle to absolute address bits, so any start
    // line and any SBO work (tools/probes/desc_probe.cu, profiles/r02_desc_probe.txt).  TMA bytes of A per K-block: 1/6.4 of
    // the one-box-per-tap scheme.  A ring: a_stages x a_stage_bytes; the B ring keeps `stages` x b_bytes.
    // halo == 2 (3x3 / stride 2 / pad 1, C <= 64): the same idea on the four parity planes of the padded input -- output (ox, oy)
    // reads padded pixel (2 ox + kx, 2 oy + ky), i.e. half-pixel (ox + (kx >> 1), oy + (ky >> 1)) of plane (ky & 1, kx & 1).  Four
    // [TH+1][TW+1] x BK boxes per channel block (one per plane, a_stage_bytes / 4 apart) replace nine [TH][TW] boxes: 0.53 of
    // the TMA bytes, which is what bound these layers (producer 47 % in TMA issue, MMA thread 34-50 % waiting for data).
    int halo, a_stages;
    uint32_t a_stage_bytes, halo_bytes, halo_pitch;   // halo_pitch = (TW+2) * row bytes (halo == 2: (TW+1) * row bytes)
    uint32_t desc_hi_a;                       // descriptor high word of the halo A operand (SBO = halo_pitch)
    // TMA epilogue (bf16 output, stride 1, BN >= 128): each group of four epilogue warps writes its 128 x 64-column slab as
    // bf16 into a 128B-swizzled shared-memory tile and one thread stores it with cp.async.bulk.tensor; the shortcut residual
    // comes in the same way (TMA load + mbarrier).  No shuffles, no staging transposes, no LSU global traffic, ~100 registers.
    int tma_epi;
    int l2_hint;              // 1: halo activation tiles and shortcut tiles are loaded with the L2 evict-first policy
    // epi_alt (with tma_epi, BN <= 128): the two groups of four epilogue warps take ALTERNATE tiles (group g: accumulator g, all
    // BN columns) instead of half the columns of every tile -- two tiles are in the epilogue at once; the per-tile epilogue of the
    // small tiles is a latency chain (TMEM load -> math -> barrier -> store), not a throughput problem.
    int epi_alt;
    // epi_bufs (TMA epilogue): OUT tiles per warp group.  With one tile a slab cannot be written before the bulk store of the previous
    // slab has finished READING the tile, which put the store's issue-to-read latency on the critical path of every slab of the
    // shallow-K layers; with two the group only waits for the store before the last one.
    int epi_bufs;
    // Fused 2x2 / stride-2 max-pool + input conversion of the NEXT integer layer (integer kinds, halo tiles of 8 x 16 pixels):
    // the epilogue reduces every 2x2 window inside the warp (lane ^ 1 = x neighbour, lane ^ 8 = y neighbour), quantises (pool_mode
    // 1: quant_i8 with pool_mult) or takes the sign (pool_mode 2: +-1 bytes) and writes bytes straight into the next layer's s8
    // input: the f32 activation and the pooled f32 tensor never reach HBM.  jshift = 1 moves every tile down one merged row so
    // that window rows (oy even, oy + 1) fall into the same tile (the padded layout puts oy = 0 on an odd merged row).
    int pool_mode, jshift;
    float pool_mult;
    signed char *pool_out; long pool_ldc; int pool_Hp, pool_Wp;   // next layer's s8 input: padded NHWC, bytes
    int sps;                                  // K-blocks per pipeline stage (amortises the per-stage barrier round trip)
    int kbs;                                  // pipeline stages per work item = ceil(kblocks / sps)
    // K-split tail (wave quantisation): the last num_work % G work items ("tail") are cut along K into slices of sk_L
    // stages, one slice per CTA (pair); a slice that does not end its work item dumps the raw f32 accumulator to
    // sk_ws, the slice that does (the owner) adds those partials in its epilogue.  sk_T == 0: off.
    int sk_T, sk_L;
    float *sk_ws;                             // [grid][BN/4][128] float4
    unsigned *sk_flags;                       // [grid][8 epilogue warps]: 1 = partial published (reset by its reader)
    uint32_t idesc, desc_hi;  // UMMA instruction descriptor; high word of the smem descriptors
    char *out; long out_ldc; int out_bf16; int n, n_store;
    const char *res; long res_ldc; int res_bf16;
    const float *bias; int act, act2;
    uint32_t tmem_cols;
    unsigned long long *stats; // YB_TC_STATS=1: per-CTA cycle counters [grid][16] (diagnostic)
    int no_coalesce;          // YB_TC_NO_COALESCE=1: per-thread row stores (the pre-staging epilogue), for A/B comparison
    int dbg;                  // YB_TC_DBG bit mask for bottleneck experiments: 1 no TMA, 2 no MMA, 4 no epilogue memory ops
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
// L2 eviction hint for data that is dead after this kernel (the C/2 tensor a 3x3 layer reads, the shortcut operand): evict-first
// leaves the L2 to the output this kernel writes, which the next layer reads right away
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
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

__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// --- cluster helpers (CG == 2: a CTA pair, tcgen05 cta_group::2) --------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {   // arrive on the same barrier of CTA `cta`
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-pair rank bit of a shared::cluster address -> leader CTA
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2_load_3d_hint(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int c0, int c1, int c2,
                                             int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// Work schedule of one CTA (CG=1) / CTA pair (CG=2), identical in every warp role: first this unit's slice of the
// K-split tail (units of pipeline stages over the first sk_T work items), then whole work items round-robin.
struct TcSched { int u, u_end, w_dp, w_step, num_work, kbs, lead0, lead1; };
template <bool KS>
__device__ __forceinline__ TcSched sched_init(const TcParams &p, int unit, int nunits) {
    TcSched s;
    s.kbs = p.kbs; s.num_work = p.num_work; s.w_step = nunits;
    if constexpr (!KS) { s.u = s.u_end = s.lead0 = s.lead1 = 0; s.w_dp = unit; return s; }
    const int U = p.sk_T * p.kbs;
    s.u = min(unit * p.sk_L, U); s.u_end = min(s.u + p.sk_L, U);
    s.w_dp = p.sk_T + unit;
    // A slice whose last segment stops short of its work item's end only PUBLISHES a partial sum; it goes first, so
    // that no partial ever waits behind a segment that itself waits for partials (which would chain the CTAs up).
    s.lead0 = s.lead1 = 0;
    if (s.u < s.u_end) {
        const int wl = (s.u_end - 1) / s.kbs;
        if (s.u_end < (wl + 1) * s.kbs) { s.lead0 = max(s.u, wl * s.kbs); s.lead1 = s.u_end; s.u_end = s.lead0; }
    }
    return s;
}
// next segment: work item w, stages [s0, s1) of its kbs stages
template <bool KS>
__device__ __forceinline__ bool sched_next(TcSched &s, int &w, int &s0, int &s1) {
    if constexpr (!KS) {
        if (s.w_dp >= s.num_work) return false;
        w = s.w_dp; s.w_dp += s.w_step; s0 = 0; s1 = s.kbs;
        return true;
    }
    if (s.lead0 < s.lead1) {
        w = s.lead0 / s.kbs; s0 = s.lead0 - w * s.kbs; s1 = s.lead1 - w * s.kbs; s.lead1 = s.lead0;
        return true;
    }
    if (s.u < s.u_end) {
        w = s.u / s.kbs; s0 = s.u - w * s.kbs; s1 = min(s.kbs, s0 + (s.u_end - s.u)); s.u += s1 - s0;
        return true;
    }
    if (s.w_dp >= s.num_work) return false;
    w = s.w_dp; s.w_dp += s.w_step; s0 = 0; s1 = s.kbs;
    return true;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// CG = 1: one CTA per 128-pixel tile.  CG = 2: a CTA pair computes a 256-pixel x BN tile with cta_group::2 MMAs --
// each CTA loads its own 128 pixels of A and HALF of the B (filter) tile, so the bytes every SM pulls through its
// TMA unit per FLOP drop by a third; that unit (~64 B/clk/SM) is what bounds the BN=256 layers
// (profiles/r01_notes.md).  The leader CTA (cluster rank 0) issues the MMAs for both.
// KS: compiled with the K-split tail schedule (TcParams::sk_T); the KS = false instantiations carry none of its code.
// ST: compiled with the per-role cycle counters of YB_TC_STATS=1 (diagnostic); the production instantiations (ST = false)
// contain no clock64() reads -- the single-thread producer / MMA roles are issue-bound on the BN <= 128 layers.
// EPI: which epilogue family is compiled in -- 0: LSU stores, float kinds (bf16 / f32 heads / fused [yolo]); 1: TMA epilogue
// (bf16 tiles stored with cp.async.bulk.tensor); 2: the integer kinds (kind::i8 requantising and XNOR-as-+-1 epilogues).  One
// kernel with all three spilled ~1.5 KB per thread (168 registers is the cap for 320 threads) and cost the LSU layers 10-25 %.
template <int CG, bool KS, bool ST, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
          const __grid_constant__ CUtensorMap tmR, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smemB = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B swizzle atoms are 1024B aligned
    const uint32_t smemA = smemB + p.bstat_bytes;                   // [resident filter matrix][halo A ring][pipeline ring]
    const uint32_t smem0 = smemA + (uint32_t)p.a_stages * p.a_stage_bytes;
    const uint32_t bars = smem0 + (uint32_t)p.stages * p.stage_bytes;
    auto full_bar = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bars + 8u * (uint32_t)(p.stages + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (uint32_t)(2 * p.stages + TC_ACC + a); };
    const uint32_t bstat_bar = bars + 8u * (uint32_t)(2 * p.stages + 2 * TC_ACC);
    auto fullA_bar = [&](int s) { return bstat_bar + 8u + 8u * (uint32_t)s; };          // halo mode: the A ring's own barriers
    auto emptyA_bar = [&](int s) { return bstat_bar + 8u + 8u * (uint32_t)(TC_MAX_ASTAGES + s); };
    auto resfull_bar = [&](int g) { return bstat_bar + 8u + 8u * (uint32_t)(2 * TC_MAX_ASTAGES + g); };   // TMA epilogue: residual landed
    const uint32_t tmem_slot = bstat_bar + 8u + 8u * (uint32_t)(2 * TC_MAX_ASTAGES + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    // work items: CG=1 -> (m_tile, n_tile); CG=2 -> (pair of m_tiles, n_tile), both CTAs of a pair iterate in lockstep
    const int w_first = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        // full: one arrival (the leader's expect_tx; the peer's bytes are covered by the transaction count);
        // tempty: one arrival per epilogue warp (of both CTAs when paired)
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < TC_ACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), CG * (p.epi_alt ? TC_EPI_WARPS / 2 : TC_EPI_WARPS)); }
        mbar_init(bstat_bar, 1);
        for (int s = 0; s < TC_MAX_ASTAGES; ++s) { mbar_init(fullA_bar(s), 1); mbar_init(emptyA_bar(s), 1); }
        mbar_init(resfull_bar(0), 1); mbar_init(resfull_bar(1), 1);
        if (p.tma_epi) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
            if (p.res && EPI == 1) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    // bias (folded batch-norm) for all filter tiles -> shared memory, once per CTA
    float *bias_s = reinterpret_cast<float *>(smem_raw + (tmem_slot + 16u - smem_u32(smem_raw)));
    // 4 KB of staging per epilogue warp (32 rows x 128 B, XOR-swizzled) for the coalescing transposes
    // fused [yolo]: one bit per filter, set where the entry is a box width/height (no logistic)
    uint32_t *ymask_s = reinterpret_cast<uint32_t *>(bias_s + p.nt * p.BN);
    const uint32_t stg_align = p.tma_epi ? 1023u : 127u;     // TMA epilogue tiles are 128B-swizzled: 1024-byte aligned
    const uint32_t stg_base = ((tmem_slot + 16u + 4u * (uint32_t)(p.nt * p.BN) + (uint32_t)(p.nt * p.BN / 8)) + stg_align) & ~stg_align;
    for (int i = threadIdx.x; i < p.nt * p.BN; i += TC_THREADS) bias_s[i] = (i < p.n) ? __ldg(p.bias + i) : 0.f;
    if (p.yolo_out) {
        for (int wd = threadIdx.x; wd < p.nt * p.BN / 32; wd += TC_THREADS) {
            uint32_t m = 0;
            for (int j = 0; j < 32; ++j) { const int e = (wd * 32 + j) % p.yolo_per; if (e == 2 || e == 3) m |= 1u << j; }
            ymask_s[wd] = m;
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch, bias ->
    // smem: weights only) overlapped the tail of the previous kernel; from here on we touch activations it wrote.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0) {
        // ======================= TMA producer (every CTA loads its own A rows and its share of B) ===========
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            long long w_empty = 0, w_tma = 0; const long long t_begin = ST ? clock64() : 0;
            // loop-invariant parameters in registers; the (tap, channel-block) walk is incremental -- the first version
            // recomputed it with two integer divisions per K-block, and that ~700-cycle dependent scalar chain in
            // this single thread was what starved the tensor pipe (profiles/r01_notes.md)
            const int sps = p.sps, kblocks = p.kblocks, cblocks = p.cblocks, BK = p.BK, fsize = p.size, stages = p.stages;
            const int xoff = p.xoff, yoff = p.yoff, stride2 = p.stride2, nt = p.nt, xt = p.xt;
            const uint32_t a_bytes = p.a_bytes, b_bytes = p.b_bytes, stage_bytes = p.stage_bytes;
            const uint32_t b_off = (uint32_t)sps * a_bytes;
            const int bstat = p.bstat;
            if (bstat) {   // resident filter matrix: kblocks boxes of [BN filters][BK], once
                mbar_arrive_expect_tx(bstat_bar, p.bstat_bytes);
                for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(smemB + (uint32_t)kb * b_bytes, &tmB, bstat_bar, kb * BK, 0);
            }
            if (p.halo) {
                // ---- halo mode: unit = (work item, channel block); the A tile of unit u+1 is requested before the nine filter
                // tiles of unit u, so the activation ring runs one unit ahead of the MMAs
                const int SA = p.a_stages;
                const uint32_t a_stage_bytes = p.a_stage_bytes, halo_bytes = p.halo_bytes;
                int sa = 0; uint32_t pha = 0;
                const uint64_t pol_first = l2_policy_evict_first();
                TcSched schA = sched_init<false>(p, w_first, w_step);
                int wA = 0, cbA = cblocks, d0, d1;
                auto next_A = [&]() {
                    if (cbA == cblocks) { if (!sched_next<false>(schA, wA, d0, d1)) return; cbA = 0; }
                    const int m = (CG == 2) ? 2 * (wA / nt) + (int)rank : wA / nt;
                    const int x0 = (m % xt) * p.TW, J0 = (m / xt) * p.TH + p.jshift;
                    if constexpr (ST) { const long long c0 = clock64(); mbar_wait(emptyA_bar(sa), pha ^ 1u, 5); w_tma += clock64() - c0; }   // [7]: wait on the A ring
                    else mbar_wait(emptyA_bar(sa), pha ^ 1u, 5);
                    const uint32_t fb = fullA_bar(sa);
                    if (leader) mbar_arrive_expect_tx(fb, (uint32_t)CG * halo_bytes);
                    const uint32_t ad = smemA + (uint32_t)sa * a_stage_bytes;
                    if (p.halo == 2) {
                        const uint32_t plane = a_stage_bytes >> 2;
#pragma unroll
                        for (int pl = 0; pl < 4; ++pl) {
                            if constexpr (CG == 2) tma2_load_5d(ad + (uint32_t)pl * plane, &tmA, fb, cbA * BK, pl & 1, x0, pl >> 1, J0);
                            else tma_load_5d(ad + (uint32_t)pl * plane, &tmA, fb, cbA * BK, pl & 1, x0, pl >> 1, J0);
                        }
                    } else if (p.l2_hint) {
                        if constexpr (CG == 2) tma2_load_3d_hint(ad, &tmA, fb, cbA * BK, x0, J0 - 1, pol_first);
                        else tma_load_3d_hint(ad, &tmA, fb, cbA * BK, x0, J0 - 1, pol_first);
                    } else if constexpr (CG == 2) tma2_load_3d(ad, &tmA, fb, cbA * BK, x0, J0 - 1);
                    else tma_load_3d(ad, &tmA, fb, cbA * BK, x0, J0 - 1);
                    ++cbA;
                    if (++sa == SA) { sa = 0; pha ^= 1u; }
                };
                next_A();
                TcSched sch = sched_init<false>(p, w_first, w_step);
                int w, seg0, seg1;
                while (sched_next<false>(sch, w, seg0, seg1)) {
                    const int n0 = (w % nt) * p.BN + (int)rank * (p.BN / CG);
                    for (int cb = 0; cb < cblocks; ++cb) {
                        next_A();
                        if (bstat) continue;
                        int kcol = cb * BK;                      // K is ordered (tap, channel): tap t of this block at t*C + cb*BK
                        for (int t = 0; t < 9; ++t, kcol += cblocks * BK) {
                            if constexpr (ST) { const long long c0 = clock64(); mbar_wait(empty_bar(stage), phase ^ 1u, 0); w_empty += clock64() - c0; }
                            else mbar_wait(empty_bar(stage), phase ^ 1u, 0);
                            const uint32_t fb = full_bar(stage);
                            if (leader) mbar_arrive_expect_tx(fb, (uint32_t)CG * b_bytes);
                            const uint32_t bd = smem0 + (uint32_t)stage * stage_bytes;
                            if constexpr (CG == 2) tma2_load_2d(bd, &tmB, fb, kcol, n0);
                            else tma_load_2d(bd, &tmB, fb, kcol, n0);
                            if (++stage == stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            } else {
            TcSched sch = sched_init<KS>(p, w_first, w_step);
            int w, seg0, seg1;
            while (sched_next<KS>(sch, w, seg0, seg1)) {
                const int n_idx = w % nt;
                const int m = (CG == 2) ? 2 * (w / nt) + (int)rank : w / nt;
                const int x0 = (m % xt) * p.TW;
                const int J0 = (m / xt) * p.TH + p.jshift;
                const int n0 = n_idx * p.BN + (int)rank * (p.BN / CG);
                const int kb_begin = seg0 * sps, kb_end = min(kblocks, seg1 * sps);
                // channel block, tap x/y, K column of the weight matrix at the first K-block of the segment
                const int tap0 = kb_begin / cblocks;
                int cb = kb_begin - tap0 * cblocks, ky = tap0 / fsize, kx = tap0 - (tap0 / fsize) * fsize, kcol = kb_begin * BK;
                for (int kb0 = kb_begin; kb0 < kb_end; kb0 += sps) {
                    const int nsub = min(sps, kb_end - kb0);
                    if constexpr (ST) { const long long c0 = clock64(); mbar_wait(empty_bar(stage), phase ^ 1u, 0); w_empty += clock64() - c0; }
                    else mbar_wait(empty_bar(stage), phase ^ 1u, 0);
                    const uint32_t fb = full_bar(stage);
                    const uint32_t a_dst = smem0 + (uint32_t)stage * stage_bytes;
                    const uint32_t b_dst = a_dst + b_off;
                    if (p.dbg & 1) {
                        if (leader) mbar_arrive(fb);
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    // the leader's barrier collects the bytes of BOTH CTAs (the 2-CTA TMA form signals the leader)
                    if (leader) mbar_arrive_expect_tx(fb, (uint32_t)(CG * nsub) * (a_bytes + (bstat ? 0u : b_bytes)));
                    const long long ct0 = ST ? clock64() : 0;
                    for (int j = 0; j < nsub; ++j) {
                        const uint32_t ad = a_dst + (uint32_t)j * a_bytes, bd = b_dst + (uint32_t)j * b_bytes;
                        const int c0 = cb * BK;
                        if constexpr (CG == 2) {
                            if (stride2) tma2_load_5d(ad, &tmA, fb, c0, kx & 1, x0 + (kx >> 1), ky & 1, J0 + (ky >> 1));
                            else tma2_load_3d(ad, &tmA, fb, c0, x0 + kx + xoff, J0 + ky + yoff);
                            tma2_load_2d(bd, &tmB, fb, kcol, n0);
                        } else {
                            if (stride2) tma_load_5d(ad, &tmA, fb, c0, kx & 1, x0 + (kx >> 1), ky & 1, J0 + (ky >> 1));
                            else tma_load_3d(ad, &tmA, fb, c0, x0 + kx + xoff, J0 + ky + yoff);
                            if (!bstat) tma_load_2d(bd, &tmB, fb, kcol, n0);
                        }
                        kcol += BK;
                        if (++cb == cblocks) { cb = 0; if (++kx == fsize) { kx = 0; ++ky; } }
                    }
                    if constexpr (ST) w_tma += clock64() - ct0;
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
            }
            if (ST && p.stats) { p.stats[blockIdx.x * 16 + 0] = (unsigned long long)w_empty; p.stats[blockIdx.x * 16 + 1] = (unsigned long long)(clock64() - t_begin); p.stats[blockIdx.x * 16 + 7] = (unsigned long long)w_tma; }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (CG=2: the leader CTA only, for both CTAs) =======================
        // One thread issues every tcgen05.mma.  Measured (tools/probes/issue_probe.cu, profiles/r02_issue_probe.txt): an MMA
        // costs the issuing thread ~53 cycles, a loop trip ~230 more, commits are free -- so MMAs are issued in straight-line
        // batches of 8-18 (two pipeline stages, three taps, or a whole channel block at once), descriptors first.  The operand
        // kind and the MMAs per K-block are compile-time: one copy of the role per (kind, kk), chosen once per launch.
        auto mma_role = [&](auto kind_c, auto kk_c) {
            constexpr int KIND = decltype(kind_c)::value;
            constexpr int KK = decltype(kk_c)::value;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const int sps = p.sps, kblocks = p.kblocks, stages = p.stages, BN = p.BN;
            const uint32_t a_bytes = p.a_bytes, b_bytes = p.b_bytes, stage_bytes = p.stage_bytes, idesc = p.idesc;
            const uint32_t b_off = (uint32_t)sps * a_bytes;
            const uint32_t bhi = p.desc_hi;
            long long w_full = 0, w_tempty = 0; const long long t_begin = ST ? clock64() : 0;
            const int bstat = p.bstat;
            if (bstat) { mbar_wait(bstat_bar, 0, 4); tc_fence_after(); }
            auto lo = [](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };   // descriptor low word: address, LBO = 1
            // the KK MMAs of one K-block: K advance of 16 bf16 / 32 s8 / 8 tf32 = 32 bytes inside the swizzle row = +2 units
            auto issue_kb = [&](uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t first) {
#pragma unroll
                for (int k = 0; k < KK; ++k) {
                    const uint64_t adesc = ((uint64_t)ahi << 32) | (uint64_t)(alo + 2u * (uint32_t)k);
                    const uint64_t bdesc = ((uint64_t)bhi << 32) | (uint64_t)(blo + 2u * (uint32_t)k);
                    const uint32_t accum = (k == 0) ? (uint32_t)(first != 0u) : 1u;
                    if constexpr (CG == 2 && KIND == 1) umma2_i8(d_tmem, adesc, bdesc, idesc, accum);
                    else if constexpr (CG == 2) umma2_bf16(d_tmem, adesc, bdesc, idesc, accum);
                    else if constexpr (KIND == 3) umma_tf32(d_tmem, adesc, bdesc, idesc, accum);
                    else if constexpr (KIND == 1) umma_i8(d_tmem, adesc, bdesc, idesc, accum);
                    else umma_bf16(d_tmem, adesc, bdesc, idesc, accum);
                }
            };
            auto wait_tempty = [&]() {   // the epilogue(s) drained this accumulator
                if constexpr (ST) { const long long c0 = clock64(); mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 1); w_tempty += clock64() - c0; }
                else mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 1);
                tc_fence_after();
            };
            auto wait_full = [&](int st_, uint32_t ph_) {   // TMA bytes of a ring stage have landed
                if constexpr (ST) { const long long c0 = clock64(); mbar_wait(full_bar(st_), ph_, 2); w_full += clock64() - c0; }
                else mbar_wait(full_bar(st_), ph_, 2);
            };
            auto release = [&](uint32_t bar) {   // arrives on `bar` (of both CTAs) when every MMA issued so far has retired
                if constexpr (CG == 2) umma2_commit_both(bar); else umma_commit(bar);
            };
            if (p.halo) {
                // ---- halo mode: per channel block one activation tile (with halo) and nine filter tiles; tap (ky, kx) reads the
                // activation tile from line ky*(TW+2) + kx on
                const int SA = p.a_stages, cblocks = p.cblocks;
                const uint32_t a_stage_bytes = p.a_stage_bytes, pitch = p.halo_pitch, rb = pitch / (uint32_t)(p.TW + (p.halo == 2 ? 1 : 2));
                const uint32_t ahi = p.desc_hi_a;
                // byte offset of tap (ky, kx) inside the activation tile: line ky*(TW+2) + kx, or (stride 2) plane (ky&1, kx&1),
                // line (ky>>1)*(TW+1) + (kx>>1)
                uint32_t toff[9];
                {
                    const uint32_t plane = a_stage_bytes >> 2;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t ky = (uint32_t)(t / 3), kx = (uint32_t)(t % 3);
                        toff[t] = p.halo == 2 ? ((ky & 1u) * 2u + (kx & 1u)) * plane + (ky >> 1) * pitch + (kx >> 1) * rb
                                              : ky * pitch + kx * rb;
                    }
                }
                int sa = 0; uint32_t pha = 0;
                TcSched sch = sched_init<false>(p, w_first, w_step);
                int w, seg0, seg1;
                while (sched_next<false>(sch, w, seg0, seg1)) {
                    wait_tempty();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    for (int cb = 0; cb < cblocks; ++cb) {
                        if constexpr (ST) { const long long c0 = clock64(); mbar_wait(fullA_bar(sa), pha, 6); w_full += clock64() - c0; }
                        else mbar_wait(fullA_bar(sa), pha, 6);
                        tc_fence_after();
                        const uint32_t a_tile = smemA + (uint32_t)sa * a_stage_bytes;
                        if (bstat) {
                            // resident filter matrix: all nine taps of the channel block in one straight-line batch
                            const uint32_t b0 = smemB + (uint32_t)cb * b_bytes, bstep = (uint32_t)cblocks * b_bytes;
#pragma unroll
                            for (int t = 0; t < 9; ++t)
                                issue_kb(d_tmem, lo(a_tile + toff[t]), ahi, lo(b0 + (uint32_t)t * bstep), (uint32_t)(cb | t));
                        } else {
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {        // one filter row = three K-blocks of straight-line code
                                // each stage is waited for right in front of its own MMAs and released right behind them (a commit
                                // covers everything issued before it).  Waiting for all three stages first made a 5-stage ring stall
                                // once per batch: only two of the next three stages can be in flight while a batch executes
                                // (role counters: MMA thread 62 % in wait_full at 75 % tensor utilisation, profiles/r02_notes.md).
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    wait_full(stage, phase);
                                    tc_fence_after();
                                    issue_kb(d_tmem, lo(a_tile + toff[ky * 3 + kx]), ahi, lo(smem0 + (uint32_t)stage * stage_bytes),
                                             (uint32_t)(cb | ky | kx));
                                    release(empty_bar(stage));
                                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                                }
                            }
                        }
                        release(emptyA_bar(sa));
                        if (++sa == SA) { sa = 0; pha ^= 1u; }
                    }
                    release(tfull_bar(acc));     // accumulator complete -> epilogue warps (of both CTAs)
                    if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1u; }
                }
            } else {
                const uint32_t ahi = p.desc_hi;
                TcSched sch = sched_init<KS>(p, w_first, w_step);
                int w, seg0, seg1;
                while (sched_next<KS>(sch, w, seg0, seg1)) {
                    wait_tempty();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                    const int kb_begin = seg0 * sps, kb_end = min(kblocks, seg1 * sps);
                    int kb0 = kb_begin;
                    while (kb0 < kb_end) {
                        // one batch = the K-blocks of one stage (sps > 1) or of two consecutive stages (sps == 1): up to 4 entries
                        // one pipeline stage = up to 4 K-blocks (sps): wait, issue its MMAs in straight-line code, release it
                        const int nsub = min(sps, kb_end - kb0);
                        wait_full(stage, phase);
                        tc_fence_after();
                        const uint32_t a_base = smem0 + (uint32_t)stage * stage_bytes, b_base = a_base + b_off;
                        const uint32_t first = (uint32_t)(kb0 - kb_begin);   // 0 on the first K-block of the segment: overwrite
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < nsub)
                                issue_kb(d_tmem, lo(a_base + (uint32_t)j * a_bytes), ahi,
                                         lo(bstat ? smemB + (uint32_t)(kb0 + j) * b_bytes : b_base + (uint32_t)j * b_bytes), first | (uint32_t)j);
                        release(empty_bar(stage));     // frees the smem stage (in both CTAs) when these MMAs retire
                        kb0 += nsub;
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                    release(tfull_bar(acc));     // accumulator complete -> epilogue warps (of both CTAs)
                    if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1u; }
                }
            }
            if (ST && p.stats) { p.stats[blockIdx.x * 16 + 2] = (unsigned long long)w_full; p.stats[blockIdx.x * 16 + 3] = (unsigned long long)w_tempty;
                                 p.stats[blockIdx.x * 16 + 4] = (unsigned long long)(clock64() - t_begin); }
        };
        if (leader && elect_one()) {
            auto by_kk = [&](auto kind_c) {
                if (p.kk == 4) mma_role(kind_c, std::integral_constant<int, 4>{});
                else if (p.kk == 2) mma_role(kind_c, std::integral_constant<int, 2>{});
                else mma_role(kind_c, std::integral_constant<int, 1>{});
            };
            if constexpr (EPI == 2) by_kk(std::integral_constant<int, 1>{});
            else if constexpr (CG == 2 || EPI == 1) by_kk(std::integral_constant<int, 0>{});
            else if (p.kind == 0) by_kk(std::integral_constant<int, 0>{});
            else by_kk(std::integral_constant<int, 3>{});
        }
    } else {
        // ======================= epilogue (warps 2..9) =======================
        // Per 64-column slab: issue both TMEM loads and the residual loads first, wait once, then do the math and
        // the stores -- global-load latency is paid once per slab instead of once per value (the first version
        // was epilogue-bound on exactly that, profiles/r01_notes.md).
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;         // which half of the columns this warp owns
        const int cbeg = p.epi_alt ? 0 : (p.BN >= 64) ? half * (p.BN >> 1) : 0;
        const int cend = p.epi_alt ? p.BN : (p.BN >= 64) ? cbeg + (p.BN >> 1) : (half == 0 ? p.BN : 0);
        const int r = q * 32 + lane;              // accumulator row == pixel within the tile
        const int tx = r & (p.TW - 1), ty = r >> p.TWlog2;
        const bool leaky = p.act == ACT_LEAKY, leaky2 = p.act2 == ACT_LEAKY;
        int acc = p.epi_alt ? half : 0; uint32_t acc_phase = 0;
        uint32_t epi_res_phase = 0;               // TMA epilogue: parity of this group's residual barrier
        bool res_requested = false;               // TMA epilogue: the residual tile of the slab about to be processed is on its way
        int out_buf = 0;                          // TMA epilogue: which of the group's epi_bufs OUT tiles the next slab uses
        int tile_cnt = 0;
        long long w_tfull = 0, w_res = 0; const long long t_begin = ST ? clock64() : 0;
        TcSched sch = sched_init<KS>(p, w_first, w_step);
        int w, seg0, seg1;
        while (sched_next<KS>(sch, w, seg0, seg1)) {
            if (p.epi_alt && ((tile_cnt++ & 1) != half)) continue;   // the other group's tile
            // K-split tail: a segment that stops short of the work item's last stage only publishes its raw accumulator;
            // the segment that ends the work item adds the npart partials of the CTAs (pairs) gA .. unit-1 before it
            const bool seg_partial = KS && seg1 < p.kbs;
            const int npart = (KS && seg0 > 0 && !seg_partial) ? w_first - (w * p.kbs) / p.sk_L : 0;
            const int n_idx = w % p.nt;
            const int m = (CG == 2) ? 2 * (w / p.nt) + (int)rank : w / p.nt;
            const int ox = (m % p.xt) * p.TW + tx;
            const int J = (m / p.xt) * p.TH + p.jshift + ty;
            const int n0 = n_idx * p.BN;
            const int img = J / p.PR;
            const int oy = J - img * p.PR - p.row_off;
            const bool valid = (img < p.N) && (oy >= 0) && (oy < p.OH) && (ox < p.OW) && !(p.dbg & 4);
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
            // the staged bf16 store paths fetch their own residual, the integer / tf32 kinds never have one: rv[] is only
            // prefetched for the per-thread store path (f32 heads, YB_TC_NO_COALESCE)
            const bool own_res = EPI != 0 || (p.out_bf16 && ((cend - cbeg) >= 64 || (cend - cbeg) == 32) && !p.no_coalesce) || !p.res;
            if (!own_res) {
                if (cbeg < cend) load_res(cbeg, rv[0]);
                if (cend - cbeg > 32) load_res(cbeg + 32, rv[1]);
            }

            const bool path64 = EPI == 0 && !seg_partial && p.out_bf16 && (cend - cbeg) >= 64 && !p.no_coalesce;

            // (the TMA epilogue requests its first residual tile before this wait and waits itself)
            const bool tfull_waited = EPI != 1;
            if (tfull_waited) {
                if constexpr (ST) { const long long c0 = clock64(); mbar_wait(tfull_bar(acc), acc_phase, 3); w_tfull += clock64() - c0; }
                else mbar_wait(tfull_bar(acc), acc_phase, 3);
                tc_fence_after();
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);

            if constexpr (KS) {
                auto flag_of = [&](int unit) { return p.sk_flags + ((size_t)(unit * CG + (int)rank) * TC_EPI_WARPS + (warp - 2)); };
                auto ws_of = [&](int unit) { return reinterpret_cast<float4 *>(p.sk_ws) + (size_t)(unit * CG + (int)rank) * (size_t)(TC_BM * 64); };
                if (seg_partial) {
                    // publish the raw f32 accumulator ([column/4][row] float4: a warp store is 512 contiguous bytes)
                    float4 *dst = ws_of(w_first);
                    for (int f0 = cbeg; f0 < cend; f0 += 32) {
                        uint32_t v0[32];
                        tmem_ld32(taddr + (uint32_t)f0, v0);
                        tmem_ld_wait();
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4)
                            __stcg(dst + (size_t)((f0 >> 2) + g4) * TC_BM + r,
                                   make_float4(__uint_as_float(v0[g4 * 4]), __uint_as_float(v0[g4 * 4 + 1]),
                                               __uint_as_float(v0[g4 * 4 + 2]), __uint_as_float(v0[g4 * 4 + 3])));
                    }
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release_u32(flag_of(w_first), 1u);
                } else if (npart) {
                    // owner: fold the partial sums of CTAs (pairs) gA .. unit-1 into the TMEM accumulator, then run the
                    // ordinary epilogue below on it
                    const int gA = w_first - npart;
                    for (int pc = 0; pc < npart; ++pc) {
                        const unsigned *fl = flag_of(gA + pc);
                        if (ld_acquire_u32(fl) == 0u) {
                            const long long t0 = clock64();
                            while (ld_acquire_u32(fl) == 0u) {
                                if (clock64() - t0 > 4000000000LL) { printf("yb k_conv_tc: K-split partial timeout (block %d warp %d)\n", blockIdx.x, warp); __trap(); }
                            }
                        }
                    }
                    // 16 float4 loads in flight per thread: the fold is a latency-bound L2 read (profiles/r01_notes.md)
                    auto add4 = [](uint32_t (&v)[32], int g4, const float4 &t) {
                        v[g4 * 4 + 0] = __float_as_uint(__uint_as_float(v[g4 * 4 + 0]) + t.x);
                        v[g4 * 4 + 1] = __float_as_uint(__uint_as_float(v[g4 * 4 + 1]) + t.y);
                        v[g4 * 4 + 2] = __float_as_uint(__uint_as_float(v[g4 * 4 + 2]) + t.z);
                        v[g4 * 4 + 3] = __float_as_uint(__uint_as_float(v[g4 * 4 + 3]) + t.w);
                    };
                    for (int f0 = cbeg; f0 < cend; f0 += 64) {
                        if (cend - f0 >= 64) {
                            uint32_t v0[32], v1[32];
                            tmem_ld32(taddr + (uint32_t)f0, v0);
                            tmem_ld32(taddr + (uint32_t)f0 + 32u, v1);
                            for (int pc = 0; pc < npart; ++pc) {
                                const float4 *src = ws_of(gA + pc) + (size_t)(f0 >> 2) * TC_BM + r;
                                float4 t[16];
#pragma unroll
                                for (int g4 = 0; g4 < 16; ++g4) t[g4] = __ldcg(src + (size_t)g4 * TC_BM);
                                if (pc == 0) tmem_ld_wait();
#pragma unroll
                                for (int g4 = 0; g4 < 8; ++g4) { add4(v0, g4, t[g4]); add4(v1, g4, t[8 + g4]); }
                            }
                            tmem_st32(taddr + (uint32_t)f0, v0);
                            tmem_st32(taddr + (uint32_t)f0 + 32u, v1);
                        } else {
                            uint32_t v[32];
                            tmem_ld32(taddr + (uint32_t)f0, v);
                            tmem_ld_wait();
                            for (int pc = 0; pc < npart; ++pc) {
                                const float4 *src = ws_of(gA + pc) + (size_t)(f0 >> 2) * TC_BM + r;
                                float4 t[8];
#pragma unroll
                                for (int g4 = 0; g4 < 8; ++g4) t[g4] = __ldcg(src + (size_t)g4 * TC_BM);
#pragma unroll
                                for (int g4 = 0; g4 < 8; ++g4) add4(v, g4, t[g4]);
                            }
                            tmem_st32(taddr + (uint32_t)f0, v);
                        }
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) for (int pc = 0; pc < npart; ++pc) *flag_of(gA + pc) = 0u;   // re-arm for the next launch
                }
            }

            auto finish = [&](const uint32_t (&v)[32], const uint4 (&rr)[4], int f0) {
                if (!valid || (n0 + f0) >= p.n_store) return;
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float a = __uint_as_float(v[j]) + bs[f0 + j];
                    x[j] = leaky ? fmaxf(a, 0.1f * a) : a;   // == a > 0 ? a : 0.1a
                }
                if (p.res) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t wv[4] = {rr[g].x, rr[g].y, rr[g].z, rr[g].w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            x[g * 8 + 2 * h] += __uint_as_float(wv[h] << 16);
                            x[g * 8 + 2 * h + 1] += __uint_as_float(wv[h] & 0xffff0000u);
                        }
                    }
                    if (leaky2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.1f * x[j]);
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
                } else if (p.yolo_out) {
                    // detection head with the [yolo] layer fused: logistic on x, y, objectness and class entries (w, h stay
                    // raw), written straight into the NCHW tensor the reference decoder reads -- the f32 NHWC copy of the
                    // head and the separate yolo kernel disappear
                    const int c0 = n0 + f0;
                    const int nvalid = p.n - c0;                       // columns >= n are padding
                    const uint32_t raw = ymask_s[c0 >> 5];             // bit j: entry (c0 + j) % (4+classes+1) is w or h -> stays raw
                    float *dst = p.yolo_out + (((size_t)img * p.n + c0) * p.OH + oy) * (size_t)p.OW + ox;
                    const size_t plane = (size_t)p.OH * p.OW;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float sg = __fdividef(1.f, 1.f + __expf(-x[j]));
                        const float v = ((raw >> j) & 1u) ? x[j] : sg;
                        if (j < nvalid) dst[(size_t)j * plane] = v;
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

            // ---- coalesced f32 store of one 32-column slab (the integer kinds and the f32 heads): a row is 128 B; the warp's 32
            // rows go through its private XOR-swizzled 4 KB staging tile so that every global store instruction writes 4 rows x
            // 128 contiguous bytes instead of 32 rows x 16 B (32 different lines per instruction: what kept the early INT8
            // layers at 0.8 TB/s in round 1)
            auto store_f32_slab = [&](const float (&y)[32], int f0) {
                const uint32_t stg = stg_base + (uint32_t)(warp - 2) * 4096u;
                const int srow = lane >> 3, schunk = lane & 7;
                const unsigned long long obase = (unsigned long long)(uintptr_t)orow;
                const int vflag = valid ? 1 : 0;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4)),
                                 "r"(__float_as_uint(y[c * 4 + 0])), "r"(__float_as_uint(y[c * 4 + 1])),
                                 "r"(__float_as_uint(y[c * 4 + 2])), "r"(__float_as_uint(y[c * 4 + 3])) : "memory");
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = i * 4 + srow;
                    const unsigned long long op = __shfl_sync(0xffffffffu, obase, row);
                    const int ok = __shfl_sync(0xffffffffu, vflag, row);
                    uint32_t w0, w1, w2, w3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                 : "r"(stg + (uint32_t)row * 128u + (uint32_t)((schunk ^ (row & 7)) << 4)) : "memory");
                    if (ok && (n0 + f0 + schunk * 4) < p.n_store)
                        *(reinterpret_cast<uint4 *>(op + (size_t)(n0 + f0) * 4) + schunk) = make_uint4(w0, w1, w2, w3);
                }
                __syncwarp();
            };

            // ---- fused 2x2/2 max-pool + conversion to the next integer layer's input (TcParams::pool_mode).  Both epilogue functions
            // are monotone non-decreasing in the s32 accumulator (truncating /32, clamp, x positive ALPHA1, + bias, leaky; resp.
            // x mean >= 0, + bias, leaky), so max over the window commutes with them EXACTLY: the window maximum is taken on the raw
            // accumulators and the float epilogue runs once per pooled value.  Window = lanes {l, l^1, l^8, l^9} (tile rows are 8
            // pixels wide).  The reduction is a reduce-scatter: lane^1 halves the 32 columns, lane^8 halves them again, every lane
            // ends up with the maxima of 8 columns of its window and finishes those -- a quarter of the float work per lane.
            // Elements outside the image count as "skipped" (INT_MIN) like the reference's out-of-range taps (additionally.c:1448-1482).
            auto pool_store_raw = [&](const uint32_t (&v)[32], int f0) {
                const bool b0 = lane & 1, b3 = lane & 8;
                int h16[16], h8[8];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int lo_ = valid ? (int)v[j] : INT_MIN, hi_ = valid ? (int)v[j + 16] : INT_MIN;
                    const int keep = b0 ? hi_ : lo_, send = b0 ? lo_ : hi_;
                    const int got = __shfl_xor_sync(0xffffffffu, send, 1);
                    h16[j] = max(keep, got);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int keep = b3 ? h16[j + 8] : h16[j], send = b3 ? h16[j] : h16[j + 8];
                    const int got = __shfl_xor_sync(0xffffffffu, send, 8);
                    h8[j] = max(keep, got);
                }
                const int cbase = f0 + (b0 ? 16 : 0) + (b3 ? 8 : 0);     // this lane's 8 columns of the slab
                uint32_t w0 = 0, w1 = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int f = n0 + cbase + j;
                    float t;
                    if (p.kind == 1) {
                        int q16 = h8[j] / 32;
                        q16 = q16 > 32767 ? 32767 : (q16 < -32767 ? -32767 : q16);
                        t = __fadd_rn(__fmul_rn((float)q16, p.alpha1), bs[cbase + j]);
                        t = (p.act == ACT_LEAKY) ? ((t > 0.f) ? t : __fdiv_rn(t, 10.f)) : t;
                    } else {
                        t = __fadd_rn(__fmul_rn((float)h8[j], (f < p.n) ? __ldg(p.mean + f) : 0.f), bs[cbase + j]);
                        t = act_exact(t, p.act);
                    }
                    int b8 = (p.pool_mode == 1) ? quant_i8(t, p.pool_mult) : (t > 0.f ? 1 : -1);
                    if (f >= p.n) b8 = 0;
                    if (j < 4) w0 |= (uint32_t)(b8 & 0xff) << (8 * j); else w1 |= (uint32_t)(b8 & 0xff) << (8 * (j - 4));
                }
                const int oye = oy & ~1, oxe = ox & ~1;                  // the window's origin: the same pooled pixel in all four lanes
                if (img < p.N && oye >= 0 && oye < p.OH && oxe < p.OW && !(p.dbg & 4)) {
                    signed char *dst = p.pool_out + ((size_t)(img * p.pool_Hp + (oye >> 1) + 1) * p.pool_Wp + (oxe >> 1) + 1) * (size_t)p.pool_ldc + n0 + cbase;
                    *reinterpret_cast<uint2 *>(dst) = make_uint2(w0, w1);
                }
            };

            // ---- TMA store of one 32-column f32 slab (integer kinds): the group's [128 pixels][32 floats] tile, 128-byte rows with
            // the 128B swizzle, written by one thread per row and stored by one cp.async.bulk.tensor
            auto tma_store_f32_slab = [&](const float (&y)[32], int f0) {
                const int g = half;
                const uint32_t out_tile = stg_base + (uint32_t)(g * p.epi_bufs + out_buf) * 16384u;
                const bool boss = (q == 0) && (lane == 0);
                const uint32_t rsw = (uint32_t)(r & 7), row_off = (uint32_t)r * 128u;
                if (boss) { if (p.epi_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read0(); }   // the store that last used this tile has finished reading it
                if (p.epi_bufs == 2) out_buf ^= 1;
                named_bar_sync(1 + g, 128);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t o0 = __float_as_uint(y[c * 4 + 0]), o1 = __float_as_uint(y[c * 4 + 1]);
                    uint32_t o2 = __float_as_uint(y[c * 4 + 2]), o3 = __float_as_uint(y[c * 4 + 3]);
                    if (!valid) { o0 = o1 = o2 = o3 = 0u; }           // border / padding rows inside the tensor stay zero
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(out_tile + row_off + (((uint32_t)c ^ rsw) << 4)),
                                 "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar_sync(1 + g, 128);
                if (boss) {
                    tma_store_3d(&tmO, out_tile, n0 + f0, (m % p.xt) * p.TW + 1, (m / p.xt) * p.TH + p.jshift);
                    tma_store_commit();
                }
            };

            if constexpr (EPI == 1) {
                // ---- TMA epilogue.  Group g = the four warps that own column half `half` (128 threads, named barrier 1 + half);
                // per slab of SW = 64 (or 32) columns: [OUT tile][RES tile], both [128 pixel rows][SW * 2 bytes] with the swizzle of
                // the tensor maps -- 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4) for 128-byte rows, at r*64 +
                // ((c ^ ((r >> 1) & 3)) << 4) for 64-byte rows: conflict-free for one thread per row.
                auto slabs = [&](auto nv_c) {
                    constexpr int NV = decltype(nv_c)::value;          // tcgen05.ld.x32 per slab: 2 (SW = 64) or 1 (SW = 32)
                    constexpr int SW = 32 * NV;
                    constexpr uint32_t tile_bytes = 128u * SW * 2u, rowb = SW * 2u;
                    const int g = half;
                    const uint32_t grp_base = stg_base + (uint32_t)g * ((uint32_t)(p.epi_bufs + 1) * tile_bytes);   // [OUT x epi_bufs][RES]
                    const uint32_t res_tile = grp_base + (uint32_t)p.epi_bufs * tile_bytes;
                    const bool boss = (q == 0) && (lane == 0);           // issues this group's TMA traffic
                    const int x0 = (m % p.xt) * p.TW, J0 = (m / p.xt) * p.TH + p.jshift;
                    const uint32_t rsw = (NV == 2) ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
                    const uint32_t row_off = (uint32_t)r * rowb;
                    const bool has_res = p.res != nullptr;
                    // residual tiles are requested one slab AHEAD, across tile boundaries (the next work item of this group is
                    // known: w + step): the TMA latency hides behind the store phase of this slab and the accumulator wait of the next
                    const int wstep_g = p.epi_alt ? 2 * w_step : w_step;
                    auto request_res = [&](int w_, int f_) {
                        const int m_ = (CG == 2) ? 2 * (w_ / p.nt) + (int)rank : w_ / p.nt;
                        mbar_arrive_expect_tx(resfull_bar(g), tile_bytes);
                        if (p.l2_hint) tma_load_3d_hint(res_tile, &tmR, resfull_bar(g), (w_ % p.nt) * p.BN + f_, (m_ % p.xt) * p.TW + 1, (m_ / p.xt) * p.TH + p.jshift, l2_policy_evict_first());
                        else tma_load_3d(res_tile, &tmR, resfull_bar(g), (w_ % p.nt) * p.BN + f_, (m_ % p.xt) * p.TW + 1, (m_ / p.xt) * p.TH + p.jshift);
                    };
                    if (has_res && boss && !res_requested) request_res(w, cbeg);   // very first slab of this group
                    res_requested = true;
                    if (!tfull_waited) {
                        if constexpr (ST) { const long long c0 = clock64(); mbar_wait(tfull_bar(acc), acc_phase, 3); w_tfull += clock64() - c0; }
                        else mbar_wait(tfull_bar(acc), acc_phase, 3);
                        tc_fence_after();
                    }
                    for (int f0 = cbeg; f0 < cend; f0 += SW) {
                        uint32_t v[NV][32];
#pragma unroll
                        for (int h2 = 0; h2 < NV; ++h2) tmem_ld32(taddr + (uint32_t)(f0 + 32 * h2), v[h2]);
                        tmem_ld_wait();
#pragma unroll
                        for (int h2 = 0; h2 < NV; ++h2)
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float a0 = __uint_as_float(v[h2][j]) + bs[f0 + 32 * h2 + j];
                                v[h2][j] = __float_as_uint(leaky ? fmaxf(a0, 0.1f * a0) : a0);
                            }
                        if (has_res) {
                            if constexpr (ST) { const long long c0 = clock64(); mbar_wait(resfull_bar(g), epi_res_phase, 7); w_res += clock64() - c0; }
                            else mbar_wait(resfull_bar(g), epi_res_phase, 7);
                            epi_res_phase ^= 1u;
#pragma unroll
                            for (int c = 0; c < 4 * NV; ++c) {             // own row of the residual tile, 16 bytes at a time
                                uint32_t w0, w1, w2, w3;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                             : "r"(res_tile + row_off + (((uint32_t)c ^ rsw) << 4)) : "memory");
                                const uint32_t wv[4] = {w0, w1, w2, w3};
#pragma unroll
                                for (int h = 0; h < 4; ++h) {
                                    uint32_t &lo_ = v[c / 4][(c % 4) * 8 + 2 * h], &hi_ = v[c / 4][(c % 4) * 8 + 2 * h + 1];
                                    float a = __uint_as_float(lo_) + __uint_as_float(wv[h] << 16);
                                    float b = __uint_as_float(hi_) + __uint_as_float(wv[h] & 0xffff0000u);
                                    if (leaky2) { a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b); }
                                    lo_ = __float_as_uint(a); hi_ = __float_as_uint(b);
                                }
                            }
                        }
                        // the TMA store that last used this OUT tile must have finished READING it before it is overwritten
                        const uint32_t out_tile = grp_base + (uint32_t)out_buf * tile_bytes;
                        if (boss) { if (p.epi_bufs == 2) tma_store_wait_read1(); else tma_store_wait_read0(); }
                        if (p.epi_bufs == 2) out_buf ^= 1;
                        named_bar_sync(1 + g, 128);
                        if (has_res && boss) {                             // everybody is done with the RES tile: request the next one
                            if (f0 + SW < cend) request_res(w, f0 + SW);
                            else if (w + wstep_g < p.num_work) request_res(w + wstep_g, cbeg);
                        }
#pragma unroll
                        for (int c = 0; c < 4 * NV; ++c) {                 // own row -> OUT tile (border / padding rows: zeros)
                            const uint32_t *src = &v[c / 4][(c % 4) * 8];
                            uint32_t o0 = pack_bf16x2(__uint_as_float(src[0]), __uint_as_float(src[1]));
                            uint32_t o1 = pack_bf16x2(__uint_as_float(src[2]), __uint_as_float(src[3]));
                            uint32_t o2 = pack_bf16x2(__uint_as_float(src[4]), __uint_as_float(src[5]));
                            uint32_t o3 = pack_bf16x2(__uint_as_float(src[6]), __uint_as_float(src[7]));
                            if (!valid) { o0 = o1 = o2 = o3 = 0u; }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(out_tile + row_off + (((uint32_t)c ^ rsw) << 4)),
                                         "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA engine
                        named_bar_sync(1 + g, 128);
                        if (boss) {
                            tma_store_3d(&tmO, out_tile, n0 + f0, x0 + 1, J0);
                            tma_store_commit();
                        }
                    }
                };
                if (p.tma_epi == 64) slabs(std::integral_constant<int, 2>{});
                else slabs(std::integral_constant<int, 1>{});
            } else if constexpr (EPI == 2) {
            if (p.kind == 2) {
                // ---- XNOR as +-1 int8: acc == 2*count - K (exact); out = act((float)acc * mean + bias) in the reference's
                // float op order (additionally.c:1531, yolov2_forward_network.c:243-261)
                for (int f0 = cbeg; f0 < cend; f0 += 32) {
                    uint32_t v0[32];
                    tmem_ld32(taddr + (uint32_t)f0, v0);
                    tmem_ld_wait();
                    if (p.pool_mode) { pool_store_raw(v0, f0); continue; }
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int f = n0 + f0 + j;
                        float t = __fmul_rn((float)(int)v0[j], (f < p.n) ? __ldg(p.mean + f) : 0.f);
                        t = __fadd_rn(t, bs[f0 + j]);
                        y[j] = act_exact(t, p.act);
                    }
                    if (p.no_coalesce) {
                        if (valid) {
                            float *orow_f = reinterpret_cast<float *>(orow);
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                if (n0 + f0 + g * 4 >= p.n_store) break;
                                *reinterpret_cast<float4 *>(orow_f + n0 + f0 + g * 4) = make_float4(y[g * 4], y[g * 4 + 1], y[g * 4 + 2], y[g * 4 + 3]);
                            }
                        }
                    } else if (p.tma_epi) tma_store_f32_slab(y, f0);
                    else store_f32_slab(y, f0);
                    if (!valid) continue;
                    if (p.acc_out) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int f = n0 + f0 + j;
                            if (f < p.n) p.acc_out[(((size_t)img * p.n + f) * p.OH + oy) * p.OW + ox] = ((int)v0[j] + p.xK) / 2;
                        }
                    }
                }
            } else
            if (p.kind == 1) {
                // ---- INT8: exact requantisation of the reference (yolov2_forward_network_quantized.c:474-490, :598-627):
                // q16 = clamp(+-32767, acc / 32) [C truncating division]; y = (float)q16 * ALPHA1; y += bias; leaky: y / 10.
                for (int f0 = cbeg; f0 < cend; f0 += 32) {
                    uint32_t v0[32];
                    tmem_ld32(taddr + (uint32_t)f0, v0);
                    tmem_ld_wait();
                    if (p.pool_mode) { pool_store_raw(v0, f0); continue; }
                    float y[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int a = (int)v0[j];
                        int q16 = a / 32;
                        q16 = q16 > 32767 ? 32767 : (q16 < -32767 ? -32767 : q16);
                        float t = __fmul_rn((float)q16, p.alpha1);
                        t = __fadd_rn(t, bs[f0 + j]);
                        y[j] = (p.act == ACT_LEAKY) ? ((t > 0.f) ? t : __fdiv_rn(t, 10.f)) : t;
                    }
                    if (p.no_coalesce) {
                        if (valid) {
                            float *orow_f = reinterpret_cast<float *>(orow);
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                if (n0 + f0 + g * 4 >= p.n_store) break;
                                *reinterpret_cast<float4 *>(orow_f + n0 + f0 + g * 4) = make_float4(y[g * 4], y[g * 4 + 1], y[g * 4 + 2], y[g * 4 + 3]);
                            }
                        }
                    } else if (p.tma_epi) tma_store_f32_slab(y, f0);
                    else store_f32_slab(y, f0);
                    if (!valid) continue;
                    if (p.acc_out) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int f = n0 + f0 + j;
                            if (f < p.n) p.acc_out[(((size_t)img * p.n + f) * p.OH + oy) * p.OW + ox] = (int)v0[j];
                        }
                    }
                }
            }
            } else {
            if (seg_partial) {
                // accumulator already published above
            } else
            if (path64) {
                // ---- coalesced path: every global access of this warp is a run of whole 128-byte lines.
                // Each warp owns 32 accumulator rows; per 64-column slab a row is 128 B of bf16.  Rows are
                // transposed through the warp's private swizzled staging tile so that one warp instruction moves
                // 4 rows x 128 B instead of 32 rows x 16 B (the latter costs 32 LSU cycles per instruction and made
                // the epilogue the bottleneck of every layer, profiles/r01_notes.md).
                const uint32_t stg = stg_base + (uint32_t)(warp - 2) * 4096u;
                const int srow = lane >> 3, schunk = lane & 7;
                const unsigned long long obase = (unsigned long long)(uintptr_t)orow;
                const unsigned long long rbase = (unsigned long long)(uintptr_t)rrow;
                const int vflag = valid ? 1 : 0;
                auto stage_addr = [&](int row, int chunk) { return stg + (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); };
                auto res_to_stage = [&](int f0) {     // coalesced global -> staging
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + srow;
                        const unsigned long long rp = __shfl_sync(0xffffffffu, rbase, row);
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (rp && (n0 + f0 + schunk * 8) < p.n_store)
                            v = __ldg(reinterpret_cast<const uint4 *>(rp + (size_t)(n0 + f0) * 2) + schunk);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr(row, schunk)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                };
                if (p.res) res_to_stage(cbeg);
                for (int f0 = cbeg; f0 < cend; f0 += 64) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + (uint32_t)f0, v0);
                    tmem_ld32(taddr + (uint32_t)f0 + 32u, v1);
                    tmem_ld_wait();
                    float x[64];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float a0 = __uint_as_float(v0[j]) + bs[f0 + j], a1 = __uint_as_float(v1[j]) + bs[f0 + 32 + j];
                        x[j] = leaky ? fmaxf(a0, 0.1f * a0) : a0;
                        x[32 + j] = leaky ? fmaxf(a1, 0.1f * a1) : a1;
                    }
                    if (p.res) {
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 8; ++c) {          // own row back from staging
                            uint32_t w0, w1, w2, w3;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(stage_addr(lane, c)) : "memory");
                            const uint32_t wv[4] = {w0, w1, w2, w3};
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                x[c * 8 + 2 * h] += __uint_as_float(wv[h] << 16);
                                x[c * 8 + 2 * h + 1] += __uint_as_float(wv[h] & 0xffff0000u);
                            }
                        }
                        if (leaky2) {
#pragma unroll
                            for (int j = 0; j < 64; ++j) x[j] = fmaxf(x[j], 0.1f * x[j]);
                        }
                        __syncwarp();
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c)                 // own row -> staging
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr(lane, c)),
                                     "r"(pack_bf16x2(x[c * 8 + 0], x[c * 8 + 1])), "r"(pack_bf16x2(x[c * 8 + 2], x[c * 8 + 3])),
                                     "r"(pack_bf16x2(x[c * 8 + 4], x[c * 8 + 5])), "r"(pack_bf16x2(x[c * 8 + 6], x[c * 8 + 7])) : "memory");
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {               // staging -> coalesced global
                        const int row = i * 4 + srow;
                        const unsigned long long op = __shfl_sync(0xffffffffu, obase, row);
                        const int ok = __shfl_sync(0xffffffffu, vflag, row);
                        uint32_t w0, w1, w2, w3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(stage_addr(row, schunk)) : "memory");
                        if (ok && (n0 + f0 + schunk * 8) < p.n_store)
                            *(reinterpret_cast<uint4 *>(op + (size_t)(n0 + f0) * 2) + schunk) = make_uint4(w0, w1, w2, w3);
                    }
                    __syncwarp();
                    if (p.res && f0 + 64 < cend) res_to_stage(f0 + 64);   // next slab's residual in flight
                }
            } else if (p.out_bf16 && (cend - cbeg) == 32 && !p.no_coalesce) {
                // ---- coalesced path for 32-column slabs (BN = 32 / 64): a row is 64 B of bf16; the warp's staging
                // tile is 32 rows x 64 B (16-byte chunk j of row r at r*64 + ((j ^ ((r >> 1) & 3)) << 4)); one
                // instruction moves 8 rows x 64 B.
                const uint32_t stg = stg_base + (uint32_t)(warp - 2) * 4096u;
                const int srow = lane >> 2, schunk = lane & 3;
                const unsigned long long obase = (unsigned long long)(uintptr_t)orow;
                const unsigned long long rbase = (unsigned long long)(uintptr_t)rrow;
                const int vflag = valid ? 1 : 0;
                auto stage_addr = [&](int row, int chunk) { return stg + (uint32_t)row * 64u + (uint32_t)((chunk ^ ((row >> 1) & 3)) << 4); };
                const int f0 = cbeg;
                if (p.res) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = i * 8 + srow;
                        const unsigned long long rp = __shfl_sync(0xffffffffu, rbase, row);
                        uint4 v = make_uint4(0u, 0u, 0u, 0u);
                        if (rp && (n0 + f0 + schunk * 8) < p.n_store)
                            v = __ldg(reinterpret_cast<const uint4 *>(rp + (size_t)(n0 + f0) * 2) + schunk);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr(row, schunk)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                }
                uint32_t v0[32];
                tmem_ld32(taddr + (uint32_t)f0, v0);
                tmem_ld_wait();
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float a0 = __uint_as_float(v0[j]) + bs[f0 + j];
                    x[j] = leaky ? fmaxf(a0, 0.1f * a0) : a0;
                }
                if (p.res) {
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t w0, w1, w2, w3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(stage_addr(lane, c)) : "memory");
                        const uint32_t wv[4] = {w0, w1, w2, w3};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            x[c * 8 + 2 * h] += __uint_as_float(wv[h] << 16);
                            x[c * 8 + 2 * h + 1] += __uint_as_float(wv[h] & 0xffff0000u);
                        }
                    }
                    if (leaky2) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.1f * x[j]);
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_addr(lane, c)),
                                 "r"(pack_bf16x2(x[c * 8 + 0], x[c * 8 + 1])), "r"(pack_bf16x2(x[c * 8 + 2], x[c * 8 + 3])),
                                 "r"(pack_bf16x2(x[c * 8 + 4], x[c * 8 + 5])), "r"(pack_bf16x2(x[c * 8 + 6], x[c * 8 + 7])) : "memory");
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = i * 8 + srow;
                    const unsigned long long op = __shfl_sync(0xffffffffu, obase, row);
                    const int ok = __shfl_sync(0xffffffffu, vflag, row);
                    uint32_t w0, w1, w2, w3;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(stage_addr(row, schunk)) : "memory");
                    if (ok && (n0 + f0 + schunk * 8) < p.n_store)
                        *(reinterpret_cast<uint4 *>(op + (size_t)(n0 + f0) * 2) + schunk) = make_uint4(w0, w1, w2, w3);
                }
                __syncwarp();
            } else
            for (int f0 = cbeg; f0 < cend; f0 += 64) {
                if (cend - f0 >= 64) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + (uint32_t)f0, v0);
                    tmem_ld32(taddr + (uint32_t)f0 + 32u, v1);
                    tmem_ld_wait();
                    uint4 r0[4], r1[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) { r0[g] = rv[0][g]; r1[g] = rv[1][g]; }
                    if (f0 + 64 < cend) {                                   // next slab's residual in flight
                        load_res(f0 + 64, rv[0]);
                        if (cend - (f0 + 64) > 32) load_res(f0 + 96, rv[1]);
                    }
                    finish(v0, r0, f0);
                    finish(v1, r1, f0 + 32);
                } else {
                    uint32_t v0[32];
                    tmem_ld32(taddr + (uint32_t)f0, v0);
                    tmem_ld_wait();
                        finish(v0, rv[0], f0);
                }
            }
            }
            tc_fence_before();
            // all epilogue threads (of both CTAs) hand the accumulator back to the (leader's) MMA warp
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_remote(tempty_bar(acc), 0); }
            if (p.epi_alt) acc_phase ^= 1u;       // this group always works on accumulator `half`
            else if (++acc == TC_ACC) { acc = 0; acc_phase ^= 1u; }
        }
        if ((EPI == 1 || (EPI == 2 && p.tma_epi)) && (warp & 3) == 0 && lane == 0) tma_store_wait_all();   // this group's bulk stores have completed
        if (ST && p.stats && warp == 2 && lane == 0) { p.stats[blockIdx.x * 16 + 5] = (unsigned long long)w_tfull; p.stats[blockIdx.x * 16 + 6] = (unsigned long long)(clock64() - t_begin);
                                                       p.stats[blockIdx.x * 16 + 8] = (unsigned long long)w_res; }
    }

    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------
// Stem convolution (3 input channels, 3x3 / stride 1 / pad 1) on the tensor cores.
// K = 27 is padded to 32: every thread gathers the 3x3x3 window of ITS pixel straight from the caller's NCHW f32
// image (the layout conversion is fused away), converts to bf16 and writes one 64-byte row of a 64B-swizzled A tile;
// one thread then issues two tcgen05.mma (M=128, N=filters, K=16) and every thread drains its TMEM row through
// bias + leaky-ReLU into bf16 NHWC.  No TMA (the gather is irregular), single-buffered, several CTAs per SM.
// ------------------------------------------------------------------------------------------------------
struct StemTcP {
    const float *in;          // NCHW f32, set per call
    const unsigned char *in8; // or: HWC 8-bit frames of exactly the network size (k_stem_tc<true>), set per call
    char *out; int out_ldc;   // bf16 padded NHWC
    const __nv_bfloat16 *w;   // [32 rows (filters, zero padded)][32 k] bf16, k = (ky,kx,c), k >= 27 zero
    const float *bias;
    int N, H, W, OHp, OWp, nf, act;
    long npix;
    int ntiles;
};

// U8: the input is the caller's 8-bit HWC frame (already of the network size): value = (float)((double)v / 255.0) exactly as
// load_image_stb computes it (additionally.c:3093-3103), through a 256-entry table (the correctly rounded f32 division v / 255.f
// gives the same 256 floats, but ~8 instructions per value made this issue-bound kernel 0.3 ms slower) -- the u8 -> planar float pass over the batch
// (71 MB written, 71 MB read back) disappears from the serving path.
template <bool U8>
__global__ void __launch_bounds__(128) k_stem_tc(StemTcP p) {
    __shared__ __align__(1024) uint8_t a_tile[128 * 64];
    __shared__ __align__(1024) uint8_t b_tile[32 * 64];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[32];
    __shared__ unsigned long long optr[128];   // global address of every pixel's output row of the current tile (0: none)
    __shared__ float lut[U8 ? 256 : 1];
    const int t = threadIdx.x, warp = t >> 5;
    if constexpr (U8) { lut[t] = (float)((double)(float)t / 255.0); lut[t + 128] = (float)((double)(float)(t + 128) / 255.0); }
    const uint32_t a_addr = smem_u32(a_tile), b_addr = smem_u32(b_tile), bar = smem_u32(&mma_bar);
    if (t < 32) bias_s[t] = (t < p.nf) ? p.bias[t] : 0.f;
    {   // weights -> swizzled B tile (row f, 16-byte chunk j at f*64 + ((j ^ ((f>>1)&3)) << 4))
        const int f = t >> 2, j = t & 3;
        const uint4 v = *reinterpret_cast<const uint4 *>(p.w + f * 32 + j * 8);
        *reinterpret_cast<uint4 *>(b_tile + f * 64 + ((j ^ ((f >> 1) & 3)) << 4)) = v;
    }
    if (t == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // descriptors: K-major, 64-byte rows (SWIZZLE_64B = 4), SBO = 8 rows * 64 B, version 1
    const uint64_t hi = (uint64_t)(((8u * 64u) >> 4) | (1u << 14) | (4u << 29)) << 32;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t parity = 0;
    const size_t plane = (size_t)p.H * p.W;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const unsigned pix = (unsigned)tile * 128u + (unsigned)t;     // npix < 2^31 (checked by the plan): 32-bit divisions
        const bool ok = pix < (unsigned)p.npix;
        int x = 0, y = 0, n = 0;
        if (ok) { const unsigned row = pix / (unsigned)p.W; x = (int)(pix - row * (unsigned)p.W); n = (int)(row / (unsigned)p.H); y = (int)(row - (unsigned)n * (unsigned)p.H); }
        // ---- gather 27 taps (k = (ky*3 + kx)*3 + c), pad to 32, as bf16
        uint32_t packed[16];
        {
            float v[32];
            if constexpr (U8) {
                const unsigned char *img8 = p.in8 + (size_t)n * 3 * plane;
                if (ok && x >= 1 && x + 1 < p.W && y >= 1 && y + 1 < p.H) {
                    const unsigned char *r1 = img8 + ((size_t)y * p.W + x) * 3, *r0 = r1 - (size_t)p.W * 3, *r2 = r1 + (size_t)p.W * 3;
#pragma unroll
                    for (int j = 0; j < 9; ++j) {          // j = kx * 3 + c: nine consecutive bytes per image row
                        v[0 * 9 + j] = lut[__ldg(r0 - 3 + j)];
                        v[1 * 9 + j] = lut[__ldg(r1 - 3 + j)];
                        v[2 * 9 + j] = lut[__ldg(r2 - 3 + j)];
                    }
                } else {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const int iy = y + ky - 1, ix = x + kx - 1;
                            const bool in_img = ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
#pragma unroll
                            for (int c = 0; c < 3; ++c)
                                v[(ky * 3 + kx) * 3 + c] = in_img ? lut[__ldg(img8 + ((size_t)iy * p.W + ix) * 3 + c)] : 0.f;
                        }
                }
            } else {
            const float *img = p.in + (size_t)n * 3 * plane;
            if (ok && x >= 1 && x + 1 < p.W && y >= 1 && y + 1 < p.H) {
                // interior pixel (all but the image frame): nine row pointers, immediate offsets -1 / 0 / +1 -- the bounds-checked
                // form below costs ~10 integer instructions per tap and made this kernel issue-bound (ncu: 627 instructions per
                // warp and tile at 54 % issue utilisation, profiles/r02_notes.md)
                const float *c0 = img + (size_t)y * p.W + x;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float *r1 = c0 + (size_t)c * plane, *r0 = r1 - p.W, *r2 = r1 + p.W;
                    v[(0 * 3 + 0) * 3 + c] = __ldg(r0 - 1); v[(0 * 3 + 1) * 3 + c] = __ldg(r0); v[(0 * 3 + 2) * 3 + c] = __ldg(r0 + 1);
                    v[(1 * 3 + 0) * 3 + c] = __ldg(r1 - 1); v[(1 * 3 + 1) * 3 + c] = __ldg(r1); v[(1 * 3 + 2) * 3 + c] = __ldg(r1 + 1);
                    v[(2 * 3 + 0) * 3 + c] = __ldg(r2 - 1); v[(2 * 3 + 1) * 3 + c] = __ldg(r2); v[(2 * 3 + 2) * 3 + c] = __ldg(r2 + 1);
                }
            } else {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int iy = y + ky - 1, ix = x + kx - 1;
                    const bool in_img = ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        v[(ky * 3 + kx) * 3 + c] = in_img ? __ldg(img + (size_t)c * plane + (size_t)iy * p.W + ix) : 0.f;
                }
            }
            }
#pragma unroll
            for (int k = 27; k < 32; ++k) v[k] = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) packed[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4 *>(a_tile + t * 64 + ((j ^ ((t >> 1) & 3)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint64_t adesc = hi | (uint64_t)((((a_addr + 32u * k) & 0x3FFFFu) >> 4) | (1u << 16));
                const uint64_t bdesc = hi | (uint64_t)((((b_addr + 32u * k) & 0x3FFFFu) >> 4) | (1u << 16));
                umma_bf16(tmem_base, adesc, bdesc, idesc, (uint32_t)(k != 0));
            }
            umma_commit(bar);
        }
        mbar_wait(bar, parity, 9);
        parity ^= 1u;
        tc_fence_after();
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
        tmem_ld_wait();
        char *orow = ok ? p.out + ((size_t)(n * p.OHp + y + 1) * p.OWp + x + 1) * (size_t)p.out_ldc * 2 : nullptr;
        if (p.nf == 32) {
            // 64-byte pixel rows: stage the tile in shared memory (the A tile is free once the MMA has retired) and write it out
            // 512 contiguous bytes per warp instruction.  One STG.128 per thread on its own row touched 16 lines per instruction:
            // the L1 store wavefronts, not HBM, bounded this kernel (ncu: l1tex 74 % busy at 2.2 TB/s, profiles/r02_notes.md).
            uint4 o[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float a = __uint_as_float(acc[g * 8 + j]) + bias_s[g * 8 + j];
                    r[j] = (p.act == ACT_LEAKY) ? fmaxf(a, 0.1f * a) : a;
                }
                o[g] = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4 *>(a_tile + t * 64 + ((g ^ ((t >> 1) & 3)) << 4)) = o[g];
            optr[t] = (unsigned long long)(uintptr_t)orow;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = t + 128 * k, px = j >> 2, part = j & 3;
                const unsigned long long dst = optr[px];
                if (dst)
                    *reinterpret_cast<uint4 *>(dst + (unsigned long long)(part * 16)) =
                        *reinterpret_cast<const uint4 *>(a_tile + px * 64 + ((part ^ ((px >> 1) & 3)) << 4));
            }
        } else if (ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g * 8 >= p.nf) break;
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float a = __uint_as_float(acc[g * 8 + j]) + bias_s[g * 8 + j];
                    r[j] = (p.act == ACT_LEAKY) ? fmaxf(a, 0.1f * a) : a;
                }
                reinterpret_cast<uint4 *>(orow)[g] = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]),
                                                                pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
            }
        }
        tc_fence_before();
        __syncthreads();   // TMEM row drained and A tile consumed before the next tile overwrites them
    }
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u) : "memory");
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
    CUtensorMap tmA, tmB, tmO, tmR;   // activation, filters; TMA epilogue: output, residual
    TcParams p;
    int grid;
    size_t smem;
    int pdl;
    char desc[96];
};

int pick_bk(int C) { return (C % 64 == 0) ? 64 : (C % 32 == 0) ? 32 : (C % 16 == 0) ? 16 : 0; }
int pick_bk_i8(int cpad) { return (cpad % 128 == 0) ? 128 : (cpad % 64 == 0) ? 64 : (cpad % 32 == 0) ? 32 : 0; }
int pick_bk_f32(int C) { return (C % 32 == 0) ? 32 : (C % 16 == 0) ? 16 : (C % 8 == 0) ? 8 : 0; }
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

static void *make_plan_common(int kind, const Layer &l, const TV &in, const TV &out, bool out_bf16, const TV &res,
                              bool res_bf16, int act2, const void *d_weights_bf16, int ldn, const float *d_bias,
                              float alpha1, int *acc_out, int wide_rows = 0, int no_halo = 0, int want_halo = 0) {
    TcPlan *plan = new TcPlan();
    memset(plan, 0, sizeof(*plan));
    TcParams &p = plan->p;
    const bool i8 = kind == 1 || kind == 2;
    const int esz = kind == 3 ? 4 : i8 ? 1 : 2;              // operand element size
    const int cin = i8 ? in.ldc : l.c;                       // s8: channels padded with zeros in both operands
    // stride-2 parity halo (TcParams::halo == 2): bf16, C <= 64 -- the layers whose per-tap activation boxes kept the TMA unit,
    // not the tensor pipe, busy.  BK = 32 keeps three activation stages (4 planes x 9 x 17 lines each) inside shared memory.
    const int s2h_maxc = getenv("YB_TC_S2_HALO_MAXC") ? atoi(getenv("YB_TC_S2_HALO_MAXC")) : 64;
    const bool s2halo = kind == 0 && l.stride == 2 && l.size == 3 && l.pad == 1 && l.c % 32 == 0 && l.c <= s2h_maxc && !no_halo &&
                        !getenv("YB_TC_NO_HALO");
    const int BK = kind == 3 ? pick_bk_f32(l.c) : i8 ? pick_bk_i8(cin) : (s2halo && l.c <= 64) ? 32 : pick_bk(l.c), BN = pick_bn(l.n);
    p.kind = kind; p.alpha1 = alpha1; p.acc_out = acc_out;
    p.kk = BK * esz / 32;
    const bool s2 = l.stride == 2;
    p.N = in.N;
    p.OH = l.out_h; p.OW = l.out_w; p.OHp = out.Hp; p.OWp = out.Wp;
    p.size = l.size; p.BK = BK; p.BN = BN;
    p.cblocks = cin / BK; p.kblocks = l.size * l.size * p.cblocks;
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
    if (wide_rows) {
        // the epilogue will scatter NCHW planes (fused [yolo]): a warp's 32 accumulator rows should be as few image-row
        // runs as possible, so take the widest tile whose padded work stays within 30 % of the minimum
        for (int tw = 64; tw > bestTW; tw /= 2) {
            const int th = 128 / tw;
            const double cost = (double)((p.OW + tw - 1) / tw) * tw * (double)((rows + th - 1) / th) * th;
            if (cost <= 1.3 * best) { bestTW = tw; break; }
        }
    }
    p.TW = bestTW; p.TH = 128 / bestTW;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    const char *cg_env = getenv("YB_TC_CG");
    const uint32_t row_bytes = (uint32_t)(BK * esz);
    p.cblocks = cin / BK; p.kblocks = l.size * l.size * p.cblocks;
    p.nt = (l.n + BN - 1) / BN;
    // everything that depends on the tile shape, for one candidate (halo or per-tap) -- returns the predicted kernel time
    // in cycles: waves x K-blocks x max(tensor / issue time, TMA time at the ~48 B/clk/SM a streaming kernel sustains)
    // BN = 128 halo layers as CTA pairs too (experiment): 0 never, 1 the stride-2 one, 2 all.  Measured: no gain on 64->128 s2 @304,
    // +25 % TIME on 64->128 s1 @152 (profiles/r02_notes.md) -- stays off.
    const int cg2_bn128 = getenv("YB_TC_CG2_BN128") ? atoi(getenv("YB_TC_CG2_BN128")) : 0;
    auto layout = [&](bool halo) -> double {
        p.halo = halo ? (s2 ? 2 : 1) : 0;
        if (halo) { p.TW = 8; p.TH = 16; } else { p.TW = bestTW; p.TH = 128 / bestTW; }
        p.TWlog2 = 0; while ((1 << p.TWlog2) < p.TW) ++p.TWlog2;
        p.jshift = (halo && !s2) ? 1 : 0;     // halo tiles start one merged row down: row 0 is a border row, and 2x2 windows then never straddle tiles
        p.xt = (p.OW + p.TW - 1) / p.TW;
        p.jt = (int)((rows - p.jshift + p.TH - 1) / p.TH);
        p.num_tiles = p.xt * p.jt * p.nt;
        // CTA pairs (cta_group::2) for the wide tiles: halves the B bytes each SM has to pull through its TMA unit
        // (and for the BN = 128 stride-2 halo layer, whose 18 filter tiles per pixel tile otherwise outweigh the activation bytes)
        p.cg = ((kind == 0 || ((kind == 1 || kind == 2) && !getenv("YB_TC_I8_CG1"))) &&
                (BN == 256 || (BN == 128 && halo && cg2_bn128 >= (s2 ? 1 : 2))) && p.xt * p.jt >= 2 &&
                !(cg_env && atoi(cg_env) == 1)) ? 2 : 1;
        p.num_work = (p.cg == 2) ? ((p.xt * p.jt + 1) / 2) * p.nt : p.num_tiles;
        p.a_bytes = (uint32_t)(TC_BM * BK * esz);
        p.b_bytes = (uint32_t)((BN / p.cg) * BK * esz);   // per CTA
        // small filter matrices stay resident in shared memory for the whole kernel (one TMA pass per CTA)
        p.bstat = (p.cg == 1 && p.nt == 1 && (size_t)p.kblocks * p.b_bytes <= 72 * 1024 && !getenv("YB_TC_NO_BSTAT")) ? 1 : 0;
        p.bstat_bytes = p.bstat ? (uint32_t)p.kblocks * p.b_bytes : 0u;
        p.halo_pitch = (uint32_t)(p.TW + (s2 ? 1 : 2)) * row_bytes;
        p.halo_bytes = !halo ? 0u : s2 ? 4u * (uint32_t)(p.TH + 1) * p.halo_pitch : (uint32_t)(p.TH + 2) * p.halo_pitch;   // TMA bytes per tile and channel block
        const double mma = (double)p.kk * std::max(128.0 * BN / 256.0, 60.0);       // per K-block: tensor time vs single-thread issue
        const double a_per_kb = halo ? (double)p.halo_bytes / 9.0 : (double)p.a_bytes;
        const double tma = (a_per_kb + (p.bstat ? 0.0 : (double)p.b_bytes)) / 48.0;
        const int G = sms / p.cg;
        const double waves = (double)((p.num_work + G - 1) / G);
        return waves * p.kblocks * std::max(mma, tma);
    };
    // TMA epilogue: bf16 NHWC output of a stride-1 layer whose column halves are whole 64-column slabs
    // slab width 64 (two 16 KB tiles per warp group) where the epilogue is on the critical path (few K-blocks per tile); 32
    // (8 KB tiles: no more shared memory than the LSU staging, so the rings stay deep) for the deep-K and the BN = 64 layers
    p.tma_epi = 0;
    const bool i8kind = kind == 1 || kind == 2;
    // (stride-2 layers keep the LSU epilogue: their tiles walk the input's merged half-rows, OH + 1 per image, while the output has
    // OH + 2 rows per image -- a per-image (c, x, y, image) store would need a negative start row for the second image of a
    // straddling tile, and bulk tensor STORES fault on negative coordinates: tools/probes/tma4d_probe.cu, profiles/r02_tma4d_probe.txt)
    if (kind == 0 && out_bf16 && !s2 && BN >= 32 && !no_halo && !getenv("YB_TC_NO_TMA_EPI") &&
        !getenv("YB_TC_NO_COALESCE") && (!res.base || res_bf16)) {
        p.tma_epi = (BN >= 128 && p.kblocks <= 24) ? 64 : 32;   // measured per layer class (profiles/r02_notes.md)
        if (getenv("YB_TC_TMA_EPI_SW")) p.tma_epi = (atoi(getenv("YB_TC_TMA_EPI_SW")) == 64 && BN >= 128) ? 64 : 32;
    }
    // integer kinds: f32 slabs of 32 columns (128-byte rows) stored by TMA; the raw-accumulator dump (tests) keeps the LSU path
    if (i8kind && !s2 && !acc_out && !getenv("YB_TC_NO_TMA_EPI") && !getenv("YB_TC_NO_COALESCE")) p.tma_epi = 32;
    // alternate tiles per group: -24 % on 32->64 3x3 @304, -17 % on the BN = 64 1x1 layers, but +5 % on BN = 128 (measured);
    // BN = 32 tiles have a single 32-column slab, so alternating is the only way to use both groups at all
    // (and the BN = 64 stride-2 layer on the LSU epilogue: epilogue-bound once its loads are halo tiles -- 150 k of 176 k cycles busy)
    p.epi_alt = ((BN == 32 && (kind == 0 || i8kind)) || (p.tma_epi && (BN <= 64 || getenv("YB_TC_EPI_ALT_128")) && BN <= 128) ||
                 (kind == 0 && s2 && BN == 64 && out_bf16 && !getenv("YB_TC_NO_EPI_ALT_S2"))) &&
                !getenv("YB_TC_NO_EPI_ALT") ? 1 : 0;
    // (experiment, YB_TC_EPI_BUFS=2: two OUT tiles per group.  Measured: no layer gains, the layers that lose ring stages slow down
    // by 11-15 % -- the store's read latency is not what the slab loop waits for; profiles/r02_notes.md)
    p.epi_bufs = (getenv("YB_TC_EPI_BUFS") && p.tma_epi && atoi(getenv("YB_TC_EPI_BUFS")) == 2) ? 2 : 1;
    const size_t epi_tile = i8kind ? 16384 : (size_t)128 * p.tma_epi * 2;
    const size_t epi_tiles_bytes = !p.tma_epi ? 0 : i8kind ? 2 * p.epi_bufs * epi_tile : 2 * (p.epi_bufs + 1) * epi_tile;
    // what is left of 227 KB beside the epilogue tiles (5 KB: barriers, bias, alignment slack)
    const size_t ring_budget = p.tma_epi ? (size_t)(227 - 5) * 1024 - std::max<size_t>(epi_tiles_bytes, 32 * 1024) : (size_t)191 * 1024;
    bool use_halo = false;
    if (l.size == 3 && l.stride == 1 && l.pad == 1 && !no_halo && !getenv("YB_TC_NO_HALO") && (kind == 0 || ((kind == 1 || kind == 2) && !getenv("YB_TC_I8_NO_HALO")))) {
        const double t_halo = layout(true), t_tap = layout(false);
        use_halo = t_halo < 0.97 * t_tap || want_halo || getenv("YB_TC_HALO") != nullptr;   // want_halo: a fused max-pool needs the 8 x 16 tiles
    }
    if (s2halo) use_halo = true;
    layout(use_halo);
    // several K-blocks per stage when they are small: the single MMA-issuing thread pays a fixed barrier round
    // trip per stage, which dominated the C<=64 layers (profiles/r01_notes.md)
    const uint32_t sps_target = getenv("YB_TC_SPS_TARGET") ? (uint32_t)atoi(getenv("YB_TC_SPS_TARGET")) : 64u * 1024u;
    const uint32_t ring_blk = (p.halo ? 0u : p.a_bytes) + (p.bstat ? 0u : p.b_bytes);
    p.sps = (int)std::max<uint32_t>(1, std::min<uint32_t>(4, sps_target / std::max(ring_blk, 1u)));
    if (p.cg == 2) p.sps = 1;   // CTA pairs: 32 KB stages, 6 deep -- finer stages beat fewer barrier round trips here
    if (getenv("YB_TC_SPS")) p.sps = std::max(1, atoi(getenv("YB_TC_SPS")));
    if (getenv("YB_TC_SPS_CG2") && p.cg == 2) p.sps = std::max(1, atoi(getenv("YB_TC_SPS_CG2")));
    p.sps = std::min(p.sps, p.kblocks);
    if (p.halo) p.sps = 1;      // the ring streams filter tiles only, one K-block per stage
    const size_t fixed_smem = sizeof(float) * (size_t)p.nt * BN;
    p.a_stages = 0; p.a_stage_bytes = 0;
    if (p.halo) {
        p.a_stage_bytes = p.halo == 2 ? 4u * ((p.halo_bytes / 4u + 1023u) & ~1023u) : (p.halo_bytes + 1023u) & ~1023u;
        const int want = getenv("YB_TC_ASTAGES") ? atoi(getenv("YB_TC_ASTAGES")) : 3;
        p.a_stages = std::max(2, std::min(TC_MAX_ASTAGES, want));
        // keep at least 4 filter stages beside the activation ring
        while (p.a_stages > 2 && !p.bstat && ring_budget - p.bstat_bytes - fixed_smem - (size_t)p.a_stages * p.a_stage_bytes < 4 * (size_t)p.b_bytes) --p.a_stages;
    }
    const size_t a_ring = (size_t)p.a_stages * p.a_stage_bytes;
    {   // keep the ring at least 3 stages deep
        const size_t avail = ring_budget - p.bstat_bytes - fixed_smem - a_ring;
        while (p.sps > 1 && avail / ((size_t)p.sps * ring_blk) < 3) --p.sps;
    }
    p.stage_bytes = (uint32_t)p.sps * ring_blk;
    p.kbs = (p.kblocks + p.sps - 1) / p.sps;
    p.sk_T = 0; p.sk_L = 1;
    const size_t max_stages = getenv("YB_TC_MAX_STAGES") ? (size_t)atoi(getenv("YB_TC_MAX_STAGES")) : 8;
    if (p.stage_bytes == 0) p.stages = 2;   // halo + resident filters: nothing streams through the ring
    else p.stages = (int)std::min<size_t>(max_stages, (ring_budget - p.bstat_bytes - fixed_smem - a_ring) / p.stage_bytes);
    if (p.stages < 2) fatal_throw("tc plan: tile does not fit shared memory");
    // UMMA instruction descriptor (kind::f16): D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * p.cg) >> 4) << 24);
    // kind::i8: D = s32 (2 at bit 4), A = B = signed 8 bit (1 at bits 7 / 10)
    if (i8) p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * p.cg) >> 4) << 24);
    // kind::tf32: D = f32, A = B = tf32 (format 2): the tensor core reads the f32 words in place
    if (kind == 3) p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    // smem descriptor high word: SBO (8 rows * row bytes) >> 4 at bits 32..45, version 1 at bit 46, swizzle at 61..63
    const uint32_t layout_type = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
    p.desc_hi = ((8u * row_bytes) >> 4) | (1u << 14) | (layout_type << 29);
    p.desc_hi_a = (p.halo_pitch >> 4) | (1u << 14) | (layout_type << 29);   // halo A operand: 8-row groups one tile line pitch apart
    p.out = out.base; p.out_ldc = out.ldc; p.out_bf16 = out_bf16 ? 1 : 0;
    p.n = l.n;
    p.n_store = out_bf16 ? l.n : std::min<int>((l.n + 3) / 4 * 4, out.ldc);
    if (!out_bf16 && (out.ldc % 4 != 0)) fatal_throw("tc plan: f32 output rows must be 16-byte aligned");
    p.res = res.base; p.res_ldc = res.ldc; p.res_bf16 = res_bf16 ? 1 : 0;
    if (res.base && (res.H != l.out_h || res.W != l.out_w || res.C != l.n)) fatal_throw("tc plan: residual shape mismatch");
    if (res.base && !res_bf16) fatal_throw("tc plan: residual must be bf16");
    if (res.base && res_bf16 && (res.ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(res.base) & 15))) fatal_throw("tc plan: residual alignment");
    p.bias = d_bias; p.act = l.activation; p.act2 = act2;
    p.dbg = getenv("YB_TC_DBG") ? atoi(getenv("YB_TC_DBG")) : 0;
    p.l2_hint = (kind == 0 && !getenv("YB_TC_NO_L2_HINT")) ? 1 : 0;
    p.no_coalesce = getenv("YB_TC_NO_COALESCE") ? 1 : 0;
    snprintf(plan->desc, sizeof(plan->desc), "%dx%dx%d -> n%d k%d s%d%s%s", l.c, l.h, l.w, l.n, l.size, l.stride, p.halo ? " halo" : "",
             p.tma_epi ? " tepi" : "");
    uint32_t cols = 32; while (cols < (uint32_t)(TC_ACC * BN)) cols *= 2;
    p.tmem_cols = cols;

    const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    EncodeTiledFn enc = encode_fn();
    const CUtensorMapDataType dtype = kind == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                    : i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r;
    if (!s2) {
        // activation view (c, x_padded, merged padded rows)
        cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)in.Wp, (cuuint64_t)in.N * in.Hp};
        cuuint64_t strides[2] = {(cuuint64_t)in.ldc * esz, (cuuint64_t)in.Wp * in.ldc * esz};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(p.halo ? p.TW + 2 : p.TW), (cuuint32_t)(p.halo ? p.TH + 2 : p.TH)};
        cuuint32_t es[3] = {1, 1, 1};
        r = enc(&plan->tmA, dtype, 3, in.base, dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // stride 2: (c, x parity, x half, y parity, merged y half)
        cuuint64_t dims[5] = {(cuuint64_t)cin, 2, (cuuint64_t)in.Wp / 2, 2, (cuuint64_t)in.N * in.Hp / 2};
        cuuint64_t strides[4] = {(cuuint64_t)in.ldc * esz, (cuuint64_t)in.ldc * 2 * esz, (cuuint64_t)in.Wp * in.ldc * esz,
                                 (cuuint64_t)in.Wp * in.ldc * 2 * esz};
        cuuint32_t box[5] = {(cuuint32_t)BK, 1, (cuuint32_t)(p.halo ? p.TW + 1 : p.TW), 1, (cuuint32_t)(p.halo ? p.TH + 1 : p.TH)};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        r = enc(&plan->tmA, dtype, 5, in.base, dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { delete plan; fatal_throw("cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r)); }
    {
        const cuuint64_t K = (cuuint64_t)l.size * l.size * cin;
        cuuint64_t dims[2] = {K, (cuuint64_t)ldn};
        cuuint64_t strides[1] = {K * esz};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(BN / p.cg)};   // CG=2: each CTA of the pair loads half of the filters
        cuuint32_t es[2] = {1, 1};
        r = enc(&plan->tmB, dtype, 2, const_cast<void *>(d_weights_bf16), dims, strides, box, es,
                CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete plan; fatal_throw("cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r)); }
    }
    plan->tmO = plan->tmA; plan->tmR = plan->tmA;   // valid placeholders when the TMA epilogue is off
    if (p.tma_epi) {
        const int oesz = out_bf16 ? 2 : 4;
        auto encode_px = [&](CUtensorMap *tm, const TV &t, const char *what) {
            // (channels, padded x, merged padded rows) of a padded-NHWC tensor (bf16, or f32 for the integer kinds); box = one
            // slab of a pixel tile
            cuuint64_t dims[3] = {(cuuint64_t)l.n, (cuuint64_t)t.Wp, (cuuint64_t)t.N * t.Hp};
            cuuint64_t strides[2] = {(cuuint64_t)t.ldc * oesz, (cuuint64_t)t.Wp * t.ldc * oesz};
            cuuint32_t box[3] = {(cuuint32_t)p.tma_epi, (cuuint32_t)p.TW, (cuuint32_t)p.TH};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult rr = enc(tm, out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, t.base, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, p.tma_epi * oesz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rr != CUDA_SUCCESS) { delete plan; fatal_throw(std::string("cuTensorMapEncodeTiled(") + what + ") failed: " + std::to_string((int)rr)); }
        };
        encode_px(&plan->tmO, out, "output");
        if (res.base && out_bf16) encode_px(&plan->tmR, res, "residual");
    }
    plan->pdl = (getenv("YB_NO_PDL") == nullptr) ? 1 : 0;
    plan->grid = (p.cg == 2) ? 2 * std::min(p.num_work, sms / 2) : std::min(p.num_tiles, sms);
    if (getenv("YB_TC_STATS")) {
        cudaMalloc(&p.stats, sizeof(unsigned long long) * 16 * plan->grid);
        cudaMemset(p.stats, 0, sizeof(unsigned long long) * 16 * plan->grid);
    }
    plan->smem = (size_t)p.stages * p.stage_bytes + a_ring + p.bstat_bytes + 1024 /*alignment slack*/ +
                 8 * (2 * p.stages + 2 * TC_ACC + 1 + 2 * TC_MAX_ASTAGES + 2) + 16 +
                 sizeof(float) * (size_t)p.nt * BN /*bias*/ + (size_t)p.nt * BN / 8 /*yolo mask*/ +
                 (p.tma_epi ? 1024 + std::max<size_t>(epi_tiles_bytes, 4 * (size_t)(128 * p.tma_epi * 2)) /*TMA epilogue: [OUT x bufs | RES] per warp group*/
                            : 128 + 4096 * TC_EPI_WARPS /*epilogue staging*/);
    if (plan->smem > 227 * 1024) { delete plan; fatal_throw("tc plan: shared memory budget exceeded"); }
    {
        const void *fns[] = {(const void *)k_conv_tc<1, false, false, 0>, (const void *)k_conv_tc<2, false, false, 0>,
                             (const void *)k_conv_tc<1, false, false, 1>, (const void *)k_conv_tc<2, false, false, 1>,
                             (const void *)k_conv_tc<1, false, false, 2>, (const void *)k_conv_tc<2, false, false, 2>,
                             (const void *)k_conv_tc<2, false, true, 2>,
                             (const void *)k_conv_tc<1, true, false, 0>, (const void *)k_conv_tc<2, true, false, 0>,
                             (const void *)k_conv_tc<1, false, true, 0>, (const void *)k_conv_tc<2, false, true, 0>,
                             (const void *)k_conv_tc<1, false, true, 1>, (const void *)k_conv_tc<2, false, true, 1>,
                             (const void *)k_conv_tc<1, false, true, 2>};
        for (const void *f : fns)
            if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
                fatal_throw("cudaFuncSetAttribute(k_conv_tc) failed");
    }
    return plan;
}

void *tc_make_plan(const Layer &l, const TV &in, const TV &out, bool out_bf16, const TV &res, bool res_bf16,
                   int act2, const void *d_weights_bf16, int ldn, const float *d_bias, int wide_rows, int no_halo) {
    return make_plan_common(0, l, in, out, out_bf16, res, res_bf16, act2, d_weights_bf16, ldn, d_bias, 0.f, nullptr, wide_rows, no_halo);
}

// FP32 convolution of the exact (INT8 / XNOR) networks on kind::tf32: f32 NHWC activations and f32 [ldn][K] weights go
// through TMA untouched, the tensor core uses the upper 19 bits of each word (10-bit mantissa, ~3e-4 relative on the
// sums).  Used for detection heads only (their error cannot reach an integer layer), f32 output.
int tc_tf32_supported(const Layer &l, const TV &in, const TV &out) {
    if (pick_bk_f32(l.c) == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(in.base) & 15) != 0 || in.P != 1 || in.ldc % 4 != 0) return 0;
    const bool s1 = l.stride == 1 && ((l.size == 3 && l.pad == 1) || (l.size == 1 && l.pad == 0));
    const bool s2 = l.stride == 2 && l.size == 3 && l.pad == 1 && (l.h % 2 == 0) && (l.w % 2 == 0);
    if (!s1 && !s2) return 0;
    if (!out.base || (reinterpret_cast<uintptr_t>(out.base) & 15) != 0 || out.ldc % 4 != 0 || l.n < 8) return 0;
    if (l.activation != YB_LEAKY && l.activation != YB_LINEAR) return 0;
    return 1;
}
void *tc_make_plan_tf32(const Layer &l, const TV &in, const TV &out, const void *d_weights_f32, int ldn, const float *d_bias,
                        int wide_rows) {
    TV none{};
    return make_plan_common(3, l, in, out, false, none, false, ACT_LINEAR, d_weights_f32, ldn, d_bias, 0.f, nullptr, wide_rows);
}

// INT8 variant (reference forward_convolutional_layer_q, yolov2_forward_network_quantized.c:527-631) on
// tcgen05.mma kind::i8: `q` is the quantised s8 activation (padded NHWC, channels zero-padded to q.ldc), weights are
// s8 [ldn][taps][q.ldc]; output f32.
int tc_i8_supported(const Layer &l, const TV &q, const TV &out) {
    if (pick_bk_i8(q.ldc) == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(q.base) & 15) != 0 || q.P != 1 || q.ldc % 16 != 0) return 0;
    const bool s1 = l.stride == 1 && ((l.size == 3 && l.pad == 1) || (l.size == 1 && l.pad == 0));
    const bool s2 = l.stride == 2 && l.size == 3 && l.pad == 1 && (l.h % 2 == 0) && (l.w % 2 == 0);
    if (!s1 && !s2) return 0;
    if (!out.base || (reinterpret_cast<uintptr_t>(out.base) & 15) != 0 || out.ldc % 4 != 0 || l.n < 8) return 0;
    if (l.activation != YB_LEAKY && l.activation != YB_LINEAR) return 0;
    return 1;
}
void tc_plan_fuse_yolo(void *vp, float *d_yolo_nchw, int classes) {
    TcPlan *plan = reinterpret_cast<TcPlan *>(vp);
    plan->p.yolo_out = d_yolo_nchw;
    plan->p.yolo_per = 4 + classes + 1;
}

// Fuse the following 2x2 / stride-2 max-pool and the next integer layer's input conversion into an integer-kind halo plan:
// mode 1 = s8 quantised with `mult` (INT8 layer next), 2 = +-1 bytes (XNOR layer on the tensor cores next).  `qnext` is that layer's
// s8 input (padded NHWC).  Returns 0 (and changes nothing) when the plan's tiling cannot do it.
int tc_plan_fuse_pool(void *vp, int mode, float mult, const TV &qnext) {
    TcPlan *plan = reinterpret_cast<TcPlan *>(vp);
    TcParams &p = plan->p;
    if (!(p.kind == 1 || p.kind == 2) || !p.halo || p.TW != 8 || p.jshift != 1 || (p.PR & 1) || (p.OW & 1) || (p.OH & 1) || p.acc_out) return 0;
    if (qnext.H != p.OH / 2 || qnext.W != p.OW / 2 || qnext.ldc % 16 != 0 || (reinterpret_cast<uintptr_t>(qnext.base) & 15) || p.n % 32 != 0) return 0;
    p.pool_mode = mode; p.pool_mult = mult;
    p.pool_out = reinterpret_cast<signed char *>(qnext.base); p.pool_ldc = qnext.ldc; p.pool_Hp = qnext.Hp; p.pool_Wp = qnext.Wp;
    return 1;
}

void *tc_make_plan_i8(const Layer &l, const TV &q, const TV &out, const void *d_weights_s8, int ldn, const float *d_bias,
                      float alpha1, int *acc_out, int want_halo) {
    TV none{};
    return make_plan_common(1, l, q, out, false, none, false, ACT_LINEAR, d_weights_s8, ldn, d_bias, alpha1, acc_out, 0, 0, want_halo);
}

// XNOR layer mapped onto kind::i8: activations and weights as +-1 bytes, so the s32 accumulator is 2*count - K.
void *tc_make_plan_xnor(const Layer &l, const TV &q, const TV &out, const void *d_weights_pm1, int ldn, const float *d_bias,
                        const float *d_mean, int *counts_out, int want_halo) {
    TV none{};
    TcPlan *plan = reinterpret_cast<TcPlan *>(
        make_plan_common(2, l, q, out, false, none, false, ACT_LINEAR, d_weights_pm1, ldn, d_bias, 0.f, counts_out, 0, 0, want_halo));
    plan->p.mean = d_mean;
    plan->p.xK = l.size * l.size * l.c;
    return plan;
}

// K-split of the tail wave (see TcParams::sk_T).  With G CTAs (pairs) and num_work = R*G + T work items, the plain
// schedule costs R+1 waves; cutting the T tail items along K into G equal slices costs R + L/kbs waves plus the
// partial-sum round trip through L2 (hidden behind the next work item's main loop when R > 0).
// `ws`: sms * 128 KB, `flags`: sms * 8 words, zero-initialised, owned by the caller (one per stream of execution).
size_t tc_ksplit_ws_bytes(int sms) { return (size_t)sms * TC_BM * 256 * sizeof(float); }
size_t tc_ksplit_flag_bytes(int sms) { return (size_t)sms * TC_EPI_WARPS * sizeof(unsigned); }
int tc_plan_enable_ksplit(void *vp, float *ws, unsigned *flags) {
    TcPlan *plan = reinterpret_cast<TcPlan *>(vp);
    TcParams &p = plan->p;
    const char *ev = getenv("YB_TC_KSPLIT");                 // 0: never (even when the option asks for it)
    if (ev && ev[0] == '0') return 0;
    if (p.kind != 0 || !ws || !flags || p.halo) return 0;   // (the halo schedule has no K-split tail)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int G = sms / p.cg;
    const int R = p.num_work / G, T = p.num_work % G;
    if (T == 0) return 0;
    // Only deep-K layers: with few stages per work item the epilogue, not the tensor pipe, bounds the tile, and the
    // partial-sum round trip (one more accumulator read + ~2 us of L2 latency per partial) costs more than the wave saves.
    const int mink = getenv("YB_TC_KSPLIT_MINK") ? atoi(getenv("YB_TC_KSPLIT_MINK")) : 16;
    if (p.kbs < mink) return 0;
    const int pmax = getenv("YB_TC_KSPLIT_PIECES") ? std::max(1, atoi(getenv("YB_TC_KSPLIT_PIECES"))) : 3;
    const int lmin = std::max(2, (p.kbs + pmax - 1) / pmax);           // at most ~pmax slices per work item
    int L = std::max(lmin, (T * p.kbs + G - 1) / G);
    if (L >= p.kbs) return 0;
    // stage-times: exposed cost of the partial round trip (hidden behind the next item unless this is the only wave)
    // Predicted gain in stage-times.  The partial round trip costs ~6 stages, and the idle tail of the plain schedule is
    // not all waste: the next kernel's prologue runs in it (programmatic dependent launch).  Measured on yolov3-608
    // (profiles/r01_notes.md): layers below ~12 % predicted gain came out slower, the 19x19 3x3 layers 5-10 % faster.
    const double ovh = (R == 0) ? 10.0 : 6.0;
    const double before = (double)(R + 1) * p.kbs, after = (double)R * p.kbs + L + ovh;
    const double mingain = getenv("YB_TC_KSPLIT_MINGAIN") ? atof(getenv("YB_TC_KSPLIT_MINGAIN")) : 0.12;
    if (before - after < mingain * before) return 0;
    p.sk_T = T; p.sk_L = L; p.sk_ws = ws; p.sk_flags = flags;
    plan->grid = p.cg * G;
    return 1;
}

struct StemPlan { StemTcP p; int grid; };
int tc_stem_supported(const Layer &l, const TV &out) {
    return l.c == 3 && l.size == 3 && l.stride == 1 && l.pad == 1 && (l.n == 16 || l.n == 32) &&
           (l.activation == YB_LEAKY || l.activation == YB_LINEAR) && out.base && out.ldc % 8 == 0 &&
           (reinterpret_cast<uintptr_t>(out.base) & 15) == 0;
}
// d_w: device buffer of 32*32 bf16 ([filter][k], k = (ky,kx,c), zero padded), d_bias: device f32[>= n]
void *tc_stem_make_plan(const Layer &l, const TV &out, const void *d_w, const float *d_bias) {
    StemPlan *sp = new StemPlan();
    memset(sp, 0, sizeof(*sp));
    StemTcP &p = sp->p;
    p.out = out.base; p.out_ldc = out.ldc; p.w = reinterpret_cast<const __nv_bfloat16 *>(d_w); p.bias = d_bias;
    p.N = out.N; p.H = l.h; p.W = l.w; p.OHp = out.Hp; p.OWp = out.Wp; p.nf = l.n; p.act = l.activation;
    p.npix = (long)out.N * l.h * l.w;
    if (p.npix >= (1L << 31) - 256) fatal_throw("stem plan: more than 2^31 pixels per batch");
    p.ntiles = (int)((p.npix + 127) / 128);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sp->grid = std::min(p.ntiles, sms * 8);
    return sp;
}
void tc_stem_launch(void *vp, const float *d_in_nchw, cudaStream_t s) {
    StemPlan *sp = reinterpret_cast<StemPlan *>(vp);
    StemTcP p = sp->p;
    p.in = d_in_nchw;
    k_stem_tc<false><<<sp->grid, 128, 0, s>>>(p);
}
// 8-bit HWC frames of exactly the network size (3 channels): no planar-float staging
void tc_stem_launch_u8(void *vp, const unsigned char *d_in_hwc, cudaStream_t s) {
    StemPlan *sp = reinterpret_cast<StemPlan *>(vp);
    StemTcP p = sp->p;
    p.in8 = d_in_hwc;
    k_stem_tc<true><<<sp->grid, 128, 0, s>>>(p);
}
void tc_stem_free_plan(void *vp) { delete reinterpret_cast<StemPlan *>(vp); }

int tc_plan_cta_group(void *vp) { return reinterpret_cast<TcPlan *>(vp)->p.cg; }

void tc_launch(void *vp, cudaStream_t s) {
    TcPlan *plan = reinterpret_cast<TcPlan *>(vp);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)plan->grid); cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = plan->smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (plan->pdl) {   // let this kernel's prologue start while the previous kernel drains
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (plan->p.cg == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    static const bool ks_always = getenv("YB_TC_KS_ALWAYS") != nullptr;   // experiment: one kernel variant for every layer
    const bool ks = plan->p.sk_T > 0 || ks_always;
    const bool st = plan->p.stats != nullptr && !ks;   // role counters: a separate instantiation (YB_TC_STATS=1)
    const int epi = (plan->p.kind == 1 || plan->p.kind == 2) ? 2 : plan->p.tma_epi ? 1 : 0;
    const TcPlan &P = *plan;
#define YB_TC_LAUNCH(CG_, KS_, ST_, EPI_) cudaLaunchKernelEx(&cfg, k_conv_tc<CG_, KS_, ST_, EPI_>, P.tmA, P.tmB, P.tmO, P.tmR, P.p)
    if (plan->p.cg == 2) {
        if (ks) YB_TC_LAUNCH(2, true, false, 0);
        else if (epi == 1) { if (st) YB_TC_LAUNCH(2, false, true, 1); else YB_TC_LAUNCH(2, false, false, 1); }
        else if (epi == 2) { if (st) YB_TC_LAUNCH(2, false, true, 2); else YB_TC_LAUNCH(2, false, false, 2); }
        else { if (st) YB_TC_LAUNCH(2, false, true, 0); else YB_TC_LAUNCH(2, false, false, 0); }
    } else {
        if (ks) YB_TC_LAUNCH(1, true, false, 0);
        else if (epi == 1) { if (st) YB_TC_LAUNCH(1, false, true, 1); else YB_TC_LAUNCH(1, false, false, 1); }
        else if (epi == 2) { if (st) YB_TC_LAUNCH(1, false, true, 2); else YB_TC_LAUNCH(1, false, false, 2); }
        else { if (st) YB_TC_LAUNCH(1, false, true, 0); else YB_TC_LAUNCH(1, false, false, 0); }
    }
#undef YB_TC_LAUNCH
}

void tc_free_plan(void *vp) {
    TcPlan *plan = reinterpret_cast<TcPlan *>(vp);
    if (plan && plan->p.stats) {   // diagnostic dump: mean cycles per CTA of the LAST launch
        std::vector<unsigned long long> h(16 * (size_t)plan->grid);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), plan->p.stats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        double m[16] = {0};
        for (int b = 0; b < plan->grid; ++b) for (int k = 0; k < 16; ++k) m[k] += (double)h[16 * b + k] / plan->grid;
        fprintf(stderr, "TCSTATS %-28s cg %d tiles/cta %.1f kb %d sps %d BN %d ksplit T%d L%d/%d | producer: wait_empty %.0f tma_issue %.0f total %.0f | mma: wait_full %.0f "
                        "wait_tempty %.0f total %.0f | epi: wait_tfull %.0f wait_res %.0f total %.0f\n", plan->desc, plan->p.cg,
                (double)plan->p.num_tiles / plan->grid * 1.0, plan->p.kblocks, plan->p.sps, plan->p.BN, plan->p.sk_T, plan->p.sk_L, plan->p.kbs, m[0], m[7], m[1], m[2], m[3], m[4], m[5], m[8], m[6]);
        cudaFree(plan->p.stats);
    }
    delete plan;
}

}  // namespace yb
