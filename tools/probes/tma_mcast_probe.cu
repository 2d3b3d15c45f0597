// tma_mcast_probe.cu -- how fast can the chip's L2 feed all SMs through TMA, and does multicast help?
//
// profiles/r01_notes.md ("After v7", item 4): the big 3x3 layers of yolov3 need 64 B/clk/SM of L2 -> SM traffic for a
// 256x256x64 CTA-pair tile and sit at ~6.3 kB/clk chip-wide.  Before building a 2x2-CTA (cluster of 4) schedule with
// TMA multicast this probe answers, on the real machine:
//   1. how many clusters of 1 / 2 / 4 / 8 CTAs (one CTA per SM, ~200 KB of shared memory each) are co-resident, i.e. how
//      many of the 148 SMs a cluster-of-4 launch can use at all;
//   2. bytes/clk/SM delivered by TMA from an L2-resident buffer when every CTA streams
//        mode 0: its own tiles               (activations: unicast, no sharing)
//        mode 1: the same tiles as all others (weights: unicast, shared -- does L2 de-duplicate?)
//        mode 2: 1/CS of each shared tile, multicast to the CS CTAs of its cluster.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_mcast_probe tma_mcast_probe.cu
// Result of the first run: profiles/r01_tma_mcast_probe.txt (64 B/clk/SM unicast, 62 with multicast, 33 clusters of 4).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int NSTAGE = 4;
constexpr int TILE_ROWS = 256, ROW_BYTES = 128, TILE_BYTES = TILE_ROWS * ROW_BYTES;   // 32 KB, 128B-swizzled rows

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 2000000000LL) { printf("probe: mbarrier timeout (block %d)\n", blockIdx.x); __trap(); }
    }
}

// mode 0 / 1: tm_full (box = whole tile); mode 2: tm_part (box = TILE_ROWS / cs rows), multicast to the cluster
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ CUtensorMap tm_full, const __grid_constant__ CUtensorMap tm_part,
                                               int mode, int cs, int iters, int ntiles, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[NSTAGE];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t rank = cs > 1 ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (cs > 1) {   // every CTA's barriers exist before a peer multicasts into them
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int s = it % NSTAGE;
            const uint32_t bar = smem_u32(&full[s]);
            if (it >= NSTAGE) mbar_wait(bar, (uint32_t)(((it / NSTAGE) - 1) & 1));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TILE_BYTES) : "memory");
            const uint32_t dst = base + (uint32_t)s * TILE_BYTES;
            if (mode == 2) {
                const int part = TILE_ROWS / cs;
                const int tile = it % 64;
                const int row = tile * TILE_ROWS + (int)rank * part;
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                    ::"r"(dst + rank * (uint32_t)(part * ROW_BYTES)), "l"(&tm_part), "r"(bar), "r"(0), "r"(row),
                      "h"((uint16_t)((1u << cs) - 1u)) : "memory");
            } else {
                const int tile = mode == 1 ? it % 64 : (int)((blockIdx.x * 37u + (unsigned)it * 151u) % (unsigned)ntiles);
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                    ::"r"(dst), "l"(&tm_full), "r"(bar), "r"(0), "r"(tile * TILE_ROWS) : "memory");
            }
        }
        for (int it = iters < NSTAGE ? 0 : iters - NSTAGE; it < iters; ++it)      // drain
            mbar_wait(smem_u32(&full[it % NSTAGE]), (uint32_t)((it / NSTAGE) & 1));
        out[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    __syncthreads();
    if (cs > 1) {   // nobody leaves while a peer may still multicast into its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void *buf, uint64_t rows, uint32_t box_rows) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, rows};                       // 64 bf16 = 128 B per row
    cuuint64_t strides[1] = {ROW_BYTES};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return tm;
}

int main() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encode fn\n"); return 1; }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int ntiles = 2048;                                // 64 MB: L2-resident (126 MB L2)
    void *buf; cudaMalloc(&buf, (size_t)ntiles * TILE_BYTES); cudaMemset(buf, 1, (size_t)ntiles * TILE_BYTES);
    unsigned long long *d_out; cudaMalloc(&d_out, sizeof(unsigned long long) * 1024);
    const size_t smem = (size_t)NSTAGE * TILE_BYTES + 1024 + 72 * 1024;   // + ballast: one CTA per SM like k_conv_tc (~200 KB)
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    printf("SMs %d, clock %.0f MHz, %zu KB dynamic smem per CTA\n", sms, khz / 1e3, smem / 1024);
    for (int cs : {1, 2, 4, 8}) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(sms / cs * cs)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_probe, &cfg);
        printf("cluster size %d: max co-resident clusters %d (%d SMs usable)%s\n", cs, nclusters, nclusters * cs,
               e == cudaSuccess ? "" : "  [query failed]");
    }
    const int iters = 2000;
    printf("%6s %4s | %10s %12s %12s\n", "mode", "cs", "ms", "B/clk/SM", "TB/s chip");
    for (int mode : {0, 1, 2})
        for (int cs : {1, 2, 4}) {
            if (mode != 2 && cs != 1) continue;
            if (mode == 2 && cs == 1) continue;
            CUtensorMap tm_full = make_map(enc, buf, (uint64_t)ntiles * TILE_ROWS, TILE_ROWS);
            CUtensorMap tm_part = make_map(enc, buf, (uint64_t)ntiles * TILE_ROWS, TILE_ROWS / cs);
            cudaLaunchConfig_t cfg{};
            const int grid = sms / cs * cs;
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                cudaError_t le = cudaLaunchKernelEx(&cfg, k_probe, tm_full, tm_part, mode, cs, iters, ntiles, d_out);
                cudaEventRecord(e1);
                cudaError_t se = cudaDeviceSynchronize();
                if (le != cudaSuccess || se != cudaSuccess) { printf("mode %d cs %d: %s / %s\n", mode, cs, cudaGetErrorString(le), cudaGetErrorString(se)); return 1; }
                cudaEventElapsedTime(&ms, e0, e1);
            }
            unsigned long long h[1024];
            cudaMemcpy(h, d_out, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost);
            double cyc = 0; for (int b = 0; b < grid; ++b) cyc += (double)h[b] / grid;
            const double bytes_per_cta = (double)iters * TILE_BYTES;   // bytes LANDING in each CTA's shared memory
            printf("%6s %4d | %10.3f %12.1f %12.2f\n", mode == 0 ? "own" : mode == 1 ? "same" : "mcast", cs, ms,
                   bytes_per_cta / cyc, bytes_per_cta * grid / (ms * 1e-3) / 1e12);
        }
    return 0;
}
