// tma4d_probe.cu -- which (tensor dims, box, coordinates) does a 4-D cp.async.bulk.tensor STORE accept?  One case per process
// (an illegal-instruction fault kills the context).  Tensor = bf16 (c, x, y, img), box = (SW, TW, TH, 1), 128B / 64B swizzle.
// usage: tma4d_probe <C> <Wp> <Hp> <N> <SW> <TW> <TH> <x> <y> <img>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma4d_probe tma4d_probe.cu -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__global__ void k(const __grid_constant__ CUtensorMap tm, int bytes, int c0, int c1, int c2, int c3) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
    uint8_t *t = smem + (base - (uint32_t)__cvta_generic_to_shared(smem));
    for (int i = threadIdx.x; i < bytes / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(t)[i] = 0x3f80;   // bf16 1.0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(&tm), "r"(base), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main(int argc, char **argv) {
    if (argc < 11) return 2;
    const int C = atoi(argv[1]), Wp = atoi(argv[2]), Hp = atoi(argv[3]), N = atoi(argv[4]), SW = atoi(argv[5]), TW = atoi(argv[6]), TH = atoi(argv[7]);
    const int x = atoi(argv[8]), y = atoi(argv[9]), img = atoi(argv[10]);
    cudaFree(0);
    const size_t elems = (size_t)C * Wp * Hp * N;
    uint16_t *d; cudaMalloc(&d, elems * 2); cudaMemset(d, 0, elems * 2);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wp * C * 2, (cuuint64_t)Hp * Wp * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)SW, (cuuint32_t)TW, (cuuint32_t)TH, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        SW * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int bytes = SW * TW * TH * 2;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    k<<<1, 128, bytes + 1024>>>(tm, bytes, 0, x, y, img);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAULT: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> h(elems);
    cudaMemcpy(h.data(), d, elems * 2, cudaMemcpyDeviceToHost);
    long written = 0, expect = 0;
    for (int n = 0; n < N; ++n) for (int yy = 0; yy < Hp; ++yy) for (int xx = 0; xx < Wp; ++xx) for (int c = 0; c < C; ++c) {
        const bool in = n == img && yy >= y && yy < y + TH && xx >= x && xx < x + TW && c < SW;
        expect += in; written += h[(((size_t)n * Hp + yy) * Wp + xx) * C + c] == 0x3f80;
    }
    printf("ok: written %ld expected %ld %s\n", written, expect, written == expect ? "" : "MISMATCH");
    return 0;
}
