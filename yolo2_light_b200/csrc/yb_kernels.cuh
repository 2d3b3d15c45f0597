// yb_kernels.cuh -- CUDA-core (SIMT) kernels of the engine: layout converters, the generic FP32 / XNOR / INT8
// convolutions and all small layers.  sm_100a only.  The tensor-core (tcgen05) convolutions live in
// yb_conv_tc.cuh.
//
// Device activation layout ("padded NHWC"): element (n, y, x, c) of a tensor with logical dims N,H,W,C lives at
//     base + (((n*(H+2P) + y+P) * (W+2P) + x+P) * ldc + c) * sizeof(T)
// with a P=1 pixel border that is zero-filled once at allocation and never written -- 3x3/pad-1 convolutions
// (and their TMA loads) read the border instead of bounds-checking.  ldc >= C lets a tensor be a channel slice
// of a wider concat buffer (route layers become aliasing).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace yb {

struct TV {            // tensor view (POD, passed by value to kernels)
    char *base;        // address of channel 0 of padded pixel (n=0, y=-P, x=-P)
    int N, H, W, C;
    int ldc;           // elements per pixel in the underlying buffer
    int P;             // border
    int Hp, Wp;        // H+2P, W+2P
};

template <typename T>
__device__ __forceinline__ T *tv_px(const TV &t, int n, int y, int x) {
    return reinterpret_cast<T *>(t.base) + ((size_t)(n * t.Hp + y + t.P) * t.Wp + (x + t.P)) * (size_t)t.ldc;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

enum { ACT_LOGISTIC = 0, ACT_RELU = 1, ACT_LINEAR = 3, ACT_LEAKY = 7 };

// activate(), reference additionally.h:126-157 (scalar build): leaky = x>0 ? x : (float)(.1 * (double)x),
// logistic = (float)(1./(1.+exp(-x))) in double.  Used by the "exact" paths.
__device__ __forceinline__ float act_exact(float x, int a) {
    if (a == ACT_LEAKY) return (x > 0.f) ? x : (float)(0.1 * (double)x);
    if (a == ACT_LINEAR) return x;
    if (a == ACT_LOGISTIC) return (float)(1.0 / (1.0 + exp(-(double)x)));
    if (a == ACT_RELU) return x * (x > 0.f);
    return x;
}

// ------------------------------------------------------------------------------------------------------
// input: NCHW f32 (host contract, reference additionally.c:3093-3103) -> padded NHWC
// ------------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void k_input_nchw_to_nhwc(const float *__restrict__ in, TV out) {
    const long total = (long)out.N * out.H * out.W;
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < total; p += (long)gridDim.x * blockDim.x) {
        const int x = (int)(p % out.W);
        const int y = (int)((p / out.W) % out.H);
        const int n = (int)(p / ((long)out.W * out.H));
        TOut *o = tv_px<TOut>(out, n, y, x);
        const float *src = in + ((size_t)n * out.C * out.H + y) * out.W + x;
        for (int c = 0; c < out.C; ++c) o[c] = from_f32<TOut>(src[(size_t)c * out.H * out.W]);
    }
}

// any padded-NHWC activation -> NCHW f32 (diagnostic fetch)
template <typename TIn>
__global__ void k_nhwc_to_nchw_f32(TV in, float *__restrict__ out) {
    const long total = (long)in.N * in.C * in.H * in.W;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % in.W);
        const int y = (int)((i / in.W) % in.H);
        const int c = (int)((i / ((long)in.W * in.H)) % in.C);
        const int n = (int)(i / ((long)in.W * in.H * in.C));
        out[i] = to_f32(tv_px<TIn>(in, n, y, x)[c]);
    }
}

// ------------------------------------------------------------------------------------------------------
// input pipeline of the reference app on the device: u8 HWC image (what stbi_load returns) ->
// planar float /255. (load_image_stb, additionally.c:3080-3103) -> resize_image's two-pass bilinear to the network
// size (additionally.c:3021-3064), fused: one thread per output element recomputes the two x-interpolated values it
// needs.  Every product and sum is rounded to float exactly where the reference rounds (no FMA contraction), so the
// result is bit-identical to the reference's scalar build.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float u8_to_unit(unsigned char v) { return (float)((double)(float)v / 255.0); }

__device__ __forceinline__ float resize_part(const unsigned char *img, int w, int c, int k, int r, int cc, int out_w, float w_scale) {
    // value of the reference's `part` image at (cc, r, k)
    if (cc == out_w - 1 || w == 1) return u8_to_unit(img[k + c * (w - 1) + c * w * r]);
    const float sx = __fmul_rn((float)cc, w_scale);
    const int ix = (int)sx;
    const float dx = __fsub_rn(sx, (float)ix);
    const float a = u8_to_unit(img[k + c * ix + c * w * r]), b = u8_to_unit(img[k + c * (ix + 1) + c * w * r]);
    return __fadd_rn(__fmul_rn(__fsub_rn(1.f, dx), a), __fmul_rn(dx, b));
}

static __global__ void k_resize_u8_to_nchw(const unsigned char *__restrict__ src, int n_img, int w, int h, int c,
                                           float *__restrict__ dst, int out_w, int out_h) {
    const long total = (long)n_img * c * out_h * out_w;
    const bool same = (out_w == w && out_h == h);
    const float w_scale = (float)(w - 1) / (float)(out_w - 1);
    const float h_scale = (float)(h - 1) / (float)(out_h - 1);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cc = (int)(i % out_w);
        const int r = (int)((i / out_w) % out_h);
        const int k = (int)((i / ((long)out_w * out_h)) % c);
        const int n = (int)(i / ((long)out_w * out_h * c));
        const unsigned char *img = src + (size_t)n * w * h * c;
        float val;
        if (same) {
            val = u8_to_unit(img[k + c * cc + c * w * r]);
        } else {
            const float sy = __fmul_rn((float)r, h_scale);
            const int iy = (int)sy;
            const float dy = __fsub_rn(sy, (float)iy);
            val = __fmul_rn(__fsub_rn(1.f, dy), resize_part(img, w, c, k, iy, cc, out_w, w_scale));
            if (!(r == out_h - 1 || h == 1))
                val = __fadd_rn(val, __fmul_rn(dy, resize_part(img, w, c, k, iy + 1, cc, out_w, w_scale)));
        }
        dst[i] = val;
    }
}

// ------------------------------------------------------------------------------------------------------
// generic FP32 convolution on CUDA cores (implicit GEMM, 64 pixels x 64 filters per CTA, 4x4 per thread).
// Semantics: forward_convolutional_layer_cpu FP32 branch (reference yolov2_forward_network.c:204-261):
// out = act(sum_{c,ky,kx} w*in + bias) with zero padding; optional fused shortcut (reference :443-449):
// out = act2(act(..) + residual).  Used for: validation precision (YB_PREC_FP32), the FP32 layers of
// XNOR / INT8 networks, the 3-channel stem, and any shape the tensor-core kernel does not take.
// Weights: [K][ldw] f32, K ordered (ky, kx, c), ldw = filters rounded up to 64.
// ------------------------------------------------------------------------------------------------------
struct ConvP {
    TV in, out, res;           // res.base == nullptr: no residual
    const void *w;
    const float *bias;
    int n;                     // filters
    int ldw;
    int size, stride, pad;
    int act, act2;
    int K;                     // size*size*C
    long M;                    // N*out_h*out_w
};

// EXACT = true (f32 activations: the exact nets and YB_PREC_FP32): K runs in the reference's own order (c, ky, kx) --
// weights [K][ldw] stored in that order -- and every product and sum is rounded separately (__fmul_rn / __fadd_rn), i.e.
// gemm_nn's `C[j] += A_PART * B[k][j]` of the scalar build (additionally.c:1272-1286): the result is bit-identical to the
// reference, so a last-bit difference can never flip `x > 0` / `(int16)(x * m)` in a following integer layer.
template <typename TIn, typename TOut, typename TRes, bool EXACT = false>
__global__ void __launch_bounds__(256) k_conv_simt(ConvP p) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int OH = p.out.H, OW = p.out.W, C = p.in.C;

    // A-load role: thread -> (pixel tid/4, 4 consecutive k starting at (tid%4)*4)
    const int lp = tid >> 2, lk = (tid & 3) * 4;
    const long lm = m0 + lp;
    const bool lvalid = lm < p.M;
    int ln = 0, liy0 = 0, lix0 = 0;
    if (lvalid) {
        const int ox = (int)(lm % OW);
        const int oy = (int)((lm / OW) % OH);
        ln = (int)(lm / ((long)OW * OH));
        liy0 = oy * p.stride - p.pad;
        lix0 = ox * p.stride - p.pad;
    }
    // B-load role: thread -> (k row tid/16, 4 consecutive filters (tid%16)*4)
    const int bk = tid >> 4, bn = (tid & 15) * 4;
    const float *wptr = reinterpret_cast<const float *>(p.w);

    const int tx = tid & 15, ty = tid >> 4;   // compute role: pixels ty*4.., filters tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            float v = 0.f;
            if (lvalid && k < p.K) {
                const int taps = p.size * p.size;
                const int tap = EXACT ? k % taps : k / C, ch = EXACT ? k / taps : k - tap * C;
                const int ky = tap / p.size, kx = tap - ky * p.size;
                const int iy = liy0 + ky, ix = lix0 + kx;
                if (iy >= 0 && iy < p.in.H && ix >= 0 && ix < p.in.W) v = to_f32(tv_px<TIn>(p.in, ln, iy, ix)[ch]);
            }
            As[lk + q][lp] = v;
        }
        {
            const int k = k0 + bk;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < p.K) v = *reinterpret_cast<const float4 *>(wptr + (size_t)k * p.ldw + n0 + bn);
            Bs[bk][bn + 0] = v.x; Bs[bk][bn + 1] = v.y; Bs[bk][bn + 2] = v.z; Bs[bk][bn + 3] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if constexpr (EXACT) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(b[j], a[i]));
                    else acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        const int ox = (int)(m % OW);
        const int oy = (int)((m / OW) % OH);
        const int n = (int)(m / ((long)OW * OH));
        TOut *o = tv_px<TOut>(p.out, n, oy, ox);
        const TRes *r = p.res.base ? tv_px<TRes>(p.res, n, oy, ox) : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = n0 + tx * 4 + j;
            if (f >= p.n) continue;
            float v = act_exact(__fadd_rn(acc[i][j], p.bias[f]), p.act);
            if (r) v = act_exact(__fadd_rn(v, to_f32(r[f])), p.act2);
            o[f] = from_f32<TOut>(v);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// stem convolution (network input, 3 channels): 3x3 / stride 1 / pad 1 straight from the caller's NCHW f32 image
// (reference input contract additionally.c:3093-3103) to padded-NHWC output -- fuses the layout conversion, so
// the image is read exactly once and no NHWC copy of it is ever written.  One thread per output pixel, all NF
// filters in registers, weights [27][NF] broadcast from shared memory.  Accumulation order (ky, kx, c) matches
// k_conv_simt.
// ------------------------------------------------------------------------------------------------------
template <int NF>
struct StemW {            // passed by value as a kernel parameter: lives in the constant bank, so every FFMA takes
    float w[27 * NF];     // its weight operand straight from c[][] -- no shared-memory traffic at all
    float b[NF];
};

// EXACT (f32 output = the exact nets): reference order (c, ky, kx) with separately rounded products and sums, see
// k_conv_simt -- the INT8 / XNOR layers behind the stem then see bit-identical inputs.
template <int NF, typename TOut, bool EXACT = false>
__global__ void __launch_bounds__(128) k_conv_stem(const float *__restrict__ in, TV out, const __grid_constant__ StemW<NF> sw,
                                                   int act, int H, int W) {
    const long total = (long)out.N * H * W;
    const long p = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (p >= total) return;
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const int n = (int)(p / ((long)W * H));
    const float *img = in + (size_t)n * 3 * H * W;
    float acc[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[f] = 0.f;
    if constexpr (EXACT) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = y + ky - 1;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = x + kx - 1;
                    const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
                    const float v = ok ? __ldg(img + ((size_t)c * H + iy) * W + ix) : 0.f;
#pragma unroll
                    for (int f = 0; f < NF; ++f)
                        acc[f] = __fadd_rn(acc[f], __fmul_rn(sw.w[((ky * 3 + kx) * 3 + c) * NF + f], v));
                }
            }
    } else {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = x + kx - 1;
            const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = ok ? __ldg(img + ((size_t)c * H + iy) * W + ix) : 0.f;
#pragma unroll
                for (int f = 0; f < NF; ++f) acc[f] = fmaf(v, sw.w[((ky * 3 + kx) * 3 + c) * NF + f], acc[f]);
            }
        }
    }
    }
    TOut *o = tv_px<TOut>(out, n, y, x);
    if constexpr (sizeof(TOut) == 2) {
        uint4 *op = reinterpret_cast<uint4 *>(o);
#pragma unroll
        for (int g = 0; g < NF / 8; ++g) {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = acc[g * 8 + j] + sw.b[g * 8 + j];
                t[j] = (act == ACT_LEAKY) ? ((a > 0.f) ? a : 0.1f * a) : act_exact(a, act);
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(t[0], t[1]), h1 = __floats2bfloat162_rn(t[2], t[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(t[4], t[5]), h3 = __floats2bfloat162_rn(t[6], t[7]);
            uint4 v;
            v.x = *reinterpret_cast<uint32_t *>(&h0); v.y = *reinterpret_cast<uint32_t *>(&h1);
            v.z = *reinterpret_cast<uint32_t *>(&h2); v.w = *reinterpret_cast<uint32_t *>(&h3);
            op[g] = v;
        }
    } else {
        float4 *op = reinterpret_cast<float4 *>(o);   // NF*4 bytes per pixel, 16-byte aligned (ldc % 4 == 0)
#pragma unroll
        for (int g = 0; g < NF / 4; ++g)
            op[g] = make_float4(act_exact(__fadd_rn(acc[g * 4 + 0], sw.b[g * 4 + 0]), act), act_exact(__fadd_rn(acc[g * 4 + 1], sw.b[g * 4 + 1]), act),
                                act_exact(__fadd_rn(acc[g * 4 + 2], sw.b[g * 4 + 2]), act), act_exact(__fadd_rn(acc[g * 4 + 3], sw.b[g * 4 + 3]), act));
    }
}

// ------------------------------------------------------------------------------------------------------
// BIT1-XNOR path (reference yolov2_forward_network.c:116-203).
// k_binarize: f32 activation -> 1 bit per channel, bit = (x > 0), 32 channels per word, padded NHWC with a
// zero (== -1, F9) border.  One warp per (pixel, word): coalesced read + __ballot_sync.
// ------------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void k_binarize(TV in, TV bits /* C = words per pixel */) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int CW = bits.C;
    const long total = (long)in.N * in.H * in.W * CW;
    for (long i = warp; i < total; i += nwarps) {
        const int wd = (int)(i % CW);
        const long pxl = i / CW;
        const int x = (int)(pxl % in.W);
        const int y = (int)((pxl / in.W) % in.H);
        const int n = (int)(pxl / ((long)in.W * in.H));
        const int c = wd * 32 + lane;
        float v = 0.f;
        if (c < in.C) v = to_f32(tv_px<TIn>(in, n, y, x)[c]);
        const unsigned m = __ballot_sync(0xffffffffu, v > 0.f);
        if (lane == 0) tv_px<uint32_t>(bits, n, y, x)[wd] = m;
    }
}

// binarize_cpu (additionally.c:128-134): x > 0 ? +1 : -1 as floats -- the input of the XNOR layers that take the reference's
// float-GEMM fallback (stride != 1 or pad != 1).  The zero border of the destination stays zero (im2col's padding value).
static __global__ void k_binarize_pm1(TV in, TV out) {
    const long total = (long)in.N * in.H * in.W * in.C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % in.C);
        const long px = i / in.C;
        const int x = (int)(px % in.W), y = (int)((px / in.W) % in.H), n = (int)(px / ((long)in.W * in.H));
        tv_px<float>(out, n, y, x)[c] = tv_px<float>(in, n, y, x)[c] > 0.f ? 1.f : -1.f;
    }
}

struct XnorP {
    TV bits;                  // input bits, C = words per pixel (CW)
    TV out;                   // f32
    const uint32_t *w;        // [ldn filters][9 taps][CW] sign bits
    const float *mean, *bias;
    int n, size, pad;
    int K;                    // true bit count size*size*C
    int padbits;              // (CW*32 - C) * size*size
    int act;
    long M;
    int32_t *counts;          // optional raw popcounts, NCHW (tests)
};

// The same convolution with the following 2x2 / stride-2 max-pool and the next XNOR layer's input conversion fused in: one thread
// per POOLED pixel computes the four popcount outputs of its window for every filter, takes the reference's max (max-pool
// semantics: elements outside the image are skipped) and writes only the SIGN the next layer would have extracted -- MODE 2: +-1
// bytes (next layer runs as +-1 on kind::i8), MODE 3: sign bits (next layer on the popcount kernels).  The f32 activation and the
// pooled f32 tensor are never written.  Bit-identical to conv -> max-pool -> binarise.
template <int CW, int MODE>
__global__ void __launch_bounds__(128) k_conv_xnor_smallk_pool(XnorP p, TV q /* next layer's input: s8 (MODE 2) or bit words (MODE 3) */) {
    extern __shared__ uint32_t wsm[];            // [n][9*CW]
    constexpr int KW = 9 * CW;
    for (int i = threadIdx.x; i < p.n * KW; i += blockDim.x) wsm[i] = p.w[i];
    __syncthreads();
    const int H = p.out.H, W = p.out.W, PH = q.H, PW = q.W;
    const long m = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (m >= (long)q.N * PH * PW) return;
    const int px = (int)(m % PW), py = (int)((m / PW) % PH), n = (int)(m / ((long)PW * PH));
    uint32_t a[16 * CW];                          // 4x4 window of input words around the 2x2 outputs
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int iy = 2 * py + t / 4 - 1, ix = 2 * px + t % 4 - 1;
        const bool in = iy >= -1 && iy <= H && ix >= -1 && ix <= W;      // the 1-pixel border exists in memory (words 0 == -1)
        const uint32_t *src = tv_px<uint32_t>(p.bits, n, in ? iy : -1, in ? ix : -1);
#pragma unroll
        for (int c = 0; c < CW; ++c) a[t * CW + c] = in ? src[c] : 0u;
    }
    uint32_t word = 0;
    for (int f = 0; f < p.n; ++f) {
        const uint32_t *wf = wsm + f * KW;
        float mx = -3.402823466e+38f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {             // window order of the reference: rows, then columns
            const int dy = k >> 1, dx = k & 1;
            if (2 * py + dy >= H || 2 * px + dx >= W) continue;
            int cnt = 0;
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int c = 0; c < CW; ++c) cnt += __popc(~(a[((t / 3 + dy) * 4 + t % 3 + dx) * CW + c] ^ wf[t * CW + c]));
            const int count = cnt - p.padbits;
            float v = __fadd_rn(__fmul_rn((float)(2 * count - p.K), p.mean[f]), p.bias[f]);
            v = act_exact(v, p.act);
            mx = v > mx ? v : mx;
        }
        if (MODE == 2) {     // four +-1 bytes per store (the filter count of an XNOR layer feeding the tensor-core path is a multiple of 32)
            word |= (mx > 0.f ? 0x01u : 0xFFu) << (8 * (f & 3));
            if ((f & 3) == 3) { reinterpret_cast<uint32_t *>(tv_px<int8_t>(q, n, py, px))[f >> 2] = word; word = 0; }
        } else {
            if (mx > 0.f) word |= 1u << (f & 31);
            if ((f & 31) == 31 || f + 1 == p.n) { tv_px<uint32_t>(q, n, py, px)[f >> 5] = word; word = 0; }
        }
    }
}

// thread-per-(pixel, word) variant: each thread reads its 32 (or fewer) channels with 16-byte loads -- far fewer,
// fatter threads than the ballot version; used whenever the channel vector is 16-byte aligned.
template <typename TIn>
__global__ void k_binarize_vec(TV in, TV bits) {
    const int CW = bits.C;
    const long total = (long)in.N * in.H * in.W * CW;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int wd = (int)(i % CW);
        const long pxl = i / CW;
        const int x = (int)(pxl % in.W);
        const int y = (int)((pxl / in.W) % in.H);
        const int n = (int)(pxl / ((long)in.W * in.H));
        const float4 *src = reinterpret_cast<const float4 *>(tv_px<float>(in, n, y, x) + wd * 32);
        const int nch = min(32, in.C - wd * 32);
        uint32_t m = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q * 4 < nch) {
                const float4 v = __ldg(src + q);
                m |= (uint32_t)(v.x > 0.f) << (q * 4) | (uint32_t)(v.y > 0.f) << (q * 4 + 1) |
                     (uint32_t)(v.z > 0.f) << (q * 4 + 2) | (uint32_t)(v.w > 0.f) << (q * 4 + 3);
            }
        }
        tv_px<uint32_t>(bits, n, y, x)[wd] = m;
    }
}

// sign(+-1) as s8 for the tensor-core XNOR mapping: +1 where x > 0, -1 otherwise (border bytes are pre-set to -1,
// the reference's "out-of-image taps are -1" rule).  16 channels per thread.
static __global__ void k_binarize_s8(TV in, TV q) {
    const int groups = in.C >> 4;
    const long total = (long)in.N * in.H * in.W * groups;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long pxl = i / groups;
        const int x = (int)(pxl % in.W);
        const int y = (int)((pxl / in.W) % in.H);
        const int n = (int)(pxl / ((long)in.W * in.H));
        const float4 *src = reinterpret_cast<const float4 *>(tv_px<float>(in, n, y, x) + g * 16);
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = __ldg(src + k);
            w[k] = (v.x > 0.f ? 0x01u : 0xFFu) | (v.y > 0.f ? 0x01u : 0xFFu) << 8 | (v.z > 0.f ? 0x01u : 0xFFu) << 16 |
                   (v.w > 0.f ? 0x01u : 0xFFu) << 24;
        }
        reinterpret_cast<uint4 *>(tv_px<int8_t>(q, n, y, x))[g] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// XNOR convolution for small K (one or two words per tap): one thread per output pixel keeps its 9 x CW input
// words in registers and walks all filters, whose sign words sit in shared memory (broadcast reads).
template <int CW>
__global__ void __launch_bounds__(128) k_conv_xnor_smallk(XnorP p) {
    extern __shared__ uint32_t wsm[];            // [n][9*CW]
    constexpr int KW = 9 * CW;
    for (int i = threadIdx.x; i < p.n * KW; i += blockDim.x) wsm[i] = p.w[i];
    __syncthreads();
    const int H = p.out.H, W = p.out.W;
    const long m = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (m >= p.M) return;
    const int x = (int)(m % W);
    const int y = (int)((m / W) % H);
    const int n = (int)(m / ((long)W * H));
    uint32_t a[KW];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const uint32_t *src = tv_px<uint32_t>(p.bits, n, y + t / 3 - 1, x + t % 3 - 1);   // border words are 0 (== -1)
#pragma unroll
        for (int c = 0; c < CW; ++c) a[t * CW + c] = src[c];
    }
    float *o = tv_px<float>(p.out, n, y, x);
    for (int f0 = 0; f0 < p.n; f0 += 4) {
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = f0 + j;
            int cnt = 0;
            if (f < p.n) {
                const uint32_t *wf = wsm + f * KW;
#pragma unroll
                for (int k = 0; k < KW; ++k) cnt += __popc(~(a[k] ^ wf[k]));
            }
            const int count = cnt - p.padbits;
            if (p.counts && f < p.n) p.counts[(((size_t)n * p.n + f) * H + y) * W + x] = count;
            float v = (f < p.n) ? __fmul_rn((float)(2 * count - p.K), p.mean[f]) : 0.f;
            v = (f < p.n) ? __fadd_rn(v, p.bias[f]) : 0.f;
            r[j] = act_exact(v, p.act);
        }
        if (f0 + 3 < p.n) *reinterpret_cast<float4 *>(o + f0) = make_float4(r[0], r[1], r[2], r[3]);
        else for (int j = 0; j < 4 && f0 + j < p.n; ++j) o[f0 + j] = r[j];
    }
}

// XNOR bit-GEMM convolution, 3x3 / stride 1 / pad 1: count = sum_taps popc(~(a ^ w)) - (pad bits), then
// out = act((2*count - K) * mean[f] + bias[f]) evaluated in the reference's float op order
// (additionally.c:1531, yolov2_forward_network.c:243-261).  64 pixels x 64 filters per CTA, 4x4 per thread,
// K streamed through shared memory in 8-word slices with 128-bit loads.

static __global__ void __launch_bounds__(256) k_conv_xnor(XnorP p) {
    constexpr int BM = 64, BN = 64, BKW = 8;
    __shared__ uint32_t As[BKW][BM + 1];
    __shared__ uint32_t Bs[BKW][BN + 1];
    const int tid = threadIdx.x;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int H = p.out.H, W = p.out.W, CW = p.bits.C;
    const int taps = p.size * p.size;
    const int KW = taps * CW;

    // loader roles: 64 rows x 8 words = 512 words per operand, 2 per thread
    const int lr = tid >> 2, lw = (tid & 3) * 2;
    const long lm = m0 + lr;
    const bool lvalid = lm < p.M;
    int ln = 0, ly = 0, lx = 0;
    if (lvalid) {
        lx = (int)(lm % W);
        ly = (int)((lm / W) % H);
        ln = (int)(lm / ((long)W * H));
    }
    const int tx = tid & 15, ty = tid >> 4;
    int acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;

    for (int k0 = 0; k0 < KW; k0 += BKW) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = k0 + lw + q;
            uint32_t a = 0, b = 0;
            if (k < KW) {
                const int tap = k / CW, wd = k - tap * CW;
                if (lvalid) {
                    const int ky = tap / p.size, kx = tap - ky * p.size;
                    const int iy = ly + ky - p.pad, ix = lx + kx - p.pad;
                    // out-of-image taps read 0-bits (== -1): border for pad<=1, explicit otherwise
                    if (iy >= -p.bits.P && iy < H + p.bits.P && ix >= -p.bits.P && ix < W + p.bits.P)
                        a = tv_px<uint32_t>(p.bits, ln, iy, ix)[wd];
                }
                const int f = n0 + lr;
                b = p.w[(size_t)f * KW + k];
            } else {
                a = 0; b = 0xffffffffu;   // a ^ b = all ones -> xnor contributes 0
            }
            As[lw + q][lr] = a;
            Bs[lw + q][lr] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BKW; ++kk) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += __popc(~(a[i] ^ b[j]));
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        const int x = (int)(m % W);
        const int y = (int)((m / W) % H);
        const int n = (int)(m / ((long)W * H));
        float *o = tv_px<float>(p.out, n, y, x);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = n0 + tx * 4 + j;
            if (f >= p.n) continue;
            const int count = acc[i][j] - p.padbits;
            if (p.counts) p.counts[(((size_t)n * p.n + f) * H + y) * W + x] = count;
            float v = __fmul_rn((float)(2 * count - p.K), p.mean[f]);   // no fma contraction: one mul, one add
            v = __fadd_rn(v, p.bias[f]);
            o[f] = act_exact(v, p.act);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// INT8 path (reference yolov2_forward_network_quantized.c:527-631).
// k_quantize: xq = clamp(+-127, (int16_t)(x * input_mult)) with x86 float->int16 semantics (cvttss2si, low
// 16 bits, indefinite -> 0).  Output s8 padded NHWC, channels padded with zeros to ldc.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int quant_i8(float x, float mult) {
    const float v = __fmul_rn(x, mult);
    int i;
    if (!(v > -2147483648.0f && v < 2147483648.0f)) i = (int)0x80000000;
    else i = __float2int_rz(v);
    int s = (int)(short)(i & 0xffff);
    if (s > 127) s = 127;
    if (s < -127) s = -127;
    return s;
}

template <typename TIn>
__global__ void k_quantize(TV in, TV q /* s8, ldc multiple of 4 */, float mult) {
    const int groups = q.ldc >> 2;
    const long total = (long)in.N * in.H * in.W * groups;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long pxl = i / groups;
        const int x = (int)(pxl % in.W);
        const int y = (int)((pxl / in.W) % in.H);
        const int n = (int)(pxl / ((long)in.W * in.H));
        const TIn *src = tv_px<TIn>(in, n, y, x);
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = g * 4 + j;
            const int s = (c < in.C) ? quant_i8(to_f32(src[c]), mult) : 0;
            packed |= (uint32_t)(s & 0xff) << (8 * j);
        }
        reinterpret_cast<uint32_t *>(tv_px<int8_t>(q, n, y, x))[g] = packed;
    }
}

// INT8 convolution on CUDA cores (dp4a): acc32 = sum xq*wq (exact), then the reference's requantisation
// epilogue: q16 = clamp(+-32767, acc32 / 32) [C truncating division]; y = (float)q16 * alpha1; y += bias;
// leaky: y > 0 ? y : y / 10  (yolov2_forward_network_quantized.c:474-490, :598-627).
// Weights: [ldn filters][taps][ldc_in/4 words] s8, K ordered (ky, kx, c) with zero channel padding.
struct Int8P {
    TV q;                     // s8 input, ldc % 4 == 0
    TV out;                   // f32
    const uint32_t *w;
    const float *bias;
    float alpha1;
    int n, size, stride, pad, act;
    int CW;                   // words per pixel = q.ldc/4
    long M;
    int32_t *acc_out;         // optional raw accumulators, NCHW (tests)
};

__device__ __forceinline__ float int8_epilogue(int acc, float alpha1, float bias, int act) {
    int q = acc / 32;
    if (q > 32767) q = 32767;
    if (q < -32767) q = -32767;
    float y = __fmul_rn((float)q, alpha1);
    y = __fadd_rn(y, bias);
    if (act == ACT_LEAKY) y = (y > 0.f) ? y : __fdiv_rn(y, 10.f);
    return y;
}

static __global__ void __launch_bounds__(256) k_conv_int8_simt(Int8P p) {
    constexpr int BM = 64, BN = 64, BKW = 8;
    __shared__ uint32_t As[BKW][BM + 1];
    __shared__ uint32_t Bs[BKW][BN + 1];
    const int tid = threadIdx.x;
    const long m0 = (long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int OH = p.out.H, OW = p.out.W, CW = p.CW;
    const int KW = p.size * p.size * CW;
    const int lr = tid >> 2, lw = (tid & 3) * 2;
    const long lm = m0 + lr;
    const bool lvalid = lm < p.M;
    int ln = 0, liy0 = 0, lix0 = 0;
    if (lvalid) {
        const int ox = (int)(lm % OW);
        const int oy = (int)((lm / OW) % OH);
        ln = (int)(lm / ((long)OW * OH));
        liy0 = oy * p.stride - p.pad;
        lix0 = ox * p.stride - p.pad;
    }
    const int tx = tid & 15, ty = tid >> 4;
    int acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;

    for (int k0 = 0; k0 < KW; k0 += BKW) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = k0 + lw + q;
            uint32_t a = 0, b = 0;
            if (k < KW) {
                const int tap = k / CW, wd = k - tap * CW;
                if (lvalid) {
                    const int ky = tap / p.size, kx = tap - ky * p.size;
                    const int iy = liy0 + ky, ix = lix0 + kx;
                    if (iy >= 0 && iy < p.q.H && ix >= 0 && ix < p.q.W)
                        a = reinterpret_cast<const uint32_t *>(tv_px<int8_t>(p.q, ln, iy, ix))[wd];
                }
                b = p.w[(size_t)(n0 + lr) * KW + k];
            }
            As[lw + q][lr] = a;
            Bs[lw + q][lr] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BKW; ++kk) {
            uint32_t a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __dp4a((int)a[i], (int)b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        const int ox = (int)(m % OW);
        const int oy = (int)((m / OW) % OH);
        const int n = (int)(m / ((long)OW * OH));
        float *o = tv_px<float>(p.out, n, oy, ox);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = n0 + tx * 4 + j;
            if (f >= p.n) continue;
            if (p.acc_out) p.acc_out[(((size_t)n * p.n + f) * OH + oy) * OW + ox] = acc[i][j];
            o[f] = int8_epilogue(acc[i][j], p.alpha1, p.bias[f], p.act);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// small layers (one thread per output element, channels innermost -> coalesced)
// ------------------------------------------------------------------------------------------------------

// forward_maxpool_layer_avx scalar build (reference additionally.c:1448-1482): window origin
// (o*stride - pad/2), out-of-image taps ignored, init -FLT_MAX.
template <typename T>
__global__ void k_maxpool(TV in, TV out, int size, int stride, int pad) {
    const long total = (long)out.N * out.H * out.W * out.C;
    const int off = -pad / 2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % out.C);
        const long pxl = i / out.C;
        const int x = (int)(pxl % out.W);
        const int y = (int)((pxl / out.W) % out.H);
        const int n = (int)(pxl / ((long)out.W * out.H));
        float m = -3.402823466e+38f;
        for (int a = 0; a < size; ++a) {
            const int iy = off + y * stride + a;
            if (iy < 0 || iy >= in.H) continue;
            for (int b = 0; b < size; ++b) {
                const int ix = off + x * stride + b;
                if (ix < 0 || ix >= in.W) continue;
                const float v = to_f32(tv_px<T>(in, n, iy, ix)[c]);
                m = (v > m) ? v : m;
            }
        }
        tv_px<T>(out, n, y, x)[c] = from_f32<T>(m);
    }
}

// same, 16 bytes (4 f32 / 8 bf16 channels) per thread
template <typename T>
__global__ void k_maxpool_vec(TV in, TV out, int size, int stride, int pad) {
    constexpr int V = 16 / sizeof(T);
    const int chunks = out.C / V;
    const long total = (long)out.N * out.H * out.W * chunks;
    const int off = -pad / 2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long pxl = i / chunks;
        const int x = (int)(pxl % out.W);
        const int y = (int)((pxl / out.W) % out.H);
        const int n = (int)(pxl / ((long)out.W * out.H));
        float m[V];
#pragma unroll
        for (int k = 0; k < V; ++k) m[k] = -3.402823466e+38f;
        for (int a = 0; a < size; ++a) {
            const int iy = off + y * stride + a;
            if (iy < 0 || iy >= in.H) continue;
            for (int b = 0; b < size; ++b) {
                const int ix = off + x * stride + b;
                if (ix < 0 || ix >= in.W) continue;
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(tv_px<T>(in, n, iy, ix)) + ch);
                const T *v = reinterpret_cast<const T *>(&raw);
#pragma unroll
                for (int k = 0; k < V; ++k) { const float f = to_f32(v[k]); m[k] = (f > m[k]) ? f : m[k]; }
            }
        }
        uint4 o;
        T *ov = reinterpret_cast<T *>(&o);
#pragma unroll
        for (int k = 0; k < V; ++k) ov[k] = from_f32<T>(m[k]);
        reinterpret_cast<uint4 *>(tv_px<T>(out, n, y, x))[ch] = o;
    }
}

// max-pool fused with the consumer's input transform (exact: same values, same operation order as max-pool followed by
// k_quantize / k_binarize_s8 / k_binarize_vec -- the f32 pooled tensor just never goes to HBM).
//   MODE 0: s8 = quant_i8(max, mult)          -> q (s8, ldc % 4 == 0, zero channel padding)   [INT8 convolutions]
//   MODE 1: s8 = max > 0 ? +1 : -1            -> q (s8, ldc == C, C % 4 == 0)                 [XNOR on kind::i8]
//   MODE 2: bit = max > 0, 32 channels / word -> bits (ldc words)                              [XNOR popcount kernels]
template <int MODE>
__global__ void k_maxpool_fused(TV in /* f32 */, TV q, int size, int stride, int pad, float mult) {
    const int groups = (MODE == 2) ? q.ldc : (q.ldc >> 2);      // output words per pixel
    const int OH = q.H, OW = q.W;
    const long total = (long)q.N * OH * OW * groups;
    const int off = -pad / 2;
    constexpr int V = (MODE == 2) ? 32 : 4;                      // channels per output word
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long pxl = i / groups;
        const int x = (int)(pxl % OW);
        const int y = (int)((pxl / OW) % OH);
        const int n = (int)(pxl / ((long)OW * OH));
        float m[V];
#pragma unroll
        for (int k = 0; k < V; ++k) m[k] = -3.402823466e+38f;
        for (int a = 0; a < size; ++a) {
            const int iy = off + y * stride + a;
            if (iy < 0 || iy >= in.H) continue;
            for (int b = 0; b < size; ++b) {
                const int ix = off + x * stride + b;
                if (ix < 0 || ix >= in.W) continue;
                const float *src = tv_px<float>(in, n, iy, ix) + g * V;
#pragma unroll
                for (int k = 0; k < V; k += 4) {
                    if (g * V + k + 3 < in.C) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + k));
                        m[k] = v.x > m[k] ? v.x : m[k]; m[k + 1] = v.y > m[k + 1] ? v.y : m[k + 1];
                        m[k + 2] = v.z > m[k + 2] ? v.z : m[k + 2]; m[k + 3] = v.w > m[k + 3] ? v.w : m[k + 3];
                    } else {
                        for (int j = 0; j < 4; ++j)
                            if (g * V + k + j < in.C) { const float f = __ldg(src + k + j); m[k + j] = f > m[k + j] ? f : m[k + j]; }
                    }
                }
            }
        }
        uint32_t word = 0;
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = g * 4 + j;
                const int s = (c < in.C) ? quant_i8(m[j], mult) : 0;
                word |= (uint32_t)(s & 0xff) << (8 * j);
            }
            reinterpret_cast<uint32_t *>(tv_px<int8_t>(q, n, y, x))[g] = word;
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) word |= (uint32_t)((m[j] > 0.f ? 1 : -1) & 0xff) << (8 * j);
            reinterpret_cast<uint32_t *>(tv_px<int8_t>(q, n, y, x))[g] = word;
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (g * 32 + j < in.C && m[j] > 0.f) word |= 1u << j;
            tv_px<uint32_t>(q, n, y, x)[g] = word;
        }
    }
}

// upsample_cpu forward (reference yolov2_forward_network.c:380-394): out = scale * in[y/stride][x/stride]
// Packed f32x2 arithmetic of sm_100 (two IEEE round-to-nearest operations per instruction, each lane rounded on its own --
// exactly __fmul_rn / __fadd_rn twice): halves the instruction count of the exact-order float convolutions.
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
// The product is written as fma(a, b, nz) with nz = (-0, -0) handed in at RUN TIME: RN(a*b + (-0)) == RN(a*b) bit for bit (sign of
// zero included), and ptxas cannot prove the addend away.  A plain mul.rn.f32x2 + add.rn.f32x2 pair -- and even fma(a, b, literal
// -0) + add -- it contracts into ONE FFMA2 (single rounding), --fmad=false or not, which silently breaks bit-exactness with the
// reference's separately rounded gemm_nn (found in the SASS, caught by test_fused_stem_pool_is_bit_identical...).
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b, unsigned long long nz) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ------------------------------------------------------------------------------------------------------
// Exact nets (INT8 / XNOR tiny models): stem convolution + 2x2/2 max-pool + the next integer layer's input conversion in ONE
// kernel.  Layers 0-2 of yolov3-tiny / tiny-yolo-obj_xnor are conv 3->16 (f32), maxpool 2/2, integer conv: unfused, the f32 stem
// output (709 MB at 416x416 batch 64) is written once and read once just to be reduced 4:1 and narrowed to one byte (or bit)
// per value.  One thread per POOLED pixel: the 4x4x3 input window (48 loads), four stem outputs x 16 filters in the
// reference's exact order (c, ky, kx; separately rounded products and sums: k_conv_stem<EXACT>), bias + activation, the
// reference's max (forward_maxpool_layer_avx scalar semantics: -FLT_MAX start, strict >), then quant_i8 / sign exactly as
// k_maxpool_fused.  Bit-identical to the three separate kernels.
// MODE 0: s8 quantised (q.ldc bytes per pixel, channels >= 16 stay zero); 1: +-1 bytes; 2: sign bits (one word per pixel).
// ------------------------------------------------------------------------------------------------------
// ACT is a template parameter: act_exact() with a run-time activation drags the double-precision logistic (exp) into each of the 64
// call sites -- 19 k SASS instructions, 300 KB of code that no instruction cache holds.
template <int MODE, int ACT>
__global__ void __launch_bounds__(128) k_stem_pool(const float *__restrict__ in, TV q, const __grid_constant__ StemW<16> sw, int /*act*/,
                                                   int H, int W, float mult, unsigned long long negzero2 /* 0x8000000080000000 */) {
    constexpr int NF = 16;
    const int OH = q.H, OW = q.W;
    const long total = (long)q.N * OH * OW;
    const long pidx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (pidx >= total) return;
    const int px = (int)(pidx % OW), py = (int)((pidx / OW) % OH), n = (int)(pidx / ((long)OW * OH));
    const float *img = in + (size_t)n * 3 * H * W;
    unsigned long long acc2[4][NF / 2];          // (filter 2j, filter 2j + 1) pairs: mul.rn.f32x2 / add.rn.f32x2
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < NF / 2; ++j) acc2[k][j] = 0ull;
    const int y0 = 2 * py - 1, x0 = 2 * px - 1;            // top-left of the 4x4 input window
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        unsigned long long win2[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int iy = y0 + a, ix = x0 + b;
                const float v = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(img + ((size_t)c * H + iy) * W + ix) : 0.f;
                win2[a][b] = f2_pack(v, v);
            }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int j = 0; j < NF / 2; ++j) {
                    const float *wp = &sw.w[((ky * 3 + kx) * 3 + c) * NF + 2 * j];
                    const unsigned long long w2 = f2_pack(wp[0], wp[1]);
#pragma unroll
                    for (int k = 0; k < 4; ++k)     // output pixel (dy, dx) = (k >> 1, k & 1); per accumulator the order is (c, ky, kx)
                        acc2[k][j] = f2_add(acc2[k][j], f2_mul(w2, win2[(k >> 1) + ky][(k & 1) + kx], negzero2));
                }
    }
    float acc[4][NF];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < NF / 2; ++j) f2_unpack(acc2[k][j], acc[k][2 * j], acc[k][2 * j + 1]);
    float m[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float mx = -3.402823466e+38f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {                       // window order of the reference: rows, then columns
            const bool inside = (2 * py + (k >> 1)) < H && (2 * px + (k & 1)) < W;
            float v = __fadd_rn(acc[k][f], sw.b[f]);
            if (ACT == ACT_LEAKY) v = (v > 0.f) ? v : (float)(0.1 * (double)v);     // activate(), additionally.h:91 (scalar build)
            if (inside) mx = v > mx ? v : mx;
        }
        m[f] = mx;
    }
    if (MODE == 0) {
        uint32_t wq[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) word |= (uint32_t)(quant_i8(m[g * 4 + j], mult) & 0xff) << (8 * j);
            wq[g] = word;
        }
        *reinterpret_cast<uint4 *>(tv_px<int8_t>(q, n, py, px)) = make_uint4(wq[0], wq[1], wq[2], wq[3]);
    } else if (MODE == 1) {     // +-1 bytes: the next XNOR layer runs as +-1 on kind::i8
        uint32_t wq[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) word |= (m[g * 4 + j] > 0.f ? 0x01u : 0xFFu) << (8 * j);
            wq[g] = word;
        }
        *reinterpret_cast<uint4 *>(tv_px<int8_t>(q, n, py, px)) = make_uint4(wq[0], wq[1], wq[2], wq[3]);
    } else {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < NF; ++j)
            if (m[j] > 0.f) word |= 1u << j;
        tv_px<uint32_t>(q, n, py, px)[0] = word;
    }
}

template <typename T>
__global__ void k_upsample(TV in, TV out, int stride, float scale) {
    const long total = (long)out.N * out.H * out.W * out.C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % out.C);
        const long pxl = i / out.C;
        const int x = (int)(pxl % out.W);
        const int y = (int)((pxl / out.W) % out.H);
        const int n = (int)(pxl / ((long)out.W * out.H));
        const float v = to_f32(tv_px<T>(in, n, y / stride, x / stride)[c]);
        tv_px<T>(out, n, y, x)[c] = from_f32<T>(__fmul_rn(scale, v));
    }
}

// same, 16 bytes per thread (scale == 1: a pure copy of bits, exact for any dtype)
static __global__ void k_upsample_vec16(TV in, TV out, int stride, int esize) {
    const int chunks = (out.C * esize) >> 4;
    const long total = (long)out.N * out.H * out.W * chunks;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % chunks);
        const long pxl = i / chunks;
        const int x = (int)(pxl % out.W);
        const int y = (int)((pxl / out.W) % out.H);
        const int n = (int)(pxl / ((long)out.W * out.H));
        const char *src = in.base + (((size_t)(n * in.Hp + y / stride + in.P) * in.Wp + (x / stride + in.P)) * (size_t)in.ldc) * esize;
        char *dst = out.base + (((size_t)(n * out.Hp + y + out.P) * out.Wp + (x + out.P)) * (size_t)out.ldc) * esize;
        reinterpret_cast<uint4 *>(dst)[ch] = __ldg(reinterpret_cast<const uint4 *>(src) + ch);
    }
}

// forward_shortcut_layer_cpu (reference yolov2_forward_network.c:443-449, shortcut_cpu :410-432):
// out = act(in + from) with the general stride/sample subsampling of shortcut_cpu.
template <typename T>
__global__ void k_shortcut(TV in, TV from, TV out, int stride, int sample, int minw, int minh, int minc, int act) {
    const long total = (long)out.N * out.H * out.W * out.C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % out.C);
        const long pxl = i / out.C;
        const int x = (int)(pxl % out.W);
        const int y = (int)((pxl / out.W) % out.H);
        const int n = (int)(pxl / ((long)out.W * out.H));
        float v = to_f32(tv_px<T>(in, n, y, x)[c]);
        if (c < minc && (y % sample) == 0 && (x % sample) == 0) {
            const int j = y / sample, ii = x / sample;
            if (j < minh && ii < minw) v = __fadd_rn(v, to_f32(tv_px<T>(from, n, j * stride, ii * stride)[c]));
        }
        tv_px<T>(out, n, y, x)[c] = from_f32<T>(act_exact(v, act));
    }
}

// route: copy one source into its channel slice of the concat buffer (reference yolov2_forward_network.c:318)
template <typename T>
__global__ void k_copy_channels(TV in, TV out /* view of the slice: C == in.C */) {
    const long total = (long)in.N * in.H * in.W * in.C;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % in.C);
        const long pxl = i / in.C;
        const int x = (int)(pxl % in.W);
        const int y = (int)((pxl / in.W) % in.H);
        const int n = (int)(pxl / ((long)in.W * in.H));
        tv_px<T>(out, n, y, x)[c] = tv_px<T>(in, n, y, x)[c];
    }
}

// forward_reorg_layer_cpu (reference yolov2_forward_network.c:337-373), darknet's space-to-depth flavour:
// out[k][j][i] = x_flat[ w2 + (out_w*stride) * (h2 + (out_h*stride) * c2) ] with the input tensor reinterpreted
// as [in_c][out_h*stride][out_w*stride] where in_c = out_c/stride^2.
template <typename T>
__global__ void k_reorg(TV in, TV out, int stride) {
    const long total = (long)out.N * out.H * out.W * out.C;
    const int in_c = out.C / (stride * stride);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % out.C);
        const long pxl = i / out.C;
        const int x = (int)(pxl % out.W);
        const int y = (int)((pxl / out.W) % out.H);
        const int n = (int)(pxl / ((long)out.W * out.H));
        const int c2 = k % in_c, offset = k / in_c;
        const int w2 = x * stride + offset % stride, h2 = y * stride + offset / stride;
        // flat index inside one image of the source viewed as [in_c][out.H*stride][out.W*stride]
        const long flat = w2 + (long)out.W * stride * (h2 + (long)out.H * stride * c2);
        // map the flat NCHW index back onto the real source dims [in.C][in.H][in.W]
        const int sx = (int)(flat % in.W);
        const int sy = (int)((flat / in.W) % in.H);
        const int sc = (int)(flat / ((long)in.W * in.H));
        tv_px<T>(out, n, y, x)[k] = tv_px<T>(in, n, sy, sx)[sc];
    }
}

// forward_yolo_layer_cpu (reference yolov2_forward_network.c:453-472): copy + logistic on entries 0,1 and
// 4..4+classes of each anchor block.  Reads the head conv's NHWC activation, writes the NCHW f32 tensor the
// reference decoder expects (additionally.c:4200 entry_index).
template <typename TIn>
__global__ void __launch_bounds__(256) k_yolo(TV in, float *__restrict__ out, int classes, int fast) {
    // 32 pixels x 32 channels per tile: channel-contiguous reads (NHWC), pixel-contiguous writes (NCHW)
    __shared__ float tile[32][33];
    const int HW = in.H * in.W;
    const int ptiles = (HW + 31) / 32, ctiles = (in.C + 31) / 32;
    const long ntiles = (long)in.N * ptiles * ctiles;
    const int per = 4 + classes + 1;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;   // 8 warps
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int ct = (int)(t % ctiles);
        const int pt = (int)((t / ctiles) % ptiles);
        const int n = (int)(t / ((long)ctiles * ptiles));
        const int c = ct * 32 + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int pl = wy * 4 + k;
            const int hw = pt * 32 + pl;
            float v = 0.f;
            if (hw < HW && c < in.C) {
                v = to_f32(tv_px<TIn>(in, n, hw / in.W, hw % in.W)[c]);
                const int e = c % per;
                // exact nets: the reference's double-precision logistic; bf16 tensor-core nets: f32 (error ~1e-7,
                // far below the path's 1e-3 bar)
                if (e != 2 && e != 3) v = fast ? 1.f / (1.f + __expf(-v)) : (float)(1.0 / (1.0 + exp(-(double)v)));
            }
            tile[pl][lane] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int cl = wy * 4 + k;
            const int cc = ct * 32 + cl;
            const int hw = pt * 32 + lane;
            if (hw < HW && cc < in.C) out[((size_t)n * in.C + cc) * HW + hw] = tile[lane][cl];
        }
        __syncthreads();
    }
}

// forward_region_layer_cpu (reference yolov2_forward_network.c:511-575): per image HWC flatten (== our NHWC
// order), float logistic on entry 4, softmax over classes (softmax_cpu :476) when softmax=1.
// One thread per (image, cell, anchor).
template <typename TIn>
__global__ void k_region(TV in, float *__restrict__ out, int nanchors, int classes, int coords, int softmax) {
    const int size = coords + classes + 1;
    const long total = (long)in.N * in.H * in.W * nanchors;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int a = (int)(i % nanchors);
        const long pxl = i / nanchors;
        const int x = (int)(pxl % in.W);
        const int y = (int)((pxl / in.W) % in.H);
        const int n = (int)(pxl / ((long)in.W * in.H));
        const TIn *src = tv_px<TIn>(in, n, y, x) + a * size;
        float *o = out + (((size_t)n * in.H + y) * in.W + x) * (size_t)(nanchors * size) + (size_t)a * size;
        for (int k = 0; k < coords; ++k) o[k] = to_f32(src[k]);
        o[coords] = 1.0f / (1.0f + expf(-to_f32(src[coords])));
        if (softmax) {
            float largest = -3.402823466e+38f;
            for (int k = 0; k < classes; ++k) { const float v = to_f32(src[coords + 1 + k]); if (v > largest) largest = v; }
            float sum = 0.f;
            for (int k = 0; k < classes; ++k) {
                const float e = expf(to_f32(src[coords + 1 + k]) - largest);
                sum += e;
                o[coords + 1 + k] = e;
            }
            for (int k = 0; k < classes; ++k) o[coords + 1 + k] = __fdiv_rn(o[coords + 1 + k], sum);
        } else {
            for (int k = 0; k < classes; ++k) o[coords + 1 + k] = to_f32(src[coords + 1 + k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// INT8 input calibration: histogram of |x| over the logical elements of image `img` of an
// activation tensor, binned exactly like the reference (yolov2_forward_network_quantized.c:1308-1316:
// lround(fabs(x) / bin_width) in double, saturated into the last bin).  Per-block shared-memory histogram,
// integer atomics -> exact counts.  max_bin <= 4096.
// ------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_abs_hist(TV in, int img, float bin_width, int max_bin, unsigned *__restrict__ hist) {
    __shared__ unsigned sh[4096];
    for (int i = threadIdx.x; i < max_bin; i += 256) sh[i] = 0u;
    __syncthreads();
    const long per = (long)in.C * in.H * in.W;
    const int last_bin = max_bin - 1;
    for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < per; e += (long)gridDim.x * 256) {
        const int c = (int)(e % in.C);
        const long px = e / in.C;
        const int x = (int)(px % in.W), y = (int)(px / in.W);
        const float v = to_f32(tv_px<T>(in, img, y, x)[c]);
        const long b = lround(fabs((double)v) / (double)bin_width);
        atomicAdd(&sh[b >= last_bin ? last_bin : (int)b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < max_bin; i += 256)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// the network input (NCHW f32, as the caller passes it)
static __global__ void __launch_bounds__(256) k_abs_hist_flat(const float *__restrict__ src, long n, float bin_width, int max_bin,
                                                              unsigned *__restrict__ hist) {
    __shared__ unsigned sh[4096];
    for (int i = threadIdx.x; i < max_bin; i += 256) sh[i] = 0u;
    __syncthreads();
    const int last_bin = max_bin - 1;
    for (long e = (long)blockIdx.x * 256 + threadIdx.x; e < n; e += (long)gridDim.x * 256) {
        const long b = lround(fabs((double)src[e]) / (double)bin_width);
        atomicAdd(&sh[b >= last_bin ? last_bin : (int)b], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < max_bin; i += 256)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

}  // namespace yb
