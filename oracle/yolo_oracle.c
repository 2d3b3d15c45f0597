/*
 * yolo_oracle.c -- TEST INFRASTRUCTURE ONLY.  A plain-C CPU restatement of the reference's forward hot path
 * (AlexeyAB/yolo2_light), used as the parity checker for the CUDA engine.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this; the product library never does.
 *
 * Pinned: tests/test_oracle_vs_reference.py runs every function here against the UNMODIFIED reference built by
 * oracle/Makefile (scalar, -O2) on the same seeded inputs -- integer paths (XNOR, INT8) and their float
 * epilogues bit-for-bit, the FP32 conv bit-for-bit as well (same k-ascending float accumulation) -- and against
 * the golden vectors in tests/golden/ that the reference itself produced (tests/golden/make_golden.py).
 *
 * Tensors are NCHW float, batch-major, exactly as the reference holds them.  Build: -O2 -fno-fast-math
 * -ffp-contract=off (see oracle/Makefile) so float expressions evaluate as written.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { YO_LOGISTIC = 0, YO_RELU = 1, YO_LINEAR = 3, YO_LEAKY = 7 };

/* activate(), additionally.h:126-157, scalar build: leaky multiplies by the DOUBLE literal .1
 * (additionally.h:91), logistic uses double exp (additionally.h:85). */
static float yo_activate(float x, int a)
{
    switch (a) {
    case YO_LINEAR: return x;
    case YO_LEAKY: return (x > 0) ? x : (float)(.1 * (double)x);
    case YO_LOGISTIC: return (float)(1. / (1. + exp(-(double)x)));
    case YO_RELU: return x * (x > 0);
    default: return x;
    }
}

void yo_activate_array(float *x, long n, int a)
{
    if (a == YO_LINEAR) return;   /* activate_array_cpu_custom, additionally.c:1436 */
    for (long i = 0; i < n; ++i) x[i] = yo_activate(x[i], a);
}

/* FP32 convolution: forward_convolutional_layer_cpu FP32 branch (yolov2_forward_network.c:204-261) =
 * im2col_cpu (additionally.c:39) + gemm_nn (additionally.c:1272) + bias (:243) + activation (:261).
 * gemm_nn accumulates C[j] += A[k]*B[k][j] with k = (c,ky,kx) ascending, zero-padded taps included, starting
 * from the zero fill at :38 -- reproduced literally so the result is bit-identical to the scalar reference. */
void yo_conv_fp32(const float *in, int batch, int c, int h, int w,
                  const float *weights, const float *biases, int n, int size, int stride, int pad,
                  int activation, float *out)
{
    const int out_h = (h + 2 * pad - size) / stride + 1;
    const int out_w = (w + 2 * pad - size) / stride + 1;
    for (int b = 0; b < batch; ++b)
        for (int f = 0; f < n; ++f)
            for (int oy = 0; oy < out_h; ++oy)
                for (int ox = 0; ox < out_w; ++ox) {
                    float acc = 0;
                    const float *wf = weights + (size_t)f * c * size * size;
                    for (int ch = 0; ch < c; ++ch)
                        for (int ky = 0; ky < size; ++ky)
                            for (int kx = 0; kx < size; ++kx) {
                                const int iy = oy * stride + ky - pad, ix = ox * stride + kx - pad;
                                const float v = (iy < 0 || ix < 0 || iy >= h || ix >= w)
                                                    ? 0.f : in[(((size_t)b * c + ch) * h + iy) * w + ix];
                                acc += wf[(ch * size + ky) * size + kx] * v;
                            }
                    acc += biases[f];
                    out[(((size_t)b * n + f) * out_h + oy) * out_w + ox] = yo_activate(acc, activation);
                }
}

/* BIT1-XNOR convolution: XNOR branch of forward_convolutional_layer_cpu (yolov2_forward_network.c:116-203),
 * stride 1 / pad 1 only.  SURVEY Appendix A: input bit = (x > 0) for in-image taps, 0 for out-of-image taps
 * (im2col zero fill, additionally.c:1369-1423, F9); weight bit = (w > 0) (binarize_weights :113 + float_to_bit
 * :1536); count = #taps with equal bits (gemm_nn_custom_bin_mean_transposed :1504-1534, alignment bits removed);
 * out = (2*count - K) * mean_arr[f]  (:1531), then + bias (:243) and activation (:261).
 * counts (optional, may be NULL) receives the raw integer popcounts for bit-exact checks. */
void yo_conv_xnor(const float *in, int batch, int c, int h, int w,
                  const float *weights, const float *biases, const float *mean_arr, int n, int size,
                  int activation, float *out, int32_t *counts)
{
    const int pad = 1, K = size * size * c;
    for (int b = 0; b < batch; ++b)
        for (int f = 0; f < n; ++f)
            for (int oy = 0; oy < h; ++oy)
                for (int ox = 0; ox < w; ++ox) {
                    int count = 0;
                    const float *wf = weights + (size_t)f * K;
                    for (int ch = 0; ch < c; ++ch)
                        for (int ky = 0; ky < size; ++ky)
                            for (int kx = 0; kx < size; ++kx) {
                                const int iy = oy + ky - pad, ix = ox + kx - pad;
                                int ib = 0;
                                if (!(iy < 0 || ix < 0 || iy >= h || ix >= w))
                                    ib = in[(((size_t)b * c + ch) * h + iy) * w + ix] > 0;
                                const int wb = wf[(ch * size + ky) * size + kx] > 0;
                                count += (ib == wb);
                            }
                    const size_t o = (((size_t)b * n + f) * h + oy) * w + ox;
                    if (counts) counts[o] = count;
                    float v = (float)(2 * count - K) * mean_arr[f];
                    v += biases[f];
                    out[o] = yo_activate(v, activation);
                }
}

static int yo_max_abs(int src, int max_val)   /* yolov2_forward_network_quantized.c:23 */
{
    if (abs(src) > abs(max_val)) src = (src > 0) ? max_val : -max_val;
    return src;
}

/* float -> int16_t as the reference's `int16_t src = x * mult;` compiles on x86-64 (cvttss2si to 32 bit,
 * low 16 bits kept; out-of-range/NaN -> 0x80000000 -> 0). */
static int16_t yo_to_i16(float v)
{
    int32_t i;
    if (!(v > -2147483648.0f && v < 2147483648.0f)) i = (int32_t)0x80000000;
    else i = (int32_t)v;
    return (int16_t)(i & 0xffff);
}

/* input quantisation of forward_convolutional_layer_q (yolov2_forward_network_quantized.c:556-560) */
void yo_quantize_input(const float *in, long n, float input_mult, int8_t *out)
{
    for (long z = 0; z < n; ++z) out[z] = (int8_t)yo_max_abs(yo_to_i16(in[z] * input_mult), 127);
}

/* INT8 convolution: forward_convolutional_layer_q (yolov2_forward_network_quantized.c:527-631) with
 * im2col_cpu_int8 (:186) and gemm_nn_int8_int16 (:469): acc32 = sum wq*xq over in-image taps;
 * q16 = clamp(+-32767, acc32 / 32) (C division truncates toward zero); y = q16 * ALPHA1,
 * ALPHA1 = 32 / (input_mult * weights_mult) (:598); y += bias (:612); leaky as y/10 (:623).
 * The reference ignores l.batch here; this restatement loops images.
 * acc_out (optional) receives the raw s32 accumulators. */
void yo_conv_int8(const float *in, int batch, int c, int h, int w,
                  const int8_t *weights_int8, const float *biases, float input_mult, float weights_mult,
                  int n, int size, int stride, int pad, int activation, float *out, int32_t *acc_out)
{
    const int out_h = (h + 2 * pad - size) / stride + 1;
    const int out_w = (w + 2 * pad - size) / stride + 1;
    const size_t in_sz = (size_t)c * h * w;
    int8_t *xq = (int8_t *)malloc(in_sz);
    const float ALPHA1 = 32 / (input_mult * weights_mult);
    for (int b = 0; b < batch; ++b) {
        yo_quantize_input(in + (size_t)b * in_sz, (long)in_sz, input_mult, xq);
        for (int f = 0; f < n; ++f)
            for (int oy = 0; oy < out_h; ++oy)
                for (int ox = 0; ox < out_w; ++ox) {
                    int32_t acc = 0;
                    const int8_t *wf = weights_int8 + (size_t)f * c * size * size;
                    for (int ch = 0; ch < c; ++ch)
                        for (int ky = 0; ky < size; ++ky)
                            for (int kx = 0; kx < size; ++kx) {
                                const int iy = oy * stride + ky - pad, ix = ox * stride + kx - pad;
                                if (iy < 0 || ix < 0 || iy >= h || ix >= w) continue;
                                acc += (int32_t)wf[(ch * size + ky) * size + kx] * (int32_t)xq[((size_t)ch * h + iy) * w + ix];
                            }
                    const size_t o = (((size_t)b * n + f) * out_h + oy) * out_w + ox;
                    if (acc_out) acc_out[o] = acc;
                    const int16_t q16 = (int16_t)yo_max_abs(acc / 32, 256 * 128 - 1);
                    float y = q16 * ALPHA1;
                    y += biases[f];
                    if (activation == YO_LEAKY) y = (y > 0) ? y : y / 10;
                    out[o] = y;
                }
    }
    free(xq);
}

/* forward_maxpool_layer_avx, scalar build (additionally.c:1448-1482): window origin (i*stride - pad/2),
 * out-of-image taps skipped, init -FLT_MAX. */
void yo_maxpool(const float *in, int batch, int c, int h, int w, int size, int stride, int pad, float *out)
{
    const int out_w = (w + pad - size) / stride + 1, out_h = (h + pad - size) / stride + 1;
    const int off = -pad / 2;
    for (int b = 0; b < batch; ++b)
        for (int k = 0; k < c; ++k)
            for (int i = 0; i < out_h; ++i)
                for (int j = 0; j < out_w; ++j) {
                    float m = -FLT_MAX;
                    for (int n = 0; n < size; ++n)
                        for (int mm = 0; mm < size; ++mm) {
                            const int ch = off + i * stride + n, cw = off + j * stride + mm;
                            if (ch >= 0 && ch < h && cw >= 0 && cw < w) {
                                const float v = in[(((size_t)b * c + k) * h + ch) * w + cw];
                                m = (v > m) ? v : m;
                            }
                        }
                    out[(((size_t)b * c + k) * out_h + i) * out_w + j] = m;
                }
}

/* upsample_cpu forward, yolov2_forward_network.c:380-394 (after fill 0: out = scale * in) */
void yo_upsample(const float *in, int batch, int c, int h, int w, int stride, float scale, float *out)
{
    for (int b = 0; b < batch; ++b)
        for (int k = 0; k < c; ++k)
            for (int j = 0; j < h * stride; ++j)
                for (int i = 0; i < w * stride; ++i)
                    out[(((size_t)b * c + k) * h * stride + j) * w * stride + i] =
                        scale * in[(((size_t)b * c + k) * h + j / stride) * w + i / stride];
}

/* forward_shortcut_layer_cpu, yolov2_forward_network.c:443-449: out = in; out += from (shortcut_cpu :410,
 * general stride/sample form); activation.  (w1,h1,c1) = `from` tensor, (w2,h2,c2) = in/out tensor. */
void yo_shortcut(const float *in, const float *from, int batch, int w1, int h1, int c1, int w2, int h2, int c2,
                 int activation, float *out)
{
    int stride = w1 / w2, sample = w2 / w1;
    if (stride < 1) stride = 1;
    if (sample < 1) sample = 1;
    const int minw = w1 < w2 ? w1 : w2, minh = h1 < h2 ? h1 : h2, minc = c1 < c2 ? c1 : c2;
    memcpy(out, in, sizeof(float) * (size_t)batch * w2 * h2 * c2);
    for (int b = 0; b < batch; ++b)
        for (int k = 0; k < minc; ++k)
            for (int j = 0; j < minh; ++j)
                for (int i = 0; i < minw; ++i)
                    out[i * sample + w2 * (j * sample + h2 * (k + (size_t)c2 * b))] +=
                        from[i * stride + w1 * (j * stride + h1 * (k + (size_t)c1 * b))];
    yo_activate_array(out, (long)batch * w2 * h2 * c2, activation);
}

/* forward_reorg_layer_cpu, yolov2_forward_network.c:337-373 (out_* are the layer's output dims) */
void yo_reorg(const float *x, int batch, int out_c, int out_h, int out_w, int stride, float *out)
{
    const int in_c = out_c / (stride * stride);
    for (int b = 0; b < batch; ++b)
        for (int k = 0; k < out_c; ++k)
            for (int j = 0; j < out_h; ++j)
                for (int i = 0; i < out_w; ++i) {
                    const size_t in_index = i + out_w * (j + out_h * (k + (size_t)out_c * b));
                    const int c2 = k % in_c, offset = k / in_c;
                    const int w2 = i * stride + offset % stride, h2 = j * stride + offset / stride;
                    const size_t out_index = w2 + (size_t)out_w * stride * (h2 + (size_t)out_h * stride * (c2 + (size_t)in_c * b));
                    out[in_index] = x[out_index];
                }
}

/* forward_yolo_layer_cpu, yolov2_forward_network.c:453-472 with entry_index (additionally.c:4200):
 * logistic on entries 0,1 and 4..4+classes of each anchor block; w,h entries raw. */
void yo_yolo(const float *in, int batch, int n, int classes, int h, int w, float *out)
{
    const int hw = h * w, per = 4 + classes + 1;
    const size_t outputs = (size_t)n * per * hw;
    memcpy(out, in, sizeof(float) * outputs * batch);
    for (int b = 0; b < batch; ++b)
        for (int a = 0; a < n; ++a) {
            float *p = out + b * outputs + (size_t)a * per * hw;
            yo_activate_array(p, 2L * hw, YO_LOGISTIC);
            yo_activate_array(p + 4 * hw, (long)(1 + classes) * hw, YO_LOGISTIC);
        }
}

/* forward_region_layer_cpu, yolov2_forward_network.c:511-575: CHW -> HWC flatten per image, float logistic
 * (expf) on entry 4 of each (cell, anchor) block, softmax_cpu (:476) over the classes when softmax=1. */
void yo_region(const float *in, int batch, int n, int classes, int coords, int h, int w, int softmax, float *out)
{
    const int size = coords + classes + 1, hw = h * w, layers = size * n;
    const size_t outputs = (size_t)hw * layers;
    for (int b = 0; b < batch; ++b)
        for (int c = 0; c < layers; ++c)
            for (int i = 0; i < hw; ++i)
                out[b * outputs + (size_t)i * layers + c] = in[b * outputs + (size_t)c * hw + i];
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < hw * n; ++i) {
            float *p = out + b * outputs + (size_t)size * i;
            p[4] = 1.0F / (1.0F + expf(-p[4]));
            if (softmax) {
                float *cls = p + 5, sum = 0, largest = -FLT_MAX;
                for (int k = 0; k < classes; ++k) if (cls[k] > largest) largest = cls[k];
                for (int k = 0; k < classes; ++k) {
                    const float e = expf(cls[k] / 1 - largest / 1);
                    sum += e;
                    cls[k] = e;
                }
                for (int k = 0; k < classes; ++k) cls[k] /= sum;
            }
        }
}

/* forward_route_layer_cpu, yolov2_forward_network.c:318-334: channel concat, one source at a time */
void yo_route_copy(const float *src, int batch, long src_size, long dst_outputs, long offset, float *dst)
{
    for (int j = 0; j < batch; ++j)
        memcpy(dst + offset + j * dst_outputs, src + j * src_size, sizeof(float) * src_size);
}

/* Input pipeline of the reference app: load_image_stb's u8 HWC -> planar float /255.
 * (additionally.c:3080-3103) followed by resize_image (additionally.c:3021-3064): two-pass bilinear, first along x
 * into `part`, then along y; every product and sum rounded to float as written there; the last column / row copy the
 * source edge.  No resize when the size already matches (load_image :3066-3078). */
void yo_load_resize_u8(const unsigned char *data, int w, int h, int c, int out_w, int out_h, float *out)
{
    float *im = (float *)malloc(sizeof(float) * (size_t)w * h * c);
    for (int k = 0; k < c; ++k)
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i)
                im[i + w * j + (size_t)w * h * k] = (float)((double)(float)data[k + c * i + c * w * j] / 255.);
    if (!(out_h && out_w) || (out_h == h && out_w == w)) {
        memcpy(out, im, sizeof(float) * (size_t)w * h * c);
        free(im);
        return;
    }
    float *part = (float *)malloc(sizeof(float) * (size_t)out_w * h * c);
    const float w_scale = (float)(w - 1) / (out_w - 1);
    const float h_scale = (float)(h - 1) / (out_h - 1);
    for (int k = 0; k < c; ++k)
        for (int r = 0; r < h; ++r)
            for (int cc = 0; cc < out_w; ++cc) {
                float val;
                if (cc == out_w - 1 || w == 1) {
                    val = im[(w - 1) + w * r + (size_t)w * h * k];
                } else {
                    const float sx = cc * w_scale;
                    const int ix = (int)sx;
                    const float dx = sx - ix;
                    val = (1 - dx) * im[ix + w * r + (size_t)w * h * k] + dx * im[ix + 1 + w * r + (size_t)w * h * k];
                }
                part[cc + out_w * r + (size_t)out_w * h * k] = val;
            }
    for (int k = 0; k < c; ++k)
        for (int r = 0; r < out_h; ++r) {
            const float sy = r * h_scale;
            const int iy = (int)sy;
            const float dy = sy - iy;
            for (int cc = 0; cc < out_w; ++cc) {
                float val = (1 - dy) * part[cc + out_w * iy + (size_t)out_w * h * k];
                if (!(r == out_h - 1 || h == 1)) val += dy * part[cc + out_w * (iy + 1) + (size_t)out_w * h * k];
                out[cc + out_w * r + (size_t)out_w * out_h * k] = val;
            }
        }
    free(part);
    free(im);
}
