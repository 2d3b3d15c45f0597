/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin plain-pointer C ABI around the *unmodified* reference implementation
 * (AlexeyAB/yolo2_light, compiled where it lies under /root/reference/src by
 * oracle/Makefile into oracle/_ref/).  It lets the Python tests drive the
 * reference's own CPU path -- parse_network_cfg / load_weights_upto_cpu /
 * yolov2_fuse_conv_batchnorm / calculate_binary_weights /
 * quantinization_and_get_multipliers / network_predict_cpu /
 * network_predict_quantized -- and read back every layer's tensors, so the
 * oracle restatement (oracle/yolo_oracle.c), the host logic and the CUDA path
 * can be pinned against the real thing.
 *
 * This file is our own code; it only #includes the reference's public header
 * (src/additionally.h) at build time.  Nothing from the reference is copied.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <fcntl.h>

#include "additionally.h"

extern int gpu_index;   /* additionally.c:22 */

/* layer forwards that the reference defines (non-static) but does not declare in its header */
void forward_maxpool_layer_cpu(const layer l, network_state state);   /* yolov2_forward_network.c:268 */
void forward_route_layer_cpu(const layer l, network_state state);     /* :318 */
void forward_reorg_layer_cpu(const layer l, network_state state);     /* :337 */
void forward_upsample_layer_cpu(const layer l, network_state net);    /* :397 */
void forward_shortcut_layer_cpu(const layer l, network_state state);  /* :443 */
void forward_yolo_layer_cpu(const layer l, network_state state);      /* :453 */
void forward_region_layer_cpu(const layer l, network_state state);    /* :511 */
void forward_convolutional_layer_q(layer l, network_state state);     /* yolov2_forward_network_quantized.c:527 */
void calculate_binary_weights(network net);                           /* additionally.c:306 */
void do_nms_sort(detection *dets, int total, int classes, float thresh); /* box.c:296 */

typedef struct refh {
    network net;
    int quantized;
} refh;

/* the reference chats on stdout/stderr (layer table, per-conv printf in the INT8 path) */
static int g_quiet = 1;
static int saved_out = -1, saved_err = -1;
static void hush(void)
{
    if (!g_quiet) return;
    fflush(stdout); fflush(stderr);
    int nul = open("/dev/null", O_WRONLY);
    saved_out = dup(1); saved_err = dup(2);
    dup2(nul, 1); dup2(nul, 2);
    close(nul);
}
static void unhush(void)
{
    if (!g_quiet || saved_out < 0) return;
    fflush(stdout); fflush(stderr);
    dup2(saved_out, 1); dup2(saved_err, 2);
    close(saved_out); close(saved_err);
    saved_out = saved_err = -1;
}

void refh_set_quiet(int q) { g_quiet = q; }

/* prep bit 1: fuse BN, bit 2: binary weights, bit 4: int8 quantisation (only if quantized) --
 * the exact sequence of main.c:160-171 when prep == 7. */
refh *refh_create(const char *cfg, const char *weights, int batch, int quantized, int prep)
{
    gpu_index = -1;     /* main.c:656 (CPU build) -- also keeps additionally.c:254 from inverting the bit weights */
    refh *h = (refh *)calloc(1, sizeof(refh));
    hush();
    h->net = parse_network_cfg((char *)cfg, batch, quantized);
    if (weights && weights[0]) load_weights_upto_cpu(&h->net, (char *)weights, h->net.n);
    if (prep & 1) yolov2_fuse_conv_batchnorm(h->net);
    if (prep & 2) calculate_binary_weights(h->net);
    if (quantized && (prep & 4)) quantinization_and_get_multipliers(h->net);
    unhush();
    h->quantized = quantized;
    return h;
}

int refh_num_layers(refh *h) { return h->net.n; }

void refh_net_ints(refh *h, int *o)
{
    o[0] = h->net.n; o[1] = h->net.batch; o[2] = h->net.h; o[3] = h->net.w; o[4] = h->net.c;
    o[5] = h->net.inputs; o[6] = h->net.outputs; o[7] = h->net.input_calibration_size; o[8] = h->net.quantized;
}

float *refh_net_input_calibration(refh *h) { return h->net.input_calibration; }

/* 40 ints of per-layer metadata */
void refh_layer_ints(refh *h, int i, int *o)
{
    layer *l = &h->net.layers[i];
    memset(o, 0, 40 * sizeof(int));
    o[0] = (int)l->type;   o[1] = (int)l->activation; o[2] = l->batch_normalize; o[3] = l->batch;
    o[4] = l->h; o[5] = l->w; o[6] = l->c; o[7] = l->n; o[8] = l->size; o[9] = l->stride; o[10] = l->pad;
    o[11] = l->out_h; o[12] = l->out_w; o[13] = l->out_c; o[14] = l->inputs; o[15] = l->outputs;
    o[16] = l->xnor; o[17] = l->binary; o[18] = l->quantized; o[19] = l->index;
    o[20] = l->classes; o[21] = l->coords; o[22] = l->softmax; o[23] = l->total; o[24] = l->reverse;
    o[25] = l->lda_align; o[26] = l->new_lda; o[27] = l->bit_align; o[28] = l->align_bit_weights_size;
    o[29] = l->dontload; o[30] = l->dontloadscales; o[31] = l->use_bin_output; o[32] = l->max_boxes;
    o[33] = l->groups;
}

void refh_layer_floats(refh *h, int i, float *o)
{
    layer *l = &h->net.layers[i];
    o[0] = l->weights_quant_multipler; o[1] = l->input_quant_multipler; o[2] = l->output_multipler;
    o[3] = l->scale; o[4] = l->bflops; o[5] = l->temperature;
}

void *refh_layer_ptr(refh *h, int i, const char *what)
{
    layer *l = &h->net.layers[i];
    if (!strcmp(what, "weights")) return l->weights;
    if (!strcmp(what, "biases")) return l->biases;
    if (!strcmp(what, "scales")) return l->scales;
    if (!strcmp(what, "rolling_mean")) return l->rolling_mean;
    if (!strcmp(what, "rolling_variance")) return l->rolling_variance;
    if (!strcmp(what, "output")) return l->output;
    if (!strcmp(what, "weights_int8")) return l->weights_int8;
    if (!strcmp(what, "biases_quant")) return l->biases_quant;
    if (!strcmp(what, "align_bit_weights")) return l->align_bit_weights;
    if (!strcmp(what, "binary_weights")) return l->binary_weights;
    if (!strcmp(what, "mean_arr")) return l->mean_arr;
    if (!strcmp(what, "mask")) return l->mask;
    if (!strcmp(what, "input_layers")) return l->input_layers;
    if (!strcmp(what, "input_sizes")) return l->input_sizes;
    return NULL;
}

/* whole-network forward through the reference's own entry points (main.c:199-219) */
float *refh_predict(refh *h, float *input)
{
    float *r;
    hush();
    if (h->quantized) r = network_predict_quantized(h->net, input);
    else r = network_predict_cpu(h->net, input);
    unhush();
    return r;
}

/* one layer, given its input tensor (NCHW, batch-major), following the dispatch of
 * yolov2_forward_network_cpu (:581) or yolov2_forward_network_q (:1027) when use_q_rule != 0.
 * ROUTE/SHORTCUT read their sources from net.layers[*].output, so those must have been
 * produced by a previous refh_predict / refh_forward_layer call. */
void refh_forward_layer(refh *h, int i, float *input, int use_q_rule)
{
    network_state state;
    memset(&state, 0, sizeof(state));
    state.net = h->net;
    state.index = i;
    state.input = input;
    state.workspace = h->net.workspace;
    layer l = h->net.layers[i];
    hush();
    switch (l.type) {
    case CONVOLUTIONAL:
        if (use_q_rule && i >= 1 && l.activation != LINEAR) forward_convolutional_layer_q(l, state);
        else forward_convolutional_layer_cpu(l, state);
        break;
    case MAXPOOL: forward_maxpool_layer_cpu(l, state); break;
    case ROUTE: forward_route_layer_cpu(l, state); break;
    case REORG: forward_reorg_layer_cpu(l, state); break;
    case UPSAMPLE: forward_upsample_layer_cpu(l, state); break;
    case SHORTCUT: forward_shortcut_layer_cpu(l, state); break;
    case YOLO: forward_yolo_layer_cpu(l, state); break;
    case REGION: forward_region_layer_cpu(l, state); break;
    default: break;
    }
    unhush();
}

/* wall-clock seconds per call of the reference predict (the reference's own clock() is CPU time) */
double refh_time_predict(refh *h, float *input, int reps)
{
    struct timespec t0, t1;
    hush();
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; ++r) {
        if (h->quantized) network_predict_quantized(h->net, input);
        else network_predict_cpu(h->net, input);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    unhush();
    return ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / (reps > 0 ? reps : 1);
}

/* detections through the reference's get_network_boxes (additionally.c:4403) + do_nms_sort (box.c:296).
 * Fills out[k*(6+classes)] = {x,y,w,h,objectness,sort_class, prob[0..classes)}; returns count. */
int refh_get_boxes(refh *h, int w, int hgt, float thresh, float nms, float *out, int max_out, int *classes_out)
{
    int nboxes = 0;
    detection *dets = get_network_boxes(&h->net, w, hgt, thresh, 0.5f, 0, 1, &nboxes, 0);
    layer l = h->net.layers[h->net.n - 1];
    if (nms > 0) do_nms_sort(dets, nboxes, l.classes, nms);
    int stride = 6 + l.classes;
    int n = nboxes < max_out ? nboxes : max_out;
    for (int k = 0; k < n; ++k) {
        float *o = out + (size_t)k * stride;
        o[0] = dets[k].bbox.x; o[1] = dets[k].bbox.y; o[2] = dets[k].bbox.w; o[3] = dets[k].bbox.h;
        o[4] = dets[k].objectness; o[5] = (float)dets[k].sort_class;
        for (int c = 0; c < l.classes; ++c) o[6 + c] = dets[k].prob[c];
    }
    if (classes_out) *classes_out = l.classes;
    free_detections(dets, nboxes);
    return nboxes;
}

/* direct access to the reference math kernels used as kernel-level oracles */
void refh_gemm_nn(int M, int N, int K, float *A, int lda, float *B, int ldb, float *C, int ldc)
{
    gemm_nn(M, N, K, 1.0f, A, lda, B, ldb, C, ldc);   /* additionally.c:1272 */
}
void refh_im2col(float *im, int c, int hgt, int w, int k, int stride, int pad, float *col)
{
    im2col_cpu(im, c, hgt, w, k, stride, pad, col);   /* additionally.c:39 */
}

/* INT8 input calibration: the reference's KL search, yolov2_forward_network_quantized.c:1292-1398,
 * with the arguments network_calibrate_cpu uses (yolov2_forward_network.c:784: bin width 1/16, 4096 bins). */
float entropy_calibration(float *src_arr, const size_t size, const float bin_width, const int max_bin);
float refh_entropy_calibration(float *src, long size, float bin_width, int max_bin)
{
    return entropy_calibration(src, (size_t)size, bin_width, max_bin);
}

/* mAP driver of the reference: validate_detector_map, additionally.c:4541-4898, with its stdout (the only
 * place the numbers go) redirected into `out_path`. */
void validate_detector_map(char *datacfg, char *cfgfile, char *weightfile, float thresh_calc_avg_iou, int quantized,
                           const float iou_thresh);
void refh_validate_map(const char *datacfg, const char *cfg, const char *weights, float thresh_calc_avg_iou, int quantized,
                       float iou_thresh, const char *out_path)
{
    fflush(stdout);
    int saved = dup(1);
    FILE *f = fopen(out_path, "w");
    if (!f) return;
    dup2(fileno(f), 1);
    validate_detector_map((char *)datacfg, (char *)cfg, (char *)weights, thresh_calc_avg_iou, quantized, iou_thresh);
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    fclose(f);
}

/* Image pipeline of the reference app: u8 HWC (what stbi_load returns) -> planar float /255. (load_image_stb,
 * additionally.c:3080-3103) -> resize_image bilinear to the network size (additionally.c:3021-3064, only when the
 * size differs, load_image :3066-3078).  out: float[c*out_h*out_w]. */
void refh_load_resize_u8(const unsigned char *data, int w, int h, int c, int out_w, int out_h, float *out)
{
    image im = make_image(w, h, c);
    int i, j, k;
    for (k = 0; k < c; ++k)
        for (j = 0; j < h; ++j)
            for (i = 0; i < w; ++i)
                im.data[i + w * j + w * h * k] = (float)data[k + c * i + c * w * j] / 255.;
    if ((out_h && out_w) && (out_h != im.h || out_w != im.w)) {
        image resized = resize_image(im, out_w, out_h);
        free_image(im);
        im = resized;
    }
    memcpy(out, im.data, sizeof(float) * (size_t)im.w * im.h * im.c);
    free_image(im);
}

#ifdef YB_DROPIN
/* drop-in check: the glue of integration/yolo2_light_b200_glue.c behind the reference's own host code */
float *network_predict_b200(network net, float *input);
float *network_predict_b200_quantized(network net, float *input);
float *refh_predict_b200(refh *h, float *input)
{
    return h->quantized ? network_predict_b200_quantized(h->net, input) : network_predict_b200(h->net, input);
}
/* the reference's per-layer call shape (by-value layer + network_state) served by the engine: the conv of layer i alone */
void forward_convolutional_layer_b200(layer l, network_state state);
void forward_convolutional_layer_b200_q(layer l, network_state state);
void refh_forward_conv_b200(refh *h, int i, float *input, int use_q_rule)
{
    network_state state;
    memset(&state, 0, sizeof(state));
    state.net = h->net;
    state.index = i;
    state.input = input;
    state.workspace = h->net.workspace;
    layer l = h->net.layers[i];
    if (use_q_rule && i >= 1 && l.activation != LINEAR) forward_convolutional_layer_b200_q(l, state);   /* the loop's rule, :1036 */
    else forward_convolutional_layer_b200(l, state);
}
/* batch extension: nimg images over ngpus GPUs from this C process; returns the last layer's nimg x outputs floats */
float *network_predict_b200_batch(network net, float *images, int nimg, int ngpus, int quantized);
float *refh_predict_b200_batch(refh *h, float *images, int nimg, int ngpus)
{
    return network_predict_b200_batch(h->net, images, nimg, ngpus, h->quantized);
}
/* wall-clock seconds per call of the drop-in pair predict + (optionally) device decode, as main.c:199-229 would run it */
detection *get_network_boxes_nms_b200(network *net, int w, int h, float thresh, float nms, int relative, int *num, int letter);
double refh_time_predict_b200(refh *h, float *input, int reps, int decode, float thresh, float nms)
{
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; ++r) {
        if (h->quantized) network_predict_b200_quantized(h->net, input);
        else network_predict_b200(h->net, input);
        if (decode) {
            int nboxes = 0;
            detection *dets = get_network_boxes_nms_b200(&h->net, h->net.w, h->net.h, thresh, nms, 1, &nboxes, 0);
            free_detections(dets, nboxes);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    return ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / (reps > 0 ? reps : 1);
}
/* the glue's device-side decode + NMS, reported in refh_get_boxes' row format */
int refh_get_boxes_b200(refh *h, int w, int hgt, float thresh, float nms, float *out, int max_out)
{
    int nboxes = 0, k, c;
    detection *dets = get_network_boxes_nms_b200(&h->net, w, hgt, thresh, nms, 1, &nboxes, 0);
    layer l = h->net.layers[h->net.n - 1];
    int stride = 6 + l.classes;
    int n = nboxes < max_out ? nboxes : max_out;
    for (k = 0; k < n; ++k) {
        float *o = out + (size_t)k * stride;
        o[0] = dets[k].bbox.x; o[1] = dets[k].bbox.y; o[2] = dets[k].bbox.w; o[3] = dets[k].bbox.h;
        o[4] = dets[k].objectness; o[5] = 0;
        for (c = 0; c < l.classes; ++c) o[6 + c] = dets[k].prob[c];
    }
    free_detections(dets, nboxes);
    return nboxes;
}
#endif
