/*
 * yolo2_light_b200.h -- C ABI of libyolo2_light_b200.so
 *
 * A Blackwell (sm_100a) forward-inference engine for darknet YOLO v2/v3 networks that sits behind the C surface
 * of AlexeyAB/yolo2_light.  Every entry point below names the reference interface it replaces (file:line relative
 * to the reference tree).  Plain pointers and sizes only; no CUDA, torch or C++ types cross this boundary.
 *
 * Two ways in:
 *   (1) stand-alone: yb_parse_network_cfg -> yb_load_weights_upto -> yb_fuse_conv_batchnorm ->
 *       yb_calculate_binary_weights -> [yb_quantinization_and_get_multipliers] -> yb_network_predict*
 *       (the exact call sequence of the reference app, src/main.c:160-219);
 *   (2) drop-in behind the reference's own parser/loader: the host program keeps its `network` and hands the
 *       prepared per-layer arrays over as yb_layer_desc[] (yb_network_from_layers); see INTEGRATION.md for the
 *       ~60-line glue file (`network_predict_b200(network net, float *input)`).
 *
 * Error convention: the reference has no status codes -- it prints and exits (additionally.c:1595-1614,
 * gpu.cu:58-83).  Default here is the same (message on stderr + abort()).  Hosts that prefer to recover call
 * yb_set_abort_on_error(0): failing calls then return NULL / non-zero and yb_last_error() holds the message.
 *
 * There is NO CPU fallback: every predict/forward entry point requires a CUDA device of compute capability 10.x
 * and fails loudly without one.
 */
#ifndef YOLO2_LIGHT_B200_H
#define YOLO2_LIGHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Numeric values are the reference's own enums so that glue code can pass `l.type` / `l.activation` through:
 * LAYER_TYPE (src/additionally.h:376-403), ACTIVATION (src/additionally.h:68-70). */
enum {
    YB_CONVOLUTIONAL = 0, YB_MAXPOOL = 3, YB_SOFTMAX = 4, YB_ROUTE = 8, YB_SHORTCUT = 13,
    YB_REGION = 21, YB_YOLO = 22, YB_UPSAMPLE = 23, YB_REORG = 24, YB_BLANK = 25
};
enum { YB_LOGISTIC = 0, YB_RELU = 1, YB_LINEAR = 3, YB_LEAKY = 7 };

/* Arithmetic used for the FP32-variant convolutions (yolov2_forward_network.c:204-211).
 *  YB_PREC_BF16_TC : bf16 operands, f32 accumulation on tcgen05 tensor cores, bf16 NHWC activations (default)
 *  YB_PREC_FP32    : f32 operands and accumulation on CUDA cores, f32 activations (validation / exact nets)
 * Networks that contain XNOR layers, and every network run through the -quantized rule, always keep f32
 * activations so the integer paths see exactly the reference's inputs. */
enum { YB_PREC_BF16_TC = 0, YB_PREC_FP32 = 1 };

/* One layer of a prepared network: the subset of the reference's `struct layer` (src/additionally.h:409-684)
 * that the forward path reads.  All pointers are host pointers owned by the caller; the
 * library copies what it needs. */
typedef struct yb_layer_desc {
    int type;                 /* YB_* layer type                                   (layer.type)        */
    int activation;           /* YB_* activation                                   (layer.activation)  */
    int batch_normalize;      /* non-zero: BN not folded yet                       (layer.batch_normalize) */
    int h, w, c;              /* input tensor                                       (layer.h/w/c)       */
    int n;                    /* conv: filters; route: #inputs; yolo/region: anchors (layer.n)          */
    int size, stride, pad;    /* conv/maxpool geometry (maxpool pad = cfg `padding`) (layer.size/stride/pad) */
    int out_h, out_w, out_c;  /* output tensor                                      (layer.out_*)       */
    int xnor;                 /* conv: BIT1-XNOR variant                            (layer.xnor)        */
    int quantized;            /* conv: parser's per-layer INT8 flag (informational)  (layer.quantized)   */
    int index;                /* shortcut: absolute index of the `from` layer        (layer.index)       */
    int classes, coords, softmax, total;   /* yolo/region                            (layer.classes ...) */
    int reverse;              /* upsample/reorg                                      (layer.reverse)     */
    float scale;              /* upsample                                            (layer.scale)       */
    const int *input_layers;  /* route: n absolute layer indices                     (layer.input_layers) */
    const int *mask;          /* yolo: n anchor ids                                  (layer.mask)        */
    const float *anchors;     /* yolo: 2*total, region: 2*n                          (layer.biases)      */
    const float *weights;     /* conv: [n][c][size][size]                            (layer.weights)     */
    const float *biases;      /* conv: [n]                                           (layer.biases)      */
    const float *scales, *rolling_mean, *rolling_variance;   /* conv with BN: [n]                         */
    const int8_t *weights_int8;       /* conv, after quantisation: [n][c][size][size] (layer.weights_int8) */
    float weights_quant_multipler;    /*                                     (layer.weights_quant_multipler) */
    float input_quant_multipler;      /*                                     (layer.input_quant_multipler)   */
    const float *mean_arr;    /* conv xnor, after calculate_binary_weights: [n]      (layer.mean_arr)    */
} yb_layer_desc;

/* Host-side model: our equivalent of the reference's `network` (src/additionally.h:703-763). Opaque. */
typedef struct yb_network yb_network;

/* ---- errors ------------------------------------------------------------------------------------------ */
void        yb_set_abort_on_error(int on);   /* default 1 = reference behaviour (print + abort) */
const char *yb_last_error(void);

/* ---- model preparation (host, one-time) ------------------------------------------------------------ */

/* replaces parse_network_cfg(char *filename, int batch, int quantized)   src/additionally.c:3955
 * Same .cfg grammar and defaults (src/additionally.c:3423-3457, :3534-3897); batch>0 overrides the cfg's. */
yb_network *yb_parse_network_cfg(const char *filename, int batch, int quantized);

/* replaces load_weights_upto_cpu(network *net, char *filename, int cutoff)   src/additionally.c:3491
 * Same .weights format (src/additionally.c:3459-3529).  Returns 0 on success. */
int yb_load_weights_upto(yb_network *net, const char *filename, int cutoff);

/* replaces yolov2_fuse_conv_batchnorm(network net)   src/additionally.c:67 */
void yb_fuse_conv_batchnorm(yb_network *net);

/* replaces calculate_binary_weights(network net)   src/additionally.c:306  (binarize_weights :113,
 * mean_arr :188): per-filter mean |w| and sign bits for every xnor=1 convolution. */
void yb_calculate_binary_weights(yb_network *net);

/* replaces quantinization_and_get_multipliers(network net)   src/yolov2_forward_network_quantized.c:1402 */
void yb_quantinization_and_get_multipliers(yb_network *net);

/* Drop-in path: build a yb_network from layers prepared by the reference's own host code (after the
 * main.c:160-171 sequence).  dims = {batch, h, w, c}.  Arrays are copied. */
yb_network *yb_network_from_layers(const yb_layer_desc *layers, int n_layers, int batch, int h, int w, int c,
                                   int quantized);

void yb_free_network(yb_network *net);   /* free_network, src/additionally.c:2058 */

/* ---- introspection ----------------------------------------------------------------------------------- */
int  yb_network_num_layers(const yb_network *net);
/* out[0..8) = {n_layers, batch, h, w, c, inputs, outputs (last layer), input_calibration_size} */
void yb_network_dims(const yb_network *net, int *out8);
/* Fills *out with layer i; pointers alias memory owned by net (valid until yb_free_network). */
int  yb_network_layer(const yb_network *net, int i, yb_layer_desc *out);
const float *yb_network_input_calibration(const yb_network *net, int *count);
/* Change the batch size of a parsed network (set_batch_network, src/additionally.c:2038). Drops any engine. */
void yb_set_batch_network(yb_network *net, int batch);

/* ---- forward (device) -------------------------------------------------------------------------------- */

/* Select the device (cuda_set_device, src/gpu.cu:97) and the FP32-conv arithmetic for engines built later. */
int  yb_network_set_device(yb_network *net, int device);
int  yb_network_set_precision(yb_network *net, int precision /* YB_PREC_* */);

/* replaces network_predict_cpu(network net, float *input)   src/yolov2_forward_network.c:632
 * (same slot as network_predict_gpu_cudnn, src/yolov2_forward_network_gpu.cu:547).
 * input: host, NCHW float, net.batch images of c*h*w in [0,1].  Returns the last layer's host output (owned by
 * net); every YOLO/REGION layer's host output is filled (yb_network_layer_output) so that box decoding works
 * exactly as after the reference call (additionally.c:4391-4398).  Unlike the reference's decoder the outputs
 * of ALL batch items are produced. */
float *yb_network_predict(yb_network *net, const float *input);

/* replaces network_predict_quantized(network net, float *input)   src/yolov2_forward_network_quantized.c:1160
 * INT8 rule of yolov2_forward_network_q (:1036): conv i uses the s8 x s8 -> s32 path iff i >= 1 and its
 * activation is not LINEAR; everything else as in yb_network_predict with f32 activations. */
float *yb_network_predict_quantized(yb_network *net, const float *input);

/* Input pipeline on the device: replaces load_image_stb's u8 -> float/255 conversion
 * (src/additionally.c:3080-3103) + resize_image (src/additionally.c:3021-3064) + network_predict_*.
 * images_hwc: net.batch interleaved 8-bit images (HWC, net.c channels, as stbi_load returns them), all w x h.
 * The bilinear resize to the network size is bit-identical to the reference's (scalar build). */
float *yb_network_predict_image_u8(yb_network *net, const unsigned char *images_hwc, int w, int h, int quantized);
/* Diagnostic: the planar float input (batch*c*h*w) the device pipeline produced for the last call. */
int    yb_network_fetch_input(yb_network *net, int quantized, float *dst);

/* Pipelined form of the two calls above for throughput serving: yb_network_submit enqueues one batch (H2D of
 * `input` on a copy stream, the forward on the compute stream, D2H of the yolo/region tensors on a third stream)
 * and returns a ticket immediately; yb_network_collect blocks until that batch is done and points the layers'
 * host outputs at its results (valid until the ticket's slot is reused, i.e. for the next 2 submits).  Up to 3
 * batches may be in flight, so the copies of batch k+1 / k-1 overlap the compute of batch k.  `input` should be
 * pinned (yb_alloc_pinned) and must stay untouched until its ticket has been collected. */
int yb_network_submit(yb_network *net, const float *input, int quantized);
int yb_network_collect(yb_network *net, int ticket, int quantized);

/* The serving loop of the reference app (src/main.c:188-229: load_image + resize_image, network_predict*, get_network_boxes,
 * do_nms_sort) as ONE pipelined call per batch: yb_network_submit_u8 enqueues the H2D of net.batch 8-bit HWC frames (all
 * w x h) and the reference's bilinear resize on a copy stream, the forward on the compute stream and the decode + NMS of the
 * whole batch (the arithmetic of yb_network_detect) on a side stream where it runs under the NEXT batch's forward; it returns
 * a ticket at once.  yb_network_collect_detections blocks until that batch is decoded and copies back exactly its candidate
 * rows: *rows = pinned float[batch][max_rows][5 + classes] owned by the library (valid until the ticket's slot is reused, i.e.
 * for the next 2 submits), *counts = int[batch] candidates per image before the max_rows cap (max_rows <= 16384),
 * *d2h_bytes (optional) = bytes that crossed PCIe for this ticket.  Returns 5 + classes, or -1.  Up to 3 batches in flight.
 * Per batch of 16 608x608 frames that is 17.7 MB in and < 1 MB out instead of 71 MB in / 124 MB out for the raw tensors. */
int yb_network_submit_u8(yb_network *net, const unsigned char *images_hwc, int w, int h, int quantized, float thresh,
                         float nms, int relative, int letter, int max_rows);
int yb_network_collect_detections(yb_network *net, int ticket, int quantized, const float **rows, const int **counts,
                                  size_t *d2h_bytes);

/* Host output (NCHW for yolo, HWC-flattened for region, as the reference lays them out) of layer i after a
 * predict call; only YOLO/REGION layers (and the last layer) are kept on the host. */
const float *yb_network_layer_output(const yb_network *net, int i, int *count);

/* Device-resident variant for pipelines that already hold their images in HBM (and for the bench's `value`):
 * d_input = device pointer to net.batch NCHW float images; stream = cudaStream_t (or NULL).  Enqueues the whole
 * forward; results stay on the device until yb_network_sync_outputs(). quantized selects the INT8 rule. */
int yb_network_forward_device(yb_network *net, const void *d_input, int quantized, void *stream);
int yb_network_sync_outputs(yb_network *net, int quantized, void *stream);   /* D2H of yolo/region tensors + sync */

/* Test/diagnostic hook: copy ANY layer's activation back as NCHW float (batch-major), whatever its device
 * layout/dtype.  dst must hold batch*out_c*out_h*out_w floats (region: batch*outputs). */
int yb_network_fetch_layer(yb_network *net, int i, int quantized, float *dst);

/* replaces forward_convolutional_layer_cpu(layer l, network_state state)  src/yolov2_forward_network.c:30 and
 * forward_convolutional_layer_q(layer l, network_state state)  src/yolov2_forward_network_quantized.c:527.
 * Runs conv layer `i` of net alone on `input` (host NCHW, batch*c*h*w) and writes host NCHW `output`
 * (batch*n*out_h*out_w).  variant: 0 = as yb_network_predict would run it, 1 = as the quantized rule would. */
int yb_forward_convolutional_layer(yb_network *net, int i, int variant, const float *input, float *output);

/* ---- multi-GPU batch extension (SURVEY 8b "Batch extension", 8e) ------------------------------------------------------
 * The reference runs one image per call on one device (src/main.c:199-219, cuda_set_device src/gpu.cu:97-102).  From the same
 * plain-C host program -- one process, no Python, no launcher -- yb_network_predict_batch runs `nimg` images (host NCHW float,
 * nimg * c*h*w) through `ngpus` replicas of the engine: contiguous shards of net.batch whole images go round-robin to the
 * replicas (pipelined per GPU like yb_network_submit), the prepared weight arena is built on the first device and reaches the
 * others by ONE ncclBroadcast at the first call (NCCL is bound at run time; without it, or for a device list with repeats,
 * peer copies -- yb_network_replication() says which), and there is no other communication.  A partial last shard is padded
 * with zero images whose results are dropped.  Results: yb_network_batch_output(net, i, &per_image) = host float[nimg][per_image]
 * for every YOLO / REGION layer and the last layer, image k bit-identical to what yb_network_predict returns for it on one GPU.
 * yb_network_set_devices chooses the devices (default 0 .. ngpus-1; repeats allowed: several replicas on one GPU). */
int yb_network_set_devices(yb_network *net, const int *devices, int ndev);
int yb_network_predict_batch(yb_network *net, const float *images, int nimg, int ngpus, int quantized);
const float *yb_network_batch_output(const yb_network *net, int i, int *per_image);
const char *yb_network_replication(const yb_network *net);   /* "nccl" | "peer-copy" | "single" | "" (not replicated yet) */

/* Weight arena of the engine (all prepared device-side weights in one allocation) -- what a multi-GPU launcher
 * broadcasts once at init (one process per GPU; the harness uses torch.distributed/NCCL on this pointer).
 * Builds the engine if needed; upload=0 allocates without uploading (non-root ranks). */
int yb_network_weight_arena(yb_network *net, int quantized, int upload, void **d_ptr, size_t *bytes);

/* Number of kernels enqueued by the last forward. */
int  yb_network_last_launches(const yb_network *net);

/* Diagnostic switches (tests): "fuse" (1: conv+shortcut fusion and route aliasing, default), "keep_counts"
 * (1: keep the raw XNOR popcounts / INT8 s32 accumulators of every integer conv), "q_index_offset", "ksplit"
 * (1: split the tail wave of the deep-K tensor-core convolutions along K: ~1 % faster on yolov3-608 b16, but the f32
 * summation order then depends on the batch size; default 0 keeps image k of any batch bit-identical to a batch of 1). */
int  yb_network_set_option(yb_network *net, const char *name, int value);
/* Engine facts (builds the engine if needed): "launches", "tc_layers" (convolutions on tcgen05), "ksplit_layers"
 * (of those, how many run with a K-split tail wave).  -1: unknown key. */
long yb_network_get_info(yb_network *net, int quantized, const char *key);
/* Raw integer results of conv layer i (NCHW, batch-major) when "keep_counts" is on; returns the element count. */
int  yb_network_fetch_counts(yb_network *net, int i, int quantized, int32_t *dst, size_t count);
int  yb_network_layer_outputs(const yb_network *net, int i);   /* layer.outputs (per image) */
const char *yb_op_kind_name(int kind);

/* Pinned host buffers for the end-to-end path (cudaHostAlloc / cudaFreeHost). */
void *yb_alloc_pinned(size_t bytes);
void  yb_free_pinned(void *p);

/* Per-op timing of one forward (CUDA events around every kernel; diagnostic, not for benchmarks).
 * Fills up to max entries: layer index, op kind code, milliseconds. Returns the number of ops. */
int yb_network_profile(yb_network *net, int quantized, const void *d_input, int *layer_idx, int *op_kind,
                       float *ms, int max);

/* ---- detection decode (host) -------------------------------------------------------- */

/* replaces get_network_boxes + do_nms_sort   src/additionally.c:4403, src/box.c:296 for batch item b.
 * out rows: {x, y, w, h, objectness, prob[classes]}; returns the number of rows written (<= max_rows).
 * Candidates with EQUAL class probability are ranked by their position in the candidate list (layer, cell, anchor), here and
 * in yb_network_detect; the reference leaves their order to qsort (box.c:311), i.e. to the C library. */
int yb_get_network_boxes(const yb_network *net, int b, int w, int h, float thresh, float nms, int relative,
                         int letter, float *out, int max_rows);

/* ---- detection decode + NMS on the device, whole batch ------------------------------- */

/* Same arithmetic as yb_get_network_boxes / the reference (get_yolo_detections src/additionally.c:4317,
 * custom_get_region_detections :4363, correct_yolo_boxes :4281, do_nms_sort src/box.c:296), run on the yolo / region
 * tensors where the last forward left them in HBM, for every image of the batch (the reference decodes batch item 0
 * only).  rows: host float[batch][max_rows][5 + classes] = {x, y, w, h, objectness, prob[classes]} in the reference's
 * candidate order (layer, cell, anchor); suppressed / below-threshold class entries are 0 like the reference leaves
 * them.  counts[b] = number of candidates of image b; when it exceeds max_rows only the first max_rows candidates were
 * decoded and took part in the NMS (the reference has no cap: size max_rows accordingly, <= 16384).
 * Returns the row length 5 + classes, or -1. */
int yb_network_detect(yb_network *net, int quantized, int w, int h, float thresh, float nms, int relative, int letter,
                      float *rows, int max_rows, int *counts);

/* ---- INT8 input calibration ---------------------------------------------------------- */

/* network_calibrate_cpu (src/yolov2_forward_network.c:731-831) with the forward pass and the |x| histograms of every
 * convolution's input on the GPU and entropy_calibration's KL search (src/yolov2_forward_network_quantized.c:1292-1398)
 * restated on the host.  One call = one batch of calibration images: multipliers[b * nconv + k] is what the reference
 * computes for image b at its k-th CONVOLUTIONAL layer (bin width 1/16, 4096 bins, :784).  Average over images and
 * write them as `input_calibration = m0, m1, ..., 16` into the cfg (:753-769).  Uses the network's precision setting
 * (YB_PREC_FP32 reproduces the reference's float activations); returns nconv, or -1. */
int yb_network_calibrate(yb_network *net, const float *input, float *multipliers, int max_values);
/* entropy_calibration (src/yolov2_forward_network_quantized.c:1292) on a host array: bit-identical multiplier. */
float yb_entropy_calibration(const float *src, size_t size, float bin_width, int max_bin);
/* |x| histogram of the input of layer `layer`, image `img`, after the last forward (diagnostic / tests). */
int yb_network_input_histogram(yb_network *net, int quantized, int layer, int img, float bin_width, int max_bin,
                               uint32_t *hist);

/* ---- mAP accounting ------------------------------------------------------------------ */

/* The bookkeeping of validate_detector_map (src/additionally.c:4541-4898) on detections produced elsewhere: `rows` =
 * the concatenated rows of all images ({x, y, w, h, objectness, prob[classes]}, relative coordinates; what
 * yb_network_detect / yb_get_network_boxes return for w = h = 1, thresh .005, nms .45 as the reference uses),
 * rows_per_image[nimages]; truth[ntruth][6] = {image, class, x, y, w, h} (the label files' content, in file order).
 * Outputs: ap_per_class[classes] (11-point interpolated AP, :4848), *map_out, stats[8] = {precision, recall, F1,
 * average IoU, TP, FP, FN at thresh_calc_avg_iou (:4872-4880), number of detections}.  Returns the number of
 * (box, class) detections ranked, or -1.  The "difficult" list is not modelled. */
int yb_map_evaluate(const float *rows, const int *rows_per_image, int nimages, int classes, const float *truth, int ntruth,
                    float iou_thresh, float thresh_calc_avg_iou, double *ap_per_class, double *map_out, float *stats);

const char *yb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* YOLO2_LIGHT_B200_H */
