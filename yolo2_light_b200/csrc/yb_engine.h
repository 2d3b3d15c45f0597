// yb_engine.h -- device-side execution plan of one prepared network ("engine"): compiled once per
// (network, rule) from the host model, then replayed as a CUDA graph.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "yb_model.h"

namespace yb {

enum OpKind {
    OP_INPUT = 0, OP_CONV_SIMT = 1, OP_CONV_TC = 2, OP_BINARIZE = 3, OP_CONV_XNOR = 4, OP_QUANTIZE = 5,
    OP_CONV_INT8 = 6, OP_MAXPOOL = 7, OP_UPSAMPLE = 8, OP_SHORTCUT = 9, OP_ROUTE_COPY = 10, OP_REORG = 11,
    OP_YOLO = 12, OP_REGION = 13, OP_CONV_TC_I8 = 14, OP_CONV_TC2 = 15,  // k_conv_tc<2>: CTA pairs (cta_group::2)
    OP_CONV_TC_TF32 = 16
};

struct EngineOptions {
    int device = 0;
    int precision = YB_PREC_BF16_TC;
    bool qrule = false;        // yolov2_forward_network_q layer rule
    bool fuse = true;          // conv + shortcut fusion, route aliasing
    bool upload = true;        // upload the weight arena (false on non-root ranks before the broadcast)
    int q_index_offset = 0;    // added to the layer index in the `i >= 1` INT8 rule (single-layer runs)
    bool keep_counts = false;  // keep raw XNOR popcounts / INT8 accumulators (tests)
    bool ksplit = false;       // K-split of the tail wave of the tensor-core convolutions (opt-in, see yb_model.h)
};

struct Engine;
std::shared_ptr<Engine> build_engine(Network *net, const EngineOptions &opt);
// d_input == nullptr: use the engine's staging buffer (filled by engine_upload_input)
void engine_upload_input(Engine *e, const float *host_input, void *stream);
void engine_upload_u8(Engine *e, const unsigned char *host_u8, int w, int h, int c, int net_w, int net_h, void *stream);
void engine_forward(Engine *e, const void *d_input, void *stream);
void engine_download_outputs(Engine *e, Network *net, void *stream);   // async D2H into pinned, then sync
int engine_submit(Engine *e, const float *host_input);
void engine_collect(Engine *e, Network *net, int ticket);
void engine_collect_ptrs(Engine *e, int ticket, std::vector<const float *> &ptrs, std::vector<size_t> &counts);
const char *engine_broadcast_arena(const std::vector<Engine *> &replicas);   // "nccl" | "peer-copy" | "single"
int engine_device_count();
// pipelined u8 frames -> detections (device-side resize, forward, decode + NMS; only candidate rows come back)
int engine_submit_u8(Engine *e, Network *net, const unsigned char *host_u8, int w, int h, float thresh, float nms,
                     int relative, int letter, int max_rows);
int engine_collect_detections(Engine *e, int ticket, const float **rows, const int **counts, size_t *d2h_bytes);
void engine_fetch_layer(Engine *e, Network *net, int layer, float *dst);
void engine_fetch_input(Engine *e, float *dst);
int engine_fetch_counts(Engine *e, int layer, int32_t *dst, size_t count);
void engine_weight_arena(Engine *e, void **ptr, size_t *bytes);
int engine_detect(Engine *e, Network *net, int w, int h, float thresh, float nms, int relative, int letter,
                  float *rows, int max_rows, int *counts);   // device-side decode + NMS of the whole batch
void engine_input_histogram(Engine *e, Network *net, int layer, int img, float bin_width, int max_bin, uint32_t *hist);
int engine_num_launches(Engine *e);
long engine_info(Engine *e, const char *key);   // "launches", "tc_layers", "ksplit_layers"; -1 unknown
int engine_profile(Engine *e, const void *d_input, int *layer_idx, int *op_kind, float *ms, int max);
void *engine_stream(Engine *e);
const char *op_kind_name(int k);

}  // namespace yb
