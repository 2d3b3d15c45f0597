// yb_engine.cu -- builds and runs the device execution plan of a prepared network.
//
// Mirrors the layer loop of the reference (yolov2_forward_network_cpu, src/yolov2_forward_network.c:581-628 and
// yolov2_forward_network_q, src/yolov2_forward_network_quantized.c:1027-1089) as a flat list of kernel
// launches with pre-resolved device pointers, replayed as a CUDA graph.  See yb_kernels.cuh for the device
// layout.  No CPU fallback: everything here requires a compute-capability-10.x device.
#include "yb_engine.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "yb_conv_tc.cuh"
#include "yb_kernels.cuh"
#include "yb_detect.cuh"

namespace yb {

#define CUDA_OK(call)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            fatal_throw(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                        std::to_string(__LINE__) + " (" #call ")");                                     \
    } while (0)

const char *op_kind_name(int k) {
    static const char *names[] = {"input", "conv_simt", "conv_tc", "binarize", "conv_xnor", "quantize",
                                  "conv_int8", "maxpool", "upsample", "shortcut", "route_copy", "reorg",
                                  "yolo", "region", "conv_tc_i8", "conv_tc2", "conv_tc_tf32"};
    return (k >= 0 && k < 17) ? names[k] : "?";
}

enum { DT_F32 = 0, DT_BF16 = 1, DT_S8 = 2, DT_BITS = 3 };
static inline size_t dt_size(int dt) { return dt == DT_F32 ? 4 : dt == DT_BF16 ? 2 : dt == DT_S8 ? 1 : 4; }

struct Op {
    int kind;
    int layer;
    std::function<void(cudaStream_t)> launch;
};

struct ConvWeights {   // offsets into the weight arena
    size_t w_f32 = (size_t)-1, w_bf16 = (size_t)-1, w_s8 = (size_t)-1, w_bits = (size_t)-1;
    size_t w_f32km = (size_t)-1;   // f32 [ldn][K], K-major (kind::tf32)
    size_t bias = (size_t)-1, mean = (size_t)-1;
    int ldw = 0;      // f32 [K][ldw]
    int ldn = 0;      // rows of the [ldn][...] layouts
    int cpad = 0;     // padded channels of the s8 / bits / bf16 layouts
};

struct Engine {
    EngineOptions opt;
    int batch = 0;
    int act_dt = DT_F32;
    cudaStream_t stream = nullptr;
    char *act_arena = nullptr; size_t act_bytes = 0;
    char *w_arena = nullptr;   size_t w_bytes = 0;
    float *d_input = nullptr;  size_t input_count = 0;
    std::vector<TV> out_tv;            // per layer (base == nullptr: no NHWC output)
    std::vector<int> out_dt;
    std::vector<float *> d_final;      // yolo / region device outputs
    std::vector<float *> h_final;      // pinned host mirrors
    std::vector<size_t> final_count;
    std::vector<int32_t *> d_counts;   // optional raw integer results per conv layer
    std::vector<size_t> counts_count;
    std::vector<Op> ops;
    std::vector<char> not_materialised;   // layer outputs that exist only in a consumer's fused form
    TV in0{};
    int in0_dt = DT_F32;
    cudaGraphExec_t graph_exec = nullptr;
    bool graph_failed = false;
    std::vector<void *> tc_plans;      // opaque per-layer state of the tensor-core path (tensor maps)
    // device-side decode + NMS workspace (engine_detect), sized for det_cap rows per image
    int n_tc = 0, n_ksplit = 0;
    float *ksplit_ws = nullptr; unsigned *ksplit_flags = nullptr;   // partial sums / flags of the K-split tail (yb_conv_tc.cu)
    std::function<void(const float *, cudaStream_t)> first_op;   // consumes the caller's NCHW images (pointer varies per call)
    std::function<void(const unsigned char *, cudaStream_t)> first_op_u8;   // same from 8-bit HWC frames of the network size (if set)
    int first_kind = OP_INPUT, first_layer = -1;
    void *stem_plan = nullptr;
    unsigned char *d_u8 = nullptr; size_t u8_bytes = 0;   // staging of the caller's u8 images (device-side input pipeline)
    // ---- pipelined end-to-end path: H2D(k+1) | compute(k) | D2H(k-1) on three streams -----------------
    struct DetWs {                      // decode + NMS workspace for `cap` candidate rows per image
        float *rows = nullptr; unsigned *mask = nullptr; int *blkcnt = nullptr, *counts = nullptr;
        int cap = 0, stride = 0, nblk = 0;
    };
    struct Slot {
        float *d_in = nullptr;
        std::vector<float *> d_out, h_out;
        cudaEvent_t ev_in = nullptr, ev_comp = nullptr, ev_done = nullptr, ev_det = nullptr;
        bool busy = false;
        // device-side input pipeline + decode of the pipelined detection path (engine_submit_u8)
        unsigned char *d_u8 = nullptr; size_t u8_bytes = 0;
        DetWs det;
        float *h_rows = nullptr; size_t h_rows_bytes = 0; int *h_counts = nullptr;
        int mode = 0;                   // 0: raw tensors (engine_submit), 1: detections (engine_submit_u8)
    };
    std::vector<Slot> slots;
    cudaStream_t s_in = nullptr, s_out = nullptr, s_det = nullptr;
    int next_slot = 0;
    DetWs det;                          // workspace of the synchronous engine_detect
    ~Engine();
};

Engine::~Engine() {
    cudaSetDevice(opt.device);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (float *p : h_final) if (p) cudaFreeHost(p);
    for (float *p : d_final) if (p) cudaFree(p);
    for (int32_t *p : d_counts) if (p) cudaFree(p);
    for (void *p : tc_plans) tc_free_plan(p);
    if (stem_plan) tc_stem_free_plan(stem_plan);
    if (det.rows) cudaFree(det.rows);
    if (det.mask) cudaFree(det.mask);
    if (det.blkcnt) cudaFree(det.blkcnt);
    if (det.counts) cudaFree(det.counts);
    if (ksplit_ws) cudaFree(ksplit_ws);
    if (ksplit_flags) cudaFree(ksplit_flags);
    if (d_u8) cudaFree(d_u8);
    for (Slot &sl : slots) {
        if (sl.d_u8) cudaFree(sl.d_u8);
        if (sl.det.rows) cudaFree(sl.det.rows);
        if (sl.det.mask) cudaFree(sl.det.mask);
        if (sl.det.blkcnt) cudaFree(sl.det.blkcnt);
        if (sl.det.counts) cudaFree(sl.det.counts);
        if (sl.h_rows) cudaFreeHost(sl.h_rows);
        if (sl.h_counts) cudaFreeHost(sl.h_counts);
        if (sl.ev_det) cudaEventDestroy(sl.ev_det);
        if (sl.d_in) cudaFree(sl.d_in);
        for (float *p : sl.d_out) if (p) cudaFree(p);
        for (float *p : sl.h_out) if (p) cudaFreeHost(p);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_comp) cudaEventDestroy(sl.ev_comp);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    if (s_det) cudaStreamDestroy(s_det);
    if (act_arena) cudaFree(act_arena);
    if (w_arena) cudaFree(w_arena);
    if (d_input) cudaFree(d_input);
    if (stream) cudaStreamDestroy(stream);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int grid_for(long total, int block = 256) {
    long g = (total + block - 1) / block;
    const long cap = 148L * 32;   // grid-stride kernels: a few waves of the 148 SMs
    return (int)std::max<long>(1, std::min(g, cap));
}

static TV make_tv(char *base, int N, int H, int W, int C, int ldc, int P, int dt, int coff) {
    TV t;
    t.base = base + (size_t)coff * dt_size(dt);
    t.N = N; t.H = H; t.W = W; t.C = C; t.ldc = ldc; t.P = P; t.Hp = H + 2 * P; t.Wp = W + 2 * P;
    return t;
}
static size_t tv_bytes(int N, int H, int W, int ldc, int P, int dt) {
    return (size_t)N * (H + 2 * P) * (W + 2 * P) * ldc * dt_size(dt);
}

// which layers read layer j's output
static std::vector<std::vector<int>> consumers_of(const Network &net) {
    const int nl = (int)net.layers.size();
    std::vector<std::vector<int>> cons(nl);
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net.layers[i];
        if (l.type == YB_ROUTE) {
            for (int s : l.input_layers) cons[s].push_back(i);
        } else if (l.type != YB_BLANK) {
            if (i > 0) cons[i - 1].push_back(i);
            if (l.type == YB_SHORTCUT) cons[l.index].push_back(i);
        }
    }
    return cons;
}

template <typename F>
static void dispatch_dt(int dt, F &&f) {
    if (dt == DT_F32) f((float *)nullptr);
    else f((__nv_bfloat16 *)nullptr);
}

std::shared_ptr<Engine> build_engine(Network *net, const EngineOptions &opt) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        fatal_throw("yolo2_light_b200: no CUDA device -- this library has no CPU fallback");
    CUDA_OK(cudaSetDevice(opt.device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, opt.device));
    if (prop.major != 10)
        fatal_throw(std::string("yolo2_light_b200: device '") + prop.name + "' is compute capability " +
                    std::to_string(prop.major) + "." + std::to_string(prop.minor) +
                    "; this build contains sm_100a code only");

    auto e = std::make_shared<Engine>();
    e->opt = opt;
    e->batch = net->batch;
    const int B = net->batch;
    const int nl = (int)net->layers.size();
    if (nl == 0) fatal_throw("empty network");
    CUDA_OK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));

    bool any_xnor = false;
    for (const Layer &l : net->layers) {
        if (l.type != YB_CONVOLUTIONAL) continue;
        if (l.batch_normalize) fatal_throw("engine: batch-norm not folded -- call yb_fuse_conv_batchnorm first");
        if (l.xnor) {
            any_xnor = true;
            if (!l.has_mean_arr) fatal_throw("engine: xnor layer without mean_arr -- call yb_calculate_binary_weights first");
        }
        if (opt.qrule && !l.has_int8)
            fatal_throw("engine: -quantized rule without int8 weights -- call yb_quantinization_and_get_multipliers first");
    }
    // f32 activations whenever an integer path must see exactly the reference's inputs
    const bool exact = opt.qrule || any_xnor || opt.precision == YB_PREC_FP32;
    e->act_dt = exact ? DT_F32 : DT_BF16;
    const int ADT = e->act_dt;

    auto conv_variant = [&](int i) -> int {   // 0 fp32, 1 xnor, 2 int8
        const Layer &l = net->layers[i];
        if (opt.qrule && (i + opt.q_index_offset) >= 1 && l.activation != YB_LINEAR) return 2;
        if (l.xnor) return 1;
        return 0;
    };

    const auto cons = consumers_of(*net);
    // XNOR layers with enough channels run on the tensor cores as +-1 int8 (dot == 2*count - K, exact); the small
    // ones stay on the popcount kernels.  YB_XNOR_TC=0 forces popcount everywhere.
    auto xnor_on_tc = [&](const Layer &l) {
        const char *ev = getenv("YB_XNOR_TC");
        if (ev && ev[0] == '0') return false;
        // channels are padded to a multiple of 32 with zero WEIGHT bytes (whatever the activation pad bytes hold contributes 0), so
        // even the 16- and 32-channel layers run here: the popcount kernels are bound by the 16 POPC/clk/SM of the integer pipe
        // (round 1: 79-85 % of that rate), the same layers as +-1 on kind::i8 take 0.55-0.75 of the time (profiles/r02_notes.md)
        const int minc = getenv("YB_XNOR_TC_MINC") ? atoi(getenv("YB_XNOR_TC_MINC")) : 16;
        return l.xnor && l.c % 16 == 0 && l.c >= minc && l.size == 3 && l.stride == 1 && l.pad == 1 && l.n >= 8;
    };

    // XNOR layers with stride != 1 or pad != 1 never reach the bit GEMM in the reference: forward_convolutional_layer_cpu
    // binarises the input to +-1 floats (binarize_cpu, additionally.c:128-134), swaps in the +-mean weights (binarize_weights,
    // :113-126) and runs the ordinary im2col + gemm_nn (yolov2_forward_network.c:40-50, :204) -- out-of-image taps count 0 there,
    // not -1.  Same here: k_binarize_pm1 + the exact-order float convolution.
    auto xnor_fallback = [&](const Layer &l) { return l.xnor && !(l.stride == 1 && l.pad == 1); };

    // integer conv i -> 2x2/2 max-pool i+1 -> integer conv i+2, nothing else reading i or i+1: the pool and the next layer's input
    // conversion can run in conv i's epilogue (tc_plan_fuse_pool).  Returns the mode (1 s8 quantised, 2 +-1 bytes) or 0.
    auto conv_pool_mode = [&](int i) -> int {
        if (!opt.fuse || opt.keep_counts || getenv("YB_NO_CONV_POOL_FUSE") || i + 2 >= nl) return 0;
        const Layer &mp = net->layers[i + 1], &c2 = net->layers[i + 2];
        if (mp.type != YB_MAXPOOL || mp.size != 2 || mp.stride != 2 || mp.pad != 1 || c2.type != YB_CONVOLUTIONAL) return 0;
        if (cons[i].size() != 1 || cons[i][0] != i + 1 || cons[i + 1].size() != 1 || cons[i + 1][0] != i + 2) return 0;
        const int v2 = conv_variant(i + 2);
        if (v2 == 2) return 1;
        if (v2 == 1 && xnor_on_tc(c2) && !xnor_fallback(c2)) return 2;
        if (v2 == 1 && !xnor_fallback(c2)) return 3;     // next XNOR layer reads sign bits (popcount kernels)
        return 0;
    };

    // ---- fusion plan: conv i + same-shape shortcut i+1 whose only reader is that shortcut -------------
    std::vector<int> fused_into(nl, -1);   // conv i writes layer fused_into[i]'s output
    std::vector<char> is_fused_sc(nl, 0);
    if (opt.fuse) {
        for (int i = 0; i + 1 < nl; ++i) {
            const Layer &l = net->layers[i], &s = net->layers[i + 1];
            if (l.type != YB_CONVOLUTIONAL || s.type != YB_SHORTCUT) continue;
            if (conv_variant(i) != 0) continue;
            if (cons[i].size() != 1 || cons[i][0] != i + 1) continue;
            if (s.index == i) continue;
            if (!(s.w == s.out_w && s.h == s.out_h && s.c == s.out_c)) continue;
            fused_into[i] = i + 1;
            is_fused_sc[i + 1] = 1;
        }
    }

    // ---- fusion plan: detection-head conv i + [yolo] i+1 (tensor-core path only; decided again when ops are emitted)
    std::vector<char> yolo_fused(nl, 0);

    // ---- output placement --------------------------------------------------------------------------
    // pass 1: decide dtype + home of every layer output (own buffer, or a channel slice of a concat buffer)
    struct Home { int owner = -1; int coff = 0; int ldc = 0; };   // owner: layer whose buffer holds it
    std::vector<Home> home(nl);
    e->out_dt.assign(nl, ADT);
    auto has_nhwc_out = [&](int i) {
        const Layer &l = net->layers[i];
        if (l.type == YB_YOLO || l.type == YB_REGION || l.type == YB_BLANK) return false;
        if (fused_into[i] >= 0) return false;
        if (l.out_h <= 0 || l.out_w <= 0 || l.out_c <= 0) return false;
        return true;
    };
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        if (l.type == YB_CONVOLUTIONAL && !cons[i].empty()) {
            bool all_final = true;
            for (int c : cons[i]) if (net->layers[c].type != YB_YOLO && net->layers[c].type != YB_REGION) all_final = false;
            if (all_final) e->out_dt[i] = DT_F32;   // detection heads stay f32 (bf16 would cost ~1e-3 rel by itself)
        }
        if (l.type == YB_CONVOLUTIONAL && cons[i].empty()) e->out_dt[i] = DT_F32;
    }
    if (opt.fuse) {
        for (int r = 0; r < nl; ++r) {
            const Layer &l = net->layers[r];
            if (l.type != YB_ROUTE || l.n < 2 || l.out_c <= 0) continue;
            int off = 0;
            for (int k = 0; k < l.n; ++k) {
                const int j = l.input_layers[k];
                const Layer &src = net->layers[j];
                const bool ok = has_nhwc_out(j) && home[j].owner < 0 && src.type != YB_ROUTE &&
                                e->out_dt[j] == ADT;
                if (ok) { home[j].owner = r; home[j].coff = off; home[j].ldc = l.out_c; }
                off += src.out_c;
            }
        }
    }
    // pass 2: sizes + arena offsets
    std::vector<size_t> buf_off(nl, (size_t)-1);
    size_t act_total = 0;
    const int P = 1;
    auto own_buffer = [&](int i) {
        const Layer &l = net->layers[i];
        buf_off[i] = act_total;
        act_total += align_up(tv_bytes(B, l.out_h, l.out_w, (int)align_up(l.out_c, 8), P, e->out_dt[i]), 1024);
    };
    const size_t in0_off = act_total;
    e->in0_dt = ADT;
    act_total += align_up(tv_bytes(B, net->h, net->w, net->c, P, ADT), 1024);
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        if (!has_nhwc_out(i)) continue;
        if (l.type == YB_ROUTE && l.n == 1 && opt.fuse) continue;   // pure alias
        if (home[i].owner >= 0) continue;                           // lives inside a concat buffer
        own_buffer(i);
    }
    // side buffers of the integer paths
    std::vector<size_t> side_off(nl, (size_t)-1);
    std::vector<int> side_ld(nl, 0);
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        if (l.type != YB_CONVOLUTIONAL) continue;
        const int v = conv_variant(i);
        if (v == 1 && xnor_fallback(l)) {
            side_ld[i] = l.c;   // +-1 floats
            side_off[i] = act_total;
            act_total += align_up(tv_bytes(B, l.h, l.w, side_ld[i], P, DT_F32), 1024);
        } else if (v == 1 && xnor_on_tc(l)) {
            side_ld[i] = (int)align_up(l.c, 32);   // +-1 bytes; pad channels meet zero weights
            side_off[i] = act_total;
            act_total += align_up(tv_bytes(B, l.h, l.w, side_ld[i], P, DT_S8), 1024);
        } else if (v == 1) {
            side_ld[i] = (l.c + 31) / 32;
            side_off[i] = act_total;
            act_total += align_up(tv_bytes(B, l.h, l.w, side_ld[i], P, DT_BITS), 1024);
        } else if (v == 2) {
            side_ld[i] = (int)align_up(l.c, 32);   // zero-padded channels: every INT8 layer fits the kind::i8 tensor-core tile
            side_off[i] = act_total;
            act_total += align_up(tv_bytes(B, l.h, l.w, side_ld[i], P, DT_S8), 1024);
        }
    }
    e->act_bytes = act_total;
    CUDA_OK(cudaMalloc(&e->act_arena, act_total));
    CUDA_OK(cudaMemsetAsync(e->act_arena, 0, act_total, e->stream));   // zero borders, once

    for (int i = 0; i < nl; ++i) {   // +-1 activation buffers: borders are -1 (out-of-image taps count as -1)
        const Layer &l = net->layers[i];
        if (l.type == YB_CONVOLUTIONAL && conv_variant(i) == 1 && xnor_on_tc(l) && !xnor_fallback(l))
            CUDA_OK(cudaMemsetAsync(e->act_arena + side_off[i], 0xFF, tv_bytes(B, l.h, l.w, side_ld[i], P, DT_S8), e->stream));
    }
    e->in0 = make_tv(e->act_arena + in0_off, B, net->h, net->w, net->c, net->c, P, ADT, 0);
    e->out_tv.assign(nl, TV{});
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        if (!has_nhwc_out(i)) continue;
        if (l.type == YB_ROUTE && l.n == 1 && opt.fuse) {
            e->out_tv[i] = e->out_tv[l.input_layers[0]];
            e->out_dt[i] = e->out_dt[l.input_layers[0]];
            continue;
        }
        if (home[i].owner >= 0) continue;
        // pixel stride rounded up to 8 channels: 16-byte aligned rows for TMA / vector stores (e.g. 255 -> 256)
        e->out_tv[i] = make_tv(e->act_arena + buf_off[i], B, l.out_h, l.out_w, l.out_c, (int)align_up(l.out_c, 8), P, e->out_dt[i], 0);
    }
    // slices (owner buffers are allocated above since routes own their buffers)
    for (int i = 0; i < nl; ++i) {
        if (home[i].owner < 0) continue;
        const Layer &l = net->layers[i];
        const int r = home[i].owner;
        e->out_tv[i] = make_tv(e->act_arena + buf_off[r], B, l.out_h, l.out_w, l.out_c, e->out_tv[r].ldc, P, ADT, home[i].coff);
    }
    // single-input route aliases may point at slices that were only resolved now
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        if (has_nhwc_out(i) && l.type == YB_ROUTE && l.n == 1 && opt.fuse) {
            e->out_tv[i] = e->out_tv[l.input_layers[0]];
            e->out_dt[i] = e->out_dt[l.input_layers[0]];
        }
    }

    // ---- final (host-visible) outputs ----------------------------------------------------------------
    e->d_final.assign(nl, nullptr);
    e->h_final.assign(nl, nullptr);
    e->final_count.assign(nl, 0);
    e->d_counts.assign(nl, nullptr);
    e->counts_count.assign(nl, 0);
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        if (l.type == YB_YOLO || l.type == YB_REGION) {
            e->final_count[i] = (size_t)l.outputs * B;
            CUDA_OK(cudaMalloc(&e->d_final[i], e->final_count[i] * sizeof(float)));
            CUDA_OK(cudaHostAlloc(&e->h_final[i], e->final_count[i] * sizeof(float), cudaHostAllocDefault));
        }
    }
    {   // the reference returns the LAST layer's output; keep a host copy for it whatever its type
        const int last = nl - 1;
        const Layer &l = net->layers[last];
        if (!e->d_final[last] && l.outputs > 0) {
            e->final_count[last] = (size_t)l.outputs * B;
            CUDA_OK(cudaMalloc(&e->d_final[last], e->final_count[last] * sizeof(float)));
            CUDA_OK(cudaHostAlloc(&e->h_final[last], e->final_count[last] * sizeof(float), cudaHostAllocDefault));
        }
    }

    // ---- weight arena ------------------------------------------------------------------------------
    std::vector<ConvWeights> cw(nl);
    std::vector<char> hostw;
    auto reserve = [&](size_t bytes) { size_t off = align_up(hostw.size(), 1024); hostw.resize(off + bytes, 0); return off; };
    std::vector<int> use_tc(nl, 0);
    e->not_materialised.assign(nl, 0);
    std::vector<char> prefilled(nl, 0);   // the producing max-pool already wrote this conv's s8 / sign input (fused)
    std::vector<char> pool_in_conv(nl, 0); // this max-pool runs inside the epilogue of the integer convolution in front of it
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        if (l.type != YB_CONVOLUTIONAL) continue;
        const int v = conv_variant(i);
        const int taps = l.size * l.size;
        const int K = taps * l.c;
        ConvWeights &w = cw[i];
        w.bias = reserve(sizeof(float) * align_up(l.n, 64));
        memcpy(&hostw[w.bias], l.biases.data(), sizeof(float) * l.n);
        if (v == 0) {
            const TV &tin = (i == 0) ? e->in0 : e->out_tv[i - 1];
            const int in_dt = (i == 0) ? e->in0_dt : e->out_dt[i - 1];
            const int odt = fused_into[i] >= 0 ? e->out_dt[fused_into[i]] : e->out_dt[i];
            const TV &tout = e->out_tv[fused_into[i] >= 0 ? fused_into[i] : i];
            use_tc[i] = (ADT == DT_BF16 && in_dt == DT_BF16) ? tc_conv_supported(l, tin, tout, odt == DT_BF16) : 0;
            // float detection heads of the INT8 / XNOR networks (default precision): kind::tf32.  Only layers whose every
            // reader is a yolo / region layer -- nothing they compute can reach an integer layer.
            if (ADT == DT_F32 && opt.precision == YB_PREC_BF16_TC && in_dt == DT_F32 && odt == DT_F32 && fused_into[i] < 0 &&
                !cons[i].empty() && !getenv("YB_NO_TF32")) {
                bool heads_only = true;
                for (int r : cons[i]) heads_only &= net->layers[r].type == YB_YOLO || net->layers[r].type == YB_REGION;
                if (heads_only && tc_tf32_supported(l, tin, tout)) use_tc[i] = 2;
            }
            if (getenv("YB_NO_TC")) use_tc[i] = 0;
            if (use_tc[i] == 2) {
                w.ldn = (int)align_up(l.n, 64);
                w.w_f32km = reserve(sizeof(float) * (size_t)w.ldn * K);
                float *dst = reinterpret_cast<float *>(&hostw[w.w_f32km]);
                for (int f = 0; f < l.n; ++f)
                    for (int c = 0; c < l.c; ++c)
                        for (int t = 0; t < taps; ++t)
                            dst[(size_t)f * K + (size_t)t * l.c + c] = l.weights[((size_t)f * l.c + c) * taps + t];
            } else if (use_tc[i]) {
                // bf16 [ldn][K], K ordered (ky, kx, c): the K-major B operand of the implicit GEMM
                w.ldn = (int)align_up(l.n, 64);
                w.w_bf16 = reserve(sizeof(__nv_bfloat16) * (size_t)w.ldn * K);
                __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(&hostw[w.w_bf16]);
                for (int f = 0; f < l.n; ++f)
                    for (int c = 0; c < l.c; ++c)
                        for (int t = 0; t < taps; ++t)
                            dst[(size_t)f * K + (size_t)t * l.c + c] =
                                __float2bfloat16_rn(l.weights[((size_t)f * l.c + c) * taps + t]);
            } else {
                // f32 [K][ldw]; K ordered (ky, kx, c), or -- f32 activations: the exact order of the reference's gemm_nn
                // (k_conv_simt<EXACT>) -- (c, ky, kx)
                w.ldw = (int)align_up(l.n, 64);
                w.w_f32 = reserve(sizeof(float) * (size_t)K * w.ldw);
                float *dst = reinterpret_cast<float *>(&hostw[w.w_f32]);
                for (int f = 0; f < l.n; ++f)
                    for (int c = 0; c < l.c; ++c)
                        for (int t = 0; t < taps; ++t)
                            dst[(ADT == DT_F32 ? (size_t)c * taps + t : (size_t)t * l.c + c) * w.ldw + f] =
                                l.weights[((size_t)f * l.c + c) * taps + t];
            }
        } else if (v == 1 && xnor_fallback(l)) {
            // f32 [K][ldw], K in the reference's (c, ky, kx) order: +mean where w > 0, -mean otherwise (binarize_weights)
            w.ldw = (int)align_up(l.n, 64);
            w.w_f32 = reserve(sizeof(float) * (size_t)K * w.ldw);
            float *dst = reinterpret_cast<float *>(&hostw[w.w_f32]);
            for (int f = 0; f < l.n; ++f)
                for (int c = 0; c < l.c; ++c)
                    for (int t = 0; t < taps; ++t)
                        dst[((size_t)c * taps + t) * w.ldw + f] = l.weights[((size_t)f * l.c + c) * taps + t] > 0 ? l.mean_arr[f] : -l.mean_arr[f];
        } else if (v == 1 && xnor_on_tc(l)) {
            // +-1 bytes [ldn][taps][cpad]: +1 where w > 0, -1 otherwise; padded filter rows and padded channels stay 0
            w.cpad = side_ld[i];
            w.ldn = (int)align_up(l.n, 64);
            w.w_s8 = reserve((size_t)w.ldn * taps * w.cpad);
            int8_t *dst = reinterpret_cast<int8_t *>(&hostw[w.w_s8]);
            for (int f = 0; f < l.n; ++f)
                for (int c = 0; c < l.c; ++c)
                    for (int t = 0; t < taps; ++t)
                        dst[((size_t)f * taps + t) * w.cpad + c] = l.weights[((size_t)f * l.c + c) * taps + t] > 0 ? 1 : -1;
            w.mean = reserve(sizeof(float) * align_up(l.n, 64));
            memcpy(&hostw[w.mean], l.mean_arr.data(), sizeof(float) * l.n);
        } else if (v == 1) {
            // sign bits [ldn][taps][CW]; bit = (w > 0) (binarize_weights additionally.c:113 + float_to_bit :1536)
            const int CW = (l.c + 31) / 32;
            w.ldn = (int)align_up(l.n, 64);
            w.cpad = CW * 32;
            w.w_bits = reserve(sizeof(uint32_t) * (size_t)w.ldn * taps * CW);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&hostw[w.w_bits]);
            for (int f = 0; f < l.n; ++f)
                for (int c = 0; c < l.c; ++c)
                    for (int t = 0; t < taps; ++t)
                        if (l.weights[((size_t)f * l.c + c) * taps + t] > 0)
                            dst[((size_t)f * taps + t) * CW + c / 32] |= 1u << (c & 31);
            w.mean = reserve(sizeof(float) * align_up(l.n, 64));
            memcpy(&hostw[w.mean], l.mean_arr.data(), sizeof(float) * l.n);
        } else {
            // s8 [ldn][taps][cpad], zero channel padding
            w.cpad = side_ld[i];
            w.ldn = (int)align_up(l.n, 64);
            w.w_s8 = reserve((size_t)w.ldn * taps * w.cpad);
            int8_t *dst = reinterpret_cast<int8_t *>(&hostw[w.w_s8]);
            for (int f = 0; f < l.n; ++f)
                for (int c = 0; c < l.c; ++c)
                    for (int t = 0; t < taps; ++t)
                        dst[((size_t)f * taps + t) * w.cpad + c] = l.weights_int8[((size_t)f * l.c + c) * taps + t];
        }
    }
    size_t stem_w_off = (size_t)-1;
    {
        const Layer &l0 = net->layers[0];
        if (ADT == DT_BF16 && l0.type == YB_CONVOLUTIONAL && conv_variant(0) == 0 && l0.c == 3 && l0.size == 3 && l0.n <= 32) {
            stem_w_off = reserve(sizeof(__nv_bfloat16) * 32 * 32);
            __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(&hostw[stem_w_off]);
            for (int i = 0; i < 32 * 32; ++i) dst[i] = __float2bfloat16_rn(0.f);
            for (int f = 0; f < l0.n; ++f)
                for (int c = 0; c < 3; ++c)
                    for (int t = 0; t < 9; ++t)
                        dst[f * 32 + t * 3 + c] = __float2bfloat16_rn(l0.weights[((size_t)f * 3 + c) * 9 + t]);
        }
    }
    e->w_bytes = align_up(std::max<size_t>(hostw.size(), 1024), 1024);
    CUDA_OK(cudaMalloc(&e->w_arena, e->w_bytes));
    if (opt.upload) CUDA_OK(cudaMemcpyAsync(e->w_arena, hostw.data(), hostw.size(), cudaMemcpyHostToDevice, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));   // hostw goes out of scope below

    // ---- input staging -----------------------------------------------------------------------------
    e->input_count = (size_t)B * net->c * net->h * net->w;
    CUDA_OK(cudaMalloc(&e->d_input, e->input_count * sizeof(float)));

    // ---- op list -----------------------------------------------------------------------------------
    Engine *E = e.get();
    bool stem_fused = false, stem_pool_fused = false;
    {
        // ops[0] consumes the caller's NCHW f32 images.  Usually that is the stem convolution itself (3 input
        // channels, 3x3/1/1), reading NCHW directly; otherwise a plain NCHW -> padded-NHWC conversion.
        const Layer &l0 = net->layers[0];
        const bool stem_ok = l0.type == YB_CONVOLUTIONAL && conv_variant(0) == 0 && !use_tc[0] && l0.c == 3 &&
                             l0.size == 3 && l0.stride == 1 && l0.pad == 1 && (l0.n == 16 || l0.n == 32) &&
                             fused_into[0] < 0 && e->out_tv[0].base && !getenv("YB_NO_STEM") &&
                             (e->out_dt[0] == DT_F32 || (e->out_tv[0].ldc % 8 == 0));
        // exact nets: stem + 2x2/2 max-pool + the integer layer's input conversion in one kernel (k_stem_pool): layers 0 and 1
        // are then never written to HBM
        bool pool_ok = stem_ok && opt.fuse && e->out_dt[0] == DT_F32 && l0.n == 16 && nl > 2 && !getenv("YB_NO_STEM_POOL_FUSE") &&
                       (l0.activation == YB_LEAKY || l0.activation == YB_LINEAR);
        if (pool_ok) {
            const Layer &mp = net->layers[1], &c2 = net->layers[2];
            pool_ok = mp.type == YB_MAXPOOL && mp.size == 2 && mp.stride == 2 && mp.pad == 1 && cons[0].size() == 1 && cons[0][0] == 1 &&
                      cons[1].size() == 1 && cons[1][0] == 2 && c2.type == YB_CONVOLUTIONAL && conv_variant(2) != 0 &&
                      side_off[2] != (size_t)-1 && !xnor_fallback(c2) && (conv_variant(2) == 1 && !xnor_on_tc(c2) ? true : side_ld[2] % 16 == 0);
        }
        if (pool_ok) {
            stem_fused = true; stem_pool_fused = true;
            const Layer &c2 = net->layers[2];
            const int v2 = conv_variant(2);
            const bool pm1 = v2 == 1 && xnor_on_tc(c2);      // next layer reads +-1 bytes
            const TV q = (v2 == 2 || pm1) ? make_tv(e->act_arena + side_off[2], B, c2.h, c2.w, c2.c, side_ld[2], P, DT_S8, 0)
                                          : make_tv(e->act_arena + side_off[2], B, c2.h, c2.w, side_ld[2], side_ld[2], P, DT_BITS, 0);
            StemW<16> w16{};
            for (int f = 0; f < 16; ++f) {
                for (int c = 0; c < 3; ++c)
                    for (int t = 0; t < 9; ++t) w16.w[(t * 3 + c) * 16 + f] = l0.weights[((size_t)f * 3 + c) * 9 + t];
                w16.b[f] = l0.biases[f];
            }
            const int act = l0.activation, H = l0.h, W = l0.w;
            const float mult = (v2 == 2) ? c2.input_quant_multipler : 0.f;
            const int grid = (int)(((long)B * c2.h * c2.w + 127) / 128);
            prefilled[2] = 1;
            e->not_materialised[0] = e->not_materialised[1] = 1;
            e->first_kind = OP_CONV_SIMT; e->first_layer = 0;
            e->first_op = [=](const float *din, cudaStream_t s) {
                const unsigned long long nz = 0x8000000080000000ull;
                if (v2 == 2 && act == ACT_LEAKY) k_stem_pool<0, ACT_LEAKY><<<grid, 128, 0, s>>>(din, q, w16, act, H, W, mult, nz);
                else if (v2 == 2) k_stem_pool<0, ACT_LINEAR><<<grid, 128, 0, s>>>(din, q, w16, act, H, W, mult, nz);
                else if (pm1 && act == ACT_LEAKY) k_stem_pool<1, ACT_LEAKY><<<grid, 128, 0, s>>>(din, q, w16, act, H, W, mult, nz);
                else if (pm1) k_stem_pool<1, ACT_LINEAR><<<grid, 128, 0, s>>>(din, q, w16, act, H, W, mult, nz);
                else if (act == ACT_LEAKY) k_stem_pool<2, ACT_LEAKY><<<grid, 128, 0, s>>>(din, q, w16, act, H, W, mult, nz);
                else k_stem_pool<2, ACT_LINEAR><<<grid, 128, 0, s>>>(din, q, w16, act, H, W, mult, nz);
            };
        } else if (stem_ok && stem_w_off != (size_t)-1 && e->out_dt[0] == DT_BF16 && tc_stem_supported(l0, e->out_tv[0]) &&
            !getenv("YB_NO_STEM_TC")) {
            // tensor-core stem: gathers the 3x3x3 window from NCHW, K padded 27 -> 32
            stem_fused = true;
            void *sp = tc_stem_make_plan(l0, e->out_tv[0], e->w_arena + stem_w_off,
                                         reinterpret_cast<const float *>(e->w_arena + cw[0].bias));
            e->stem_plan = sp;
            e->first_kind = OP_CONV_TC; e->first_layer = 0;
            e->first_op = [sp](const float *din, cudaStream_t s) { tc_stem_launch(sp, din, s); };
            if (!getenv("YB_NO_STEM_U8")) e->first_op_u8 = [sp](const unsigned char *d8, cudaStream_t s) { tc_stem_launch_u8(sp, d8, s); };
        } else if (stem_ok) {
            stem_fused = true;
            const TV tout = e->out_tv[0];
            const int act = l0.activation, H = l0.h, W = l0.w, nf = l0.n, odt = e->out_dt[0];
            // weights go to the kernel as by-value constants: [27 = (ky,kx,c)][n] + bias
            StemW<32> w32{}; StemW<16> w16{};
            for (int f = 0; f < nf; ++f) {
                for (int c = 0; c < 3; ++c)
                    for (int t = 0; t < 9; ++t) {
                        const float v = l0.weights[((size_t)f * 3 + c) * 9 + t];
                        if (nf == 32) w32.w[(t * 3 + c) * 32 + f] = v; else w16.w[(t * 3 + c) * 16 + f] = v;
                    }
                if (nf == 32) w32.b[f] = l0.biases[f]; else w16.b[f] = l0.biases[f];
            }
            const long total = (long)B * H * W;
            const int grid = (int)((total + 127) / 128);
            e->first_kind = OP_CONV_SIMT; e->first_layer = 0;
            e->first_op = [=](const float *din, cudaStream_t s) {
                if (nf == 32 && odt == DT_BF16) k_conv_stem<32, __nv_bfloat16><<<grid, 128, 0, s>>>(din, tout, w32, act, H, W);
                else if (nf == 32) k_conv_stem<32, float, true><<<grid, 128, 0, s>>>(din, tout, w32, act, H, W);
                else if (odt == DT_BF16) k_conv_stem<16, __nv_bfloat16><<<grid, 128, 0, s>>>(din, tout, w16, act, H, W);
                else k_conv_stem<16, float, true><<<grid, 128, 0, s>>>(din, tout, w16, act, H, W);
            };
        } else {
            const TV in0 = e->in0;
            const int dt = e->in0_dt;
            const int g = grid_for((long)in0.N * in0.H * in0.W);
            e->first_op = [=](const float *din, cudaStream_t s) {
                if (dt == DT_F32) k_input_nchw_to_nhwc<float><<<g, 256, 0, s>>>(din, in0);
                else k_input_nchw_to_nhwc<__nv_bfloat16><<<g, 256, 0, s>>>(din, in0);
            };
        }
        e->ops.push_back(Op{e->first_kind, e->first_layer, nullptr});   // launched specially: pointer varies per call
    }
    for (int i = 0; i < nl; ++i) {
        const Layer &l = net->layers[i];
        const TV tin = (i == 0) ? e->in0 : e->out_tv[i - 1];
        const int in_dt = (i == 0) ? e->in0_dt : e->out_dt[i - 1];
        const bool prev_ok = (i == 0) || e->out_tv[i - 1].base != nullptr;
        auto need_prev = [&]() {
            if (!prev_ok) fatal_throw("engine: layer " + std::to_string(i) + " has no image input");
        };
        if (i == 0 && stem_fused) continue;
        if (i == 1 && stem_pool_fused) continue;
        switch (l.type) {
        case YB_CONVOLUTIONAL: {
            need_prev();
            const int v = conv_variant(i);
            const int tgt = fused_into[i] >= 0 ? fused_into[i] : i;
            const TV tout = e->out_tv[tgt];
            const int odt = e->out_dt[tgt];
            if (!tout.base) fatal_throw("engine: conv output not placed");
            if (tin.C != l.c || tin.H != l.h || tin.W != l.w) fatal_throw("engine: conv input shape mismatch");
            const long M = (long)B * l.out_h * l.out_w;
            if (v == 0) {
                TV res{}; int rdt = DT_F32; int act2 = ACT_LINEAR;
                if (fused_into[i] >= 0) {
                    const Layer &s = net->layers[tgt];
                    res = e->out_tv[s.index];
                    rdt = e->out_dt[s.index];
                    act2 = s.activation;
                    if (!res.base) fatal_throw("engine: shortcut source not placed");
                }
                if (use_tc[i] == 2) {
                    const bool fuse_yolo = opt.fuse && i + 1 < nl && net->layers[i + 1].type == YB_YOLO && cons[i].size() == 1 &&
                                           cons[i][0] == i + 1 && e->d_final[i + 1] && !getenv("YB_NO_YOLO_FUSE");
                    void *plan = tc_make_plan_tf32(l, tin, tout, e->w_arena + cw[i].w_f32km, cw[i].ldn,
                                                   reinterpret_cast<const float *>(e->w_arena + cw[i].bias), fuse_yolo ? 1 : 0);
                    e->tc_plans.push_back(plan);
                    ++e->n_tc;
                    if (fuse_yolo) {
                        tc_plan_fuse_yolo(plan, e->d_final[i + 1], net->layers[i + 1].classes);
                        yolo_fused[i + 1] = 1;
                    }
                    e->ops.push_back(Op{OP_CONV_TC_TF32, i, [plan](cudaStream_t s) { tc_launch(plan, s); }});
                } else if (use_tc[i]) {
                    const bool fuse_yolo = opt.fuse && fused_into[i] < 0 && odt == DT_F32 && i + 1 < nl &&
                                           net->layers[i + 1].type == YB_YOLO && cons[i].size() == 1 && cons[i][0] == i + 1 &&
                                           e->d_final[i + 1] && !getenv("YB_NO_YOLO_FUSE");
                    const char *ks_env = getenv("YB_TC_KSPLIT");   // 1: on without the option (A/B runs)
                    const bool want_ksplit = opt.ksplit || (ks_env && ks_env[0] == '1');
                    void *plan = tc_make_plan(l, tin, tout, odt == DT_BF16, res, rdt == DT_BF16, act2,
                                              e->w_arena + cw[i].w_bf16, cw[i].ldn,
                                              reinterpret_cast<const float *>(e->w_arena + cw[i].bias), fuse_yolo ? 1 : 0,
                                              want_ksplit ? 1 : 0);
                    e->tc_plans.push_back(plan);
                    if (!e->ksplit_ws) {
                        int sms = 148;
                        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, opt.device);
                        CUDA_OK(cudaMalloc(&e->ksplit_ws, tc_ksplit_ws_bytes(sms)));
                        CUDA_OK(cudaMalloc(&e->ksplit_flags, tc_ksplit_flag_bytes(sms)));
                        CUDA_OK(cudaMemset(e->ksplit_flags, 0, tc_ksplit_flag_bytes(sms)));
                    }
                    ++e->n_tc;
                    if (want_ksplit)
                        e->n_ksplit += tc_plan_enable_ksplit(plan, e->ksplit_ws, e->ksplit_flags);
                    if (fuse_yolo) {
                        tc_plan_fuse_yolo(plan, e->d_final[i + 1], net->layers[i + 1].classes);
                        yolo_fused[i + 1] = 1;
                    }
                    e->ops.push_back(Op{tc_plan_cta_group(plan) == 2 ? OP_CONV_TC2 : OP_CONV_TC, i, [plan](cudaStream_t s) { tc_launch(plan, s); }});
                } else {
                    ConvP p{};
                    p.in = tin; p.out = tout; p.res = res;
                    p.w = e->w_arena + cw[i].w_f32;
                    p.bias = reinterpret_cast<const float *>(e->w_arena + cw[i].bias);
                    p.n = l.n; p.ldw = cw[i].ldw; p.size = l.size; p.stride = l.stride; p.pad = l.pad;
                    p.act = l.activation; p.act2 = act2; p.K = l.size * l.size * l.c; p.M = M;
                    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((l.n + 63) / 64));
                    const int key = in_dt * 4 + odt * 2 + rdt;
                    e->ops.push_back(Op{OP_CONV_SIMT, i, [p, grid, key](cudaStream_t s) {
                        switch (key) {
                        case 0: k_conv_simt<float, float, float, true><<<grid, 256, 0, s>>>(p); break;   // reference order, bit-exact
                        case 1: k_conv_simt<float, float, __nv_bfloat16><<<grid, 256, 0, s>>>(p); break;
                        case 2: k_conv_simt<float, __nv_bfloat16, float><<<grid, 256, 0, s>>>(p); break;
                        case 3: k_conv_simt<float, __nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(p); break;
                        case 4: k_conv_simt<__nv_bfloat16, float, float><<<grid, 256, 0, s>>>(p); break;
                        case 5: k_conv_simt<__nv_bfloat16, float, __nv_bfloat16><<<grid, 256, 0, s>>>(p); break;
                        case 6: k_conv_simt<__nv_bfloat16, __nv_bfloat16, float><<<grid, 256, 0, s>>>(p); break;
                        default: k_conv_simt<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(p); break;
                        }
                    }});
                }
            } else if (v == 1) {
                if (in_dt != DT_F32 || odt != DT_F32) fatal_throw("engine: xnor path needs f32 activations");
                if (xnor_fallback(l)) {
                    TV pm1 = make_tv(e->act_arena + side_off[i], B, l.h, l.w, l.c, l.c, P, DT_F32, 0);
                    const int gb = grid_for((long)B * l.h * l.w * l.c);
                    e->ops.push_back(Op{OP_BINARIZE, i, [tin, pm1, gb](cudaStream_t s) { k_binarize_pm1<<<gb, 256, 0, s>>>(tin, pm1); }});
                    ConvP p{};
                    p.in = pm1; p.out = tout; p.res = TV{};
                    p.w = e->w_arena + cw[i].w_f32;
                    p.bias = reinterpret_cast<const float *>(e->w_arena + cw[i].bias);
                    p.n = l.n; p.ldw = cw[i].ldw; p.size = l.size; p.stride = l.stride; p.pad = l.pad;
                    p.act = l.activation; p.act2 = ACT_LINEAR; p.K = l.size * l.size * l.c; p.M = M;
                    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((l.n + 63) / 64));
                    e->ops.push_back(Op{OP_CONV_SIMT, i, [p, grid](cudaStream_t s) { k_conv_simt<float, float, float, true><<<grid, 256, 0, s>>>(p); }});
                    break;
                }
                int32_t *cnt_dbg = nullptr;
                if (opt.keep_counts) {
                    e->counts_count[i] = (size_t)B * l.n * l.out_h * l.out_w;
                    CUDA_OK(cudaMalloc(&e->d_counts[i], e->counts_count[i] * sizeof(int32_t)));
                    cnt_dbg = e->d_counts[i];
                }
                const bool in_vec = (tin.ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(tin.base) & 15) == 0;
                if (xnor_on_tc(l) && in_vec) {
                    TV q = make_tv(e->act_arena + side_off[i], B, l.h, l.w, l.c, side_ld[i], P, DT_S8, 0);
                    if (tc_i8_supported(l, q, tout)) {
                        const int g = grid_for((long)B * l.h * l.w * (l.c / 16));
                        if (!prefilled[i])
                            e->ops.push_back(Op{OP_BINARIZE, i, [tin, q, g](cudaStream_t s) { k_binarize_s8<<<g, 256, 0, s>>>(tin, q); }});
                        void *plan = tc_make_plan_xnor(l, q, tout, e->w_arena + cw[i].w_s8, cw[i].ldn,
                                                       reinterpret_cast<const float *>(e->w_arena + cw[i].bias),
                                                       reinterpret_cast<const float *>(e->w_arena + cw[i].mean), cnt_dbg,
                                                       conv_pool_mode(i) != 0);
                        e->tc_plans.push_back(plan);
                        if (const int pm = conv_pool_mode(i)) {
                            const Layer &c2 = net->layers[i + 2];
                            TV qn = make_tv(e->act_arena + side_off[i + 2], B, c2.h, c2.w, c2.c, side_ld[i + 2], P, DT_S8, 0);
                            if (tc_plan_fuse_pool(plan, pm, pm == 1 ? c2.input_quant_multipler : 0.f, qn)) {
                                prefilled[i + 2] = 1; pool_in_conv[i + 1] = 1;
                                e->not_materialised[i] = e->not_materialised[i + 1] = 1;
                            }
                        }
                        e->ops.push_back(Op{OP_CONV_TC_I8, i, [plan](cudaStream_t s) { tc_launch(plan, s); }});
                        break;
                    }
                    fatal_throw("engine: xnor tensor-core layer not supported by the i8 tile");
                }
                const int CW = side_ld[i];
                TV bits = make_tv(e->act_arena + side_off[i], B, l.h, l.w, CW, CW, P, DT_BITS, 0);
                {
                    if (in_vec) {
                        const int g = grid_for((long)B * l.h * l.w * CW);
                        if (!prefilled[i])
                            e->ops.push_back(Op{OP_BINARIZE, i, [tin, bits, g](cudaStream_t s) { k_binarize_vec<float><<<g, 256, 0, s>>>(tin, bits); }});
                    } else {
                        const long total = (long)B * l.h * l.w * CW * 32;
                        const int g = grid_for(total);
                        if (!prefilled[i])
                            e->ops.push_back(Op{OP_BINARIZE, i, [tin, bits, g](cudaStream_t s) { k_binarize<float><<<g, 256, 0, s>>>(tin, bits); }});
                    }
                }
                XnorP p{};
                p.bits = bits; p.out = tout;
                p.w = reinterpret_cast<const uint32_t *>(e->w_arena + cw[i].w_bits);
                p.mean = reinterpret_cast<const float *>(e->w_arena + cw[i].mean);
                p.bias = reinterpret_cast<const float *>(e->w_arena + cw[i].bias);
                p.n = l.n; p.size = l.size; p.pad = l.pad; p.K = l.size * l.size * l.c;
                p.padbits = (CW * 32 - l.c) * l.size * l.size;
                p.act = l.activation; p.M = M; p.counts = cnt_dbg;
                if (CW <= 2 && l.size == 3 && (size_t)l.n * 9 * CW * 4 <= 40 * 1024 && tout.ldc % 4 == 0) {
                    // small K: one thread per pixel, all filters (weights broadcast from shared memory)
                    const unsigned gsm = (unsigned)((M + 127) / 128);
                    const size_t smem = (size_t)l.n * 9 * CW * 4;
                    const int cw1 = CW;
                    const int pm = (cnt_dbg == nullptr) ? conv_pool_mode(i) : 0;
                    if (pm == 2 || pm == 3) {
                        // the 2x2 max-pool behind this layer and the next XNOR layer's sign extraction run in this kernel
                        const Layer &c2 = net->layers[i + 2];
                        const TV qn = (pm == 2) ? make_tv(e->act_arena + side_off[i + 2], B, c2.h, c2.w, c2.c, side_ld[i + 2], P, DT_S8, 0)
                                                : make_tv(e->act_arena + side_off[i + 2], B, c2.h, c2.w, side_ld[i + 2], side_ld[i + 2], P, DT_BITS, 0);
                        const unsigned gp = (unsigned)(((long)B * c2.h * c2.w + 127) / 128);
                        prefilled[i + 2] = 1; pool_in_conv[i + 1] = 1;
                        e->not_materialised[i] = e->not_materialised[i + 1] = 1;
                        e->ops.push_back(Op{OP_CONV_XNOR, i, [p, qn, gp, smem, cw1, pm](cudaStream_t s) {
                            if (cw1 == 1 && pm == 2) k_conv_xnor_smallk_pool<1, 2><<<gp, 128, smem, s>>>(p, qn);
                            else if (cw1 == 1) k_conv_xnor_smallk_pool<1, 3><<<gp, 128, smem, s>>>(p, qn);
                            else if (pm == 2) k_conv_xnor_smallk_pool<2, 2><<<gp, 128, smem, s>>>(p, qn);
                            else k_conv_xnor_smallk_pool<2, 3><<<gp, 128, smem, s>>>(p, qn);
                        }});
                        break;
                    }
                    e->ops.push_back(Op{OP_CONV_XNOR, i, [p, gsm, smem, cw1](cudaStream_t s) {
                        if (cw1 == 1) k_conv_xnor_smallk<1><<<gsm, 128, smem, s>>>(p);
                        else k_conv_xnor_smallk<2><<<gsm, 128, smem, s>>>(p);
                    }});
                    break;
                }
                dim3 grid((unsigned)((M + 63) / 64), (unsigned)((l.n + 63) / 64));
                e->ops.push_back(Op{OP_CONV_XNOR, i, [p, grid](cudaStream_t s) { k_conv_xnor<<<grid, 256, 0, s>>>(p); }});
            } else {
                if (in_dt != DT_F32 || odt != DT_F32) fatal_throw("engine: int8 path needs f32 activations");
                const int cpad = side_ld[i];
                TV q = make_tv(e->act_arena + side_off[i], B, l.h, l.w, l.c, cpad, P, DT_S8, 0);
                const float mult = l.input_quant_multipler;
                {
                    const long total = (long)B * l.h * l.w * (cpad / 4);
                    const int g = grid_for(total);
                    if (!prefilled[i])
                        e->ops.push_back(Op{OP_QUANTIZE, i, [tin, q, mult, g](cudaStream_t s) { k_quantize<float><<<g, 256, 0, s>>>(tin, q, mult); }});
                }
                int *acc_dbg = nullptr;
                if (opt.keep_counts) {
                    e->counts_count[i] = (size_t)B * l.n * l.out_h * l.out_w;
                    CUDA_OK(cudaMalloc(&e->d_counts[i], e->counts_count[i] * sizeof(int32_t)));
                    acc_dbg = e->d_counts[i];
                }
                const float alpha1 = 32 / (l.input_quant_multipler * l.weights_quant_multipler);   // ALPHA1, ..._quantized.c:598
                if (!getenv("YB_NO_TC") && tc_i8_supported(l, q, tout)) {
                    // s8 x s8 -> s32 on tcgen05 (kind::i8); weights [ldn][taps][cpad] are already K-major
                    void *plan = tc_make_plan_i8(l, q, tout, e->w_arena + cw[i].w_s8, cw[i].ldn,
                                                 reinterpret_cast<const float *>(e->w_arena + cw[i].bias), alpha1, acc_dbg,
                                                 conv_pool_mode(i) != 0);
                    e->tc_plans.push_back(plan);
                    if (const int pm = conv_pool_mode(i)) {
                        const Layer &c2 = net->layers[i + 2];
                        TV qn = make_tv(e->act_arena + side_off[i + 2], B, c2.h, c2.w, c2.c, side_ld[i + 2], P, DT_S8, 0);
                        if (tc_plan_fuse_pool(plan, pm, pm == 1 ? c2.input_quant_multipler : 0.f, qn)) {
                            prefilled[i + 2] = 1; pool_in_conv[i + 1] = 1;
                            e->not_materialised[i] = e->not_materialised[i + 1] = 1;
                        }
                    }
                    e->ops.push_back(Op{OP_CONV_TC_I8, i, [plan](cudaStream_t s) { tc_launch(plan, s); }});
                    break;
                }
                Int8P p{};
                p.q = q; p.out = tout;
                p.w = reinterpret_cast<const uint32_t *>(e->w_arena + cw[i].w_s8);
                p.bias = reinterpret_cast<const float *>(e->w_arena + cw[i].bias);
                p.alpha1 = alpha1;
                p.n = l.n; p.size = l.size; p.stride = l.stride; p.pad = l.pad; p.act = l.activation;
                p.CW = cpad / 4; p.M = M; p.acc_out = acc_dbg;
                dim3 grid((unsigned)((M + 63) / 64), (unsigned)((l.n + 63) / 64));
                e->ops.push_back(Op{OP_CONV_INT8, i, [p, grid](cudaStream_t s) { k_conv_int8_simt<<<grid, 256, 0, s>>>(p); }});
            }
            break;
        }
        case YB_MAXPOOL: {
            if (pool_in_conv[i]) break;    // done in the epilogue of the integer convolution in front of it
            need_prev();
            const TV tout = e->out_tv[i];
            const int size = l.size, stride = l.stride, pad = l.pad;
            // max-pool -> integer convolution: write the convolution's s8 / sign input directly (same values in the same
            // order as max-pool + quantise / binarise; the pooled f32 tensor never goes to HBM)
            if (opt.fuse && in_dt == DT_F32 && i + 1 < nl && cons[i].size() == 1 && cons[i][0] == i + 1 &&
                net->layers[i + 1].type == YB_CONVOLUTIONAL && conv_variant(i + 1) != 0 && side_off[i + 1] != (size_t)-1 &&
                !xnor_fallback(net->layers[i + 1]) && !getenv("YB_NO_POOL_FUSE")) {
                const Layer &c = net->layers[i + 1];
                const int v = conv_variant(i + 1);
                if (v == 2) {
                    TV q = make_tv(e->act_arena + side_off[i + 1], B, c.h, c.w, c.c, side_ld[i + 1], P, DT_S8, 0);
                    const float mult = c.input_quant_multipler;
                    const int g = grid_for((long)B * c.h * c.w * (q.ldc / 4));
                    e->ops.push_back(Op{OP_MAXPOOL, i, [tin, q, size, stride, pad, mult, g](cudaStream_t s) {
                        k_maxpool_fused<0><<<g, 256, 0, s>>>(tin, q, size, stride, pad, mult); }});
                } else if (xnor_on_tc(c)) {
                    TV q = make_tv(e->act_arena + side_off[i + 1], B, c.h, c.w, c.c, side_ld[i + 1], P, DT_S8, 0);
                    const int g = grid_for((long)B * c.h * c.w * (q.ldc / 4));
                    e->ops.push_back(Op{OP_MAXPOOL, i, [tin, q, size, stride, pad, g](cudaStream_t s) {
                        k_maxpool_fused<1><<<g, 256, 0, s>>>(tin, q, size, stride, pad, 0.f); }});
                } else {
                    const int CW = side_ld[i + 1];
                    TV bits = make_tv(e->act_arena + side_off[i + 1], B, c.h, c.w, CW, CW, P, DT_BITS, 0);
                    const int g = grid_for((long)B * c.h * c.w * CW);
                    e->ops.push_back(Op{OP_MAXPOOL, i, [tin, bits, size, stride, pad, g](cudaStream_t s) {
                        k_maxpool_fused<2><<<g, 256, 0, s>>>(tin, bits, size, stride, pad, 0.f); }});
                }
                prefilled[i + 1] = 1;
                e->not_materialised[i] = 1;      // fetch_layer reports it
                break;
            }
            const int g = grid_for((long)B * l.out_h * l.out_w * l.out_c);
            const int dt = in_dt;
            const int esz = (int)dt_size(dt);
            const bool vec = (l.out_c * esz) % 16 == 0 && (tin.ldc * esz) % 16 == 0 && (tout.ldc * esz) % 16 == 0 &&
                             (reinterpret_cast<uintptr_t>(tin.base) & 15) == 0 && (reinterpret_cast<uintptr_t>(tout.base) & 15) == 0;
            const int gv = grid_for((long)B * l.out_h * l.out_w * ((l.out_c * esz) / 16 + 1));
            e->ops.push_back(Op{OP_MAXPOOL, i, [tin, tout, size, stride, pad, g, dt, vec, gv](cudaStream_t s) {
                if (vec && dt == DT_F32) k_maxpool_vec<float><<<gv, 256, 0, s>>>(tin, tout, size, stride, pad);
                else if (vec) k_maxpool_vec<__nv_bfloat16><<<gv, 256, 0, s>>>(tin, tout, size, stride, pad);
                else if (dt == DT_F32) k_maxpool<float><<<g, 256, 0, s>>>(tin, tout, size, stride, pad);
                else k_maxpool<__nv_bfloat16><<<g, 256, 0, s>>>(tin, tout, size, stride, pad);
            }});
            break;
        }
        case YB_UPSAMPLE: {
            need_prev();
            if (l.reverse) fatal_throw("engine: reverse upsample (downsample) is not supported");
            const TV tout = e->out_tv[i];
            const int stride = l.stride; const float scale = l.scale;
            const int g = grid_for((long)B * l.out_h * l.out_w * l.out_c);
            const int dt = in_dt;
            const int esz = (int)dt_size(dt);
            const bool vec = scale == 1.f && (l.out_c * esz) % 16 == 0 && (tin.ldc * esz) % 16 == 0 && (tout.ldc * esz) % 16 == 0 &&
                             (reinterpret_cast<uintptr_t>(tin.base) & 15) == 0 && (reinterpret_cast<uintptr_t>(tout.base) & 15) == 0;
            const int gv = grid_for((long)B * l.out_h * l.out_w * ((l.out_c * esz) / 16 + 1));
            e->ops.push_back(Op{OP_UPSAMPLE, i, [tin, tout, stride, scale, g, dt, vec, gv, esz](cudaStream_t s) {
                if (vec) k_upsample_vec16<<<gv, 256, 0, s>>>(tin, tout, stride, esz);
                else if (dt == DT_F32) k_upsample<float><<<g, 256, 0, s>>>(tin, tout, stride, scale);
                else k_upsample<__nv_bfloat16><<<g, 256, 0, s>>>(tin, tout, stride, scale);
            }});
            break;
        }
        case YB_SHORTCUT: {
            if (is_fused_sc[i]) break;
            need_prev();
            const TV tout = e->out_tv[i];
            const TV from = e->out_tv[l.index];
            if (!from.base) fatal_throw("engine: shortcut source not placed");
            if (e->out_dt[l.index] != in_dt) fatal_throw("engine: shortcut dtype mismatch");
            // shortcut_cpu(batch, w1=l.w, h1=l.h, c1=l.c (from), add, w2=l.out_w, ...): yolov2_forward_network.c:410
            int stride = l.w / l.out_w, sample = l.out_w / l.w;
            if (stride < 1) stride = 1;
            if (sample < 1) sample = 1;
            const int minw = std::min(l.w, l.out_w), minh = std::min(l.h, l.out_h), minc = std::min(l.c, l.out_c);
            const int act = l.activation;
            const int g = grid_for((long)B * l.out_h * l.out_w * l.out_c);
            const int dt = in_dt;
            e->ops.push_back(Op{OP_SHORTCUT, i, [=](cudaStream_t s) {
                if (dt == DT_F32) k_shortcut<float><<<g, 256, 0, s>>>(tin, from, tout, stride, sample, minw, minh, minc, act);
                else k_shortcut<__nv_bfloat16><<<g, 256, 0, s>>>(tin, from, tout, stride, sample, minw, minh, minc, act);
            }});
            break;
        }
        case YB_ROUTE: {
            if (!has_nhwc_out(i)) fatal_throw("engine: route over layers of different spatial size is not supported");
            if (l.n == 1 && opt.fuse) break;   // alias
            int off = 0;
            for (int k = 0; k < l.n; ++k) {
                const int j = l.input_layers[k];
                const Layer &src = net->layers[j];
                if (!(home[j].owner == i)) {
                    const TV tsrc = e->out_tv[j];
                    if (!tsrc.base) fatal_throw("engine: route source not placed");
                    if (e->out_dt[j] != e->out_dt[i]) fatal_throw("engine: route dtype mismatch");
                    TV slice = e->out_tv[i];
                    slice.base += (size_t)off * dt_size(e->out_dt[i]);
                    slice.C = src.out_c;
                    const int g = grid_for((long)B * src.out_h * src.out_w * src.out_c);
                    const int dt = e->out_dt[i];
                    e->ops.push_back(Op{OP_ROUTE_COPY, i, [tsrc, slice, g, dt](cudaStream_t s) {
                        if (dt == DT_F32) k_copy_channels<float><<<g, 256, 0, s>>>(tsrc, slice);
                        else k_copy_channels<__nv_bfloat16><<<g, 256, 0, s>>>(tsrc, slice);
                    }});
                }
                off += src.out_c;
            }
            break;
        }
        case YB_REORG: {
            need_prev();
            if (l.reverse) fatal_throw("engine: reverse reorg is not supported");
            const TV tout = e->out_tv[i];
            const int stride = l.stride;
            const int g = grid_for((long)B * l.out_h * l.out_w * l.out_c);
            const int dt = in_dt;
            e->ops.push_back(Op{OP_REORG, i, [tin, tout, stride, g, dt](cudaStream_t s) {
                if (dt == DT_F32) k_reorg<float><<<g, 256, 0, s>>>(tin, tout, stride);
                else k_reorg<__nv_bfloat16><<<g, 256, 0, s>>>(tin, tout, stride);
            }});
            break;
        }
        case YB_YOLO: {
            if (yolo_fused[i]) break;   // written by the head convolution's epilogue
            need_prev();
            float *dst = e->d_final[i];
            const int classes = l.classes;
            const int g = grid_for((long)B * ((l.h * l.w + 31) / 32) * ((l.c + 31) / 32) * 256);
            const int dt = in_dt;
            const int fast = (ADT == DT_BF16) ? 1 : 0;
            e->ops.push_back(Op{OP_YOLO, i, [tin, dst, classes, g, dt, fast](cudaStream_t s) {
                if (dt == DT_F32) k_yolo<float><<<g, 256, 0, s>>>(tin, dst, classes, fast);
                else k_yolo<__nv_bfloat16><<<g, 256, 0, s>>>(tin, dst, classes, fast);
            }});
            break;
        }
        case YB_REGION: {
            need_prev();
            float *dst = e->d_final[i];
            const int n = l.n, classes = l.classes, coords = l.coords, softmax = l.softmax;
            const int g = grid_for((long)B * l.h * l.w * l.n);
            const int dt = in_dt;
            e->ops.push_back(Op{OP_REGION, i, [tin, dst, n, classes, coords, softmax, g, dt](cudaStream_t s) {
                if (dt == DT_F32) k_region<float><<<g, 256, 0, s>>>(tin, dst, n, classes, coords, softmax);
                else k_region<__nv_bfloat16><<<g, 256, 0, s>>>(tin, dst, n, classes, coords, softmax);
            }});
            break;
        }
        default:
            break;
        }
    }
    // last layer that is not yolo/region: keep an NCHW f32 copy as "the network output"
    {
        const int last = nl - 1;
        const Layer &l = net->layers[last];
        if (l.type != YB_YOLO && l.type != YB_REGION && e->d_final[last]) {
            int src = last;
            if (fused_into[last] >= 0) src = fused_into[last];
            const TV t = e->out_tv[src];
            if (t.base) {
                float *dst = e->d_final[last];
                const int dt = e->out_dt[src];
                const int g = grid_for((long)B * l.outputs);
                e->ops.push_back(Op{OP_YOLO, last, [t, dst, g, dt](cudaStream_t s) {
                    if (dt == DT_F32) k_nhwc_to_nchw_f32<float><<<g, 256, 0, s>>>(t, dst);
                    else k_nhwc_to_nchw_f32<__nv_bfloat16><<<g, 256, 0, s>>>(t, dst);
                }});
            }
        }
    }
    (void)E;
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaGetLastError());
    return e;
}

static void launch_input(Engine *e, const float *d_in, cudaStream_t s) { e->first_op(d_in, s); }
static void engine_forward_impl(Engine *e, const void *d_input, const unsigned char *d_u8_frames, void *stream);
void engine_forward(Engine *e, const void *d_input, void *stream) { engine_forward_impl(e, d_input, nullptr, stream); }

void engine_upload_input(Engine *e, const float *host_input, void *stream) {
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    CUDA_OK(cudaSetDevice(e->opt.device));
    CUDA_OK(cudaMemcpyAsync(e->d_input, host_input, e->input_count * sizeof(float), cudaMemcpyHostToDevice, s));
}

// u8 HWC images (all `w` x `h` x `c`) -> resized planar float in the engine's input staging buffer
void engine_upload_u8(Engine *e, const unsigned char *host_u8, int w, int h, int c, int net_w, int net_h, void *stream) {
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    CUDA_OK(cudaSetDevice(e->opt.device));
    const size_t bytes = (size_t)e->batch * w * h * c;
    if (bytes > e->u8_bytes) {
        if (e->d_u8) cudaFree(e->d_u8);
        CUDA_OK(cudaMalloc(&e->d_u8, bytes));
        e->u8_bytes = bytes;
    }
    CUDA_OK(cudaMemcpyAsync(e->d_u8, host_u8, bytes, cudaMemcpyHostToDevice, s));
    const long total = (long)e->batch * c * net_h * net_w;
    k_resize_u8_to_nchw<<<grid_for(total), 256, 0, s>>>(e->d_u8, e->batch, w, h, c, e->d_input, net_w, net_h);
    CUDA_OK(cudaGetLastError());
}

void *engine_stream(Engine *e) { return e->stream; }

static void engine_forward_impl(Engine *e, const void *d_input, const unsigned char *d_u8_frames, void *stream) {
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    CUDA_OK(cudaSetDevice(e->opt.device));   // thread identity may change per call (threading)
    const float *din = d_input ? reinterpret_cast<const float *>(d_input) : e->d_input;
    if (d_u8_frames) e->first_op_u8(d_u8_frames, s);   // stem straight from the 8-bit frames
    else launch_input(e, din, s);
    if (!e->graph_exec && !e->graph_failed && getenv("YB_NO_GRAPH")) e->graph_failed = true;   // profiling aid
    if (!e->graph_exec && !e->graph_failed) {
        // capture everything after the input conversion once
        cudaGraph_t graph = nullptr;
        cudaError_t st = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (st == cudaSuccess) {
            for (size_t k = 1; k < e->ops.size(); ++k) e->ops[k].launch(s);
            st = cudaStreamEndCapture(s, &graph);
        }
        if (st == cudaSuccess && graph) {
            st = cudaGraphInstantiate(&e->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
        }
        if (st != cudaSuccess || !e->graph_exec) {
            e->graph_failed = true;
            e->graph_exec = nullptr;
            cudaGetLastError();
        }
    }
    if (e->graph_exec) {
        CUDA_OK(cudaGraphLaunch(e->graph_exec, s));
    } else {
        for (size_t k = 1; k < e->ops.size(); ++k) e->ops[k].launch(s);
    }
    CUDA_OK(cudaGetLastError());
}

void engine_download_outputs(Engine *e, Network *net, void *stream) {
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    CUDA_OK(cudaSetDevice(e->opt.device));
    for (size_t i = 0; i < e->d_final.size(); ++i) {
        if (!e->d_final[i]) continue;
        CUDA_OK(cudaMemcpyAsync(e->h_final[i], e->d_final[i], e->final_count[i] * sizeof(float), cudaMemcpyDeviceToHost, s));
        net->layers[i].output = e->h_final[i];
        net->layers[i].output_count = e->final_count[i];
    }
    CUDA_OK(cudaStreamSynchronize(s));
}

static void ensure_slots(Engine *e) {
    if (!e->slots.empty()) return;
    const int NS = 3;
    CUDA_OK(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
    e->slots.resize(NS);
    for (auto &sl : e->slots) {
        CUDA_OK(cudaMalloc(&sl.d_in, e->input_count * sizeof(float)));
        sl.d_out.assign(e->d_final.size(), nullptr);
        sl.h_out.assign(e->d_final.size(), nullptr);
        for (size_t i = 0; i < e->d_final.size(); ++i) {
            if (!e->d_final[i]) continue;
            CUDA_OK(cudaMalloc(&sl.d_out[i], e->final_count[i] * sizeof(float)));
            CUDA_OK(cudaHostAlloc(&sl.h_out[i], e->final_count[i] * sizeof(float), cudaHostAllocDefault));
        }
        CUDA_OK(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&sl.ev_comp, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
        CUDA_OK(cudaEventCreateWithFlags(&sl.ev_det, cudaEventDisableTiming));
    }
    CUDA_OK(cudaStreamCreateWithFlags(&e->s_det, cudaStreamNonBlocking));
}

// Enqueue one batch: H2D on the copy-in stream, forward on the compute stream, D2H on the copy-out stream.
// Returns the ticket to pass to engine_collect.  Up to 3 batches may be in flight.
int engine_submit(Engine *e, const float *host_input) {
    CUDA_OK(cudaSetDevice(e->opt.device));
    ensure_slots(e);
    const int k = e->next_slot;
    Engine::Slot &sl = e->slots[k];
    if (sl.busy) fatal_throw("submit: pipeline full (3 batches in flight) -- collect the oldest ticket first");
    e->next_slot = (k + 1) % (int)e->slots.size();
    // the previous forward that read d_in[k] must have finished before it is overwritten
    CUDA_OK(cudaStreamWaitEvent(e->s_in, sl.ev_comp, 0));
    CUDA_OK(cudaMemcpyAsync(sl.d_in, host_input, e->input_count * sizeof(float), cudaMemcpyHostToDevice, e->s_in));
    CUDA_OK(cudaEventRecord(sl.ev_in, e->s_in));
    CUDA_OK(cudaStreamWaitEvent(e->stream, sl.ev_in, 0));
    engine_forward(e, sl.d_in, e->stream);
    // the previous D2H out of d_out[k] (or decode of it) must have finished before it is overwritten
    CUDA_OK(cudaStreamWaitEvent(e->stream, sl.ev_done, 0));
    CUDA_OK(cudaStreamWaitEvent(e->stream, sl.ev_det, 0));
    for (size_t i = 0; i < e->d_final.size(); ++i)
        if (e->d_final[i])
            CUDA_OK(cudaMemcpyAsync(sl.d_out[i], e->d_final[i], e->final_count[i] * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
    CUDA_OK(cudaEventRecord(sl.ev_comp, e->stream));
    CUDA_OK(cudaStreamWaitEvent(e->s_out, sl.ev_comp, 0));
    for (size_t i = 0; i < e->d_final.size(); ++i)
        if (e->d_final[i])
            CUDA_OK(cudaMemcpyAsync(sl.h_out[i], sl.d_out[i], e->final_count[i] * sizeof(float), cudaMemcpyDeviceToHost, e->s_out));
    CUDA_OK(cudaEventRecord(sl.ev_done, e->s_out));
    sl.busy = true; sl.mode = 0;
    return k;
}

// ptrs[i] = pinned host copy of layer i's output for this ticket (nullptr where the layer has none); valid until the slot
// is reused.  The form the multi-GPU batch call uses: several engines feed ONE host model.
void engine_collect_ptrs(Engine *e, int ticket, std::vector<const float *> &ptrs, std::vector<size_t> &counts) {
    if (ticket < 0 || ticket >= (int)e->slots.size() || !e->slots[ticket].busy || e->slots[ticket].mode != 0)
        fatal_throw("collect: bad ticket");
    CUDA_OK(cudaSetDevice(e->opt.device));
    Engine::Slot &sl = e->slots[ticket];
    CUDA_OK(cudaEventSynchronize(sl.ev_done));
    ptrs.assign(e->d_final.size(), nullptr);
    counts.assign(e->d_final.size(), 0);
    for (size_t i = 0; i < e->d_final.size(); ++i) {
        if (!e->d_final[i]) continue;
        ptrs[i] = sl.h_out[i];
        counts[i] = e->final_count[i];
    }
    sl.busy = false;
}

void engine_collect(Engine *e, Network *net, int ticket) {
    std::vector<const float *> ptrs; std::vector<size_t> counts;
    engine_collect_ptrs(e, ticket, ptrs, counts);
    for (size_t i = 0; i < ptrs.size(); ++i) {
        if (!ptrs[i]) continue;
        net->layers[i].output = const_cast<float *>(ptrs[i]);
        net->layers[i].output_count = counts[i];
    }
}

// ---- weight replication across the GPUs of one process (ONE broadcast of the prepared arena at init) ---------
// NCCL is bound at run time (dlopen): the library has no link-time dependency on it, and a host without NCCL -- or a device
// list with repeats, which NCCL refuses -- falls back to peer copies out of replica 0.
const char *engine_broadcast_arena(const std::vector<Engine *> &reps) {
    if (reps.size() < 2) return "single";
    const size_t bytes = reps[0]->w_bytes;
    for (Engine *r : reps) if (r->w_bytes != bytes) fatal_throw("broadcast: replicas disagree on the arena size");
    bool distinct = true;
    for (size_t a = 0; a < reps.size(); ++a)
        for (size_t b = a + 1; b < reps.size(); ++b) distinct &= reps[a]->opt.device != reps[b]->opt.device;
    typedef int (*InitAllFn)(void **, int, const int *);
    typedef int (*BcastFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    typedef int (*VoidFn)(void);
    typedef int (*DestroyFn)(void *);
    void *lib = (distinct && !getenv("YB_NO_NCCL")) ? dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL) : nullptr;
    if (lib) {
        InitAllFn init_all = (InitAllFn)dlsym(lib, "ncclCommInitAll");
        BcastFn bcast = (BcastFn)dlsym(lib, "ncclBroadcast");
        VoidFn gstart = (VoidFn)dlsym(lib, "ncclGroupStart"), gend = (VoidFn)dlsym(lib, "ncclGroupEnd");
        DestroyFn destroy = (DestroyFn)dlsym(lib, "ncclCommDestroy");
        if (init_all && bcast && gstart && gend && destroy) {
            std::vector<void *> comms(reps.size(), nullptr);
            std::vector<int> devs;
            for (Engine *r : reps) devs.push_back(r->opt.device);
            if (init_all(comms.data(), (int)reps.size(), devs.data()) == 0) {
                int rc = gstart();
                for (size_t k = 0; k < reps.size() && rc == 0; ++k) {
                    CUDA_OK(cudaSetDevice(reps[k]->opt.device));
                    rc = bcast(reps[k]->w_arena, reps[k]->w_arena, bytes, /*ncclChar*/ 0, /*root*/ 0, comms[k], reps[k]->stream);
                }
                rc |= gend();
                for (Engine *r : reps) { CUDA_OK(cudaSetDevice(r->opt.device)); CUDA_OK(cudaStreamSynchronize(r->stream)); }
                for (void *c : comms) if (c) destroy(c);
                if (rc == 0) return "nccl";
            }
        }
    }
    for (size_t k = 1; k < reps.size(); ++k) {
        CUDA_OK(cudaSetDevice(reps[0]->opt.device));
        CUDA_OK(cudaMemcpyPeer(reps[k]->w_arena, reps[k]->opt.device, reps[0]->w_arena, reps[0]->opt.device, bytes));
    }
    CUDA_OK(cudaDeviceSynchronize());
    return "peer-copy";
}
int engine_device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

void engine_fetch_layer(Engine *e, Network *net, int layer, float *dst) {
    CUDA_OK(cudaSetDevice(e->opt.device));
    const Layer &l = net->layers[layer];
    const size_t count = (size_t)l.outputs * e->batch;
    if (e->d_final[layer] && (l.type == YB_YOLO || l.type == YB_REGION)) {
        CUDA_OK(cudaMemcpy(dst, e->d_final[layer], count * sizeof(float), cudaMemcpyDeviceToHost));
        return;
    }
 const TV t = e->out_tv[layer];
    if (!t.base || e->not_materialised[layer]) fatal_throw("fetch_layer: layer " + std::to_string(layer) + " has no materialised output "
                             "(fused or aliased away; build the engine with fusion off)");
    float *tmp = nullptr;
    CUDA_OK(cudaMalloc(&tmp, count * sizeof(float)));
    const int g = grid_for((long)count);
    if (e->out_dt[layer] == DT_F32) k_nhwc_to_nchw_f32<float><<<g, 256, 0, e->stream>>>(t, tmp);
    else k_nhwc_to_nchw_f32<__nv_bfloat16><<<g, 256, 0, e->stream>>>(t, tmp);
    CUDA_OK(cudaMemcpyAsync(dst, tmp, count * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    cudaFree(tmp);
}

void engine_fetch_input(Engine *e, float *dst) {
    CUDA_OK(cudaSetDevice(e->opt.device));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    CUDA_OK(cudaMemcpy(dst, e->d_input, e->input_count * sizeof(float), cudaMemcpyDeviceToHost));
}

int engine_fetch_counts(Engine *e, int layer, int32_t *dst, size_t count) {
    if (layer < 0 || layer >= (int)e->d_counts.size() || !e->d_counts[layer]) return -1;
    if (count < e->counts_count[layer]) return -2;
    CUDA_OK(cudaSetDevice(e->opt.device));
    CUDA_OK(cudaMemcpy(dst, e->d_counts[layer], e->counts_count[layer] * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return (int)e->counts_count[layer];
}

void engine_weight_arena(Engine *e, void **ptr, size_t *bytes) { *ptr = e->w_arena; *bytes = e->w_bytes; }
// INT8 input calibration: |x| histogram of the INPUT of layer `layer` (image `img` of the batch) after a
// forward, binned like the reference's entropy_calibration.  hist: host uint32[max_bin].
void engine_input_histogram(Engine *e, Network *net, int layer, int img, float bin_width, int max_bin, uint32_t *hist) {
    CUDA_OK(cudaSetDevice(e->opt.device));
    if (max_bin < 129 || max_bin > 4096) fatal_throw("calibrate: max_bin must be in 129..4096");
    if (img < 0 || img >= e->batch) fatal_throw("calibrate: image index out of range");
    unsigned *d_hist = nullptr;
    CUDA_OK(cudaMalloc(&d_hist, (size_t)max_bin * sizeof(unsigned)));
    CUDA_OK(cudaMemsetAsync(d_hist, 0, (size_t)max_bin * sizeof(unsigned), e->stream));
    if (layer == 0) {
        const long n = (long)net->c * net->h * net->w;
        k_abs_hist_flat<<<grid_for(n), 256, 0, e->stream>>>(e->d_input + (size_t)img * n, n, bin_width, max_bin, d_hist);
    } else {
        const TV t = e->out_tv[layer - 1];
        if (!t.base || e->not_materialised[layer - 1]) { cudaFree(d_hist); fatal_throw("calibrate: the input of layer " + std::to_string(layer) +
                                                     " is not materialised (fused away; set option fuse=0)"); }
        const long n = (long)t.C * t.H * t.W;
        if (e->out_dt[layer - 1] == DT_F32) k_abs_hist<float><<<grid_for(n), 256, 0, e->stream>>>(t, img, bin_width, max_bin, d_hist);
        else k_abs_hist<__nv_bfloat16><<<grid_for(n), 256, 0, e->stream>>>(t, img, bin_width, max_bin, d_hist);
    }
    CUDA_OK(cudaMemcpyAsync(hist, d_hist, (size_t)max_bin * sizeof(unsigned), cudaMemcpyDeviceToHost, e->stream));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    cudaFree(d_hist);
}

// ---- batched decode + NMS on the device (yb_detect.cuh) ---------------------------------------------------------
static constexpr int DET_MAX_ROWS = 16384;   // the per-(class, image) sort lives in shared memory: 8 B per (power-of-two) row

static DetParams det_params(Engine *e, Network *net, const std::vector<float *> &finals, int w, int h, float thresh, float nms,
                            int relative, int letter, int max_rows) {
    if (max_rows <= 0 || max_rows > DET_MAX_ROWS) fatal_throw("detect: max_rows must be in 1.." + std::to_string(DET_MAX_ROWS));
    DetParams P{};
    int total = 0;
    for (size_t i = 0; i < net->layers.size(); ++i) {
        const Layer &l = net->layers[i];
        if (l.type != YB_YOLO && l.type != YB_REGION) continue;
        if (!finals[i]) fatal_throw("detect: detection layer has no device output");
        if (P.nl == DET_MAX_LAYERS) fatal_throw("detect: too many detection layers");
        if (l.n > DET_MAX_ANCHORS) fatal_throw("detect: too many anchors per layer");
        if (P.nl && l.classes != P.classes) fatal_throw("detect: detection layers disagree on the class count");
        if ((l.type == YB_YOLO && (int)l.mask.size() < l.n) || (int)l.anchors.size() < 2 * l.n)
            fatal_throw("detect: detection layer without mask / anchors");
        DetLayer &d = P.L[P.nl++];
        d.p = finals[i]; d.type = l.type; d.w = l.w; d.h = l.h; d.n = l.n; d.classes = l.classes; d.outputs = l.outputs;
        d.base = total; d.nbox = l.w * l.h * l.n; total += d.nbox;
        for (int a = 0; a < l.n; ++a) {
            const int k = (l.type == YB_YOLO) ? l.mask[a] : a;
            d.aw[a] = l.anchors[2 * k]; d.ah[a] = l.anchors[2 * k + 1];
        }
        P.classes = l.classes;
    }
    if (!P.nl) fatal_throw("detect: the network has no yolo / region layer");
    P.total = total; P.netw = net->w; P.neth = net->h; P.imw = w; P.imh = h; P.relative = relative;
    P.new_w = net->w; P.new_h = net->h;
    if (letter) {   // correct_yolo_boxes, additionally.c:4287-4296
        if (((float)net->w / w) < ((float)net->h / h)) { P.new_w = net->w; P.new_h = (h * net->w) / w; }
        else { P.new_h = net->h; P.new_w = (w * net->h) / h; }
    }
    P.thresh = thresh; P.nms = nms; P.max_rows = max_rows; P.nblk = (total + 255) / 256;
    (void)e;
    return P;
}

static void det_ws_ensure(Engine::DetWs &ws, int B, const DetParams &P) {
    const int stride = 5 + P.classes, words = (P.max_rows + 31) / 32;
    if (ws.cap == P.max_rows && ws.stride == stride && ws.nblk >= P.nblk) return;   // pitch == cap == max_rows: the NMS sees
                                                                                    // exactly the rows the caller asked for
    if (ws.rows) cudaFree(ws.rows);
    if (ws.mask) cudaFree(ws.mask);
    if (ws.blkcnt) cudaFree(ws.blkcnt);
    if (ws.counts) cudaFree(ws.counts);
    CUDA_OK(cudaMalloc(&ws.rows, (size_t)B * P.max_rows * stride * sizeof(float)));
    CUDA_OK(cudaMalloc(&ws.mask, (size_t)B * P.max_rows * words * sizeof(unsigned)));
    CUDA_OK(cudaMalloc(&ws.blkcnt, (size_t)B * P.nblk * sizeof(int)));
    CUDA_OK(cudaMalloc(&ws.counts, (size_t)B * sizeof(int)));
    ws.cap = P.max_rows; ws.stride = stride; ws.nblk = P.nblk;
}

static void det_launch_count_emit(const DetParams &P, Engine::DetWs &ws, int B, cudaStream_t s) {
    k_det_count<<<dim3((unsigned)P.nblk, (unsigned)B), 256, 0, s>>>(P, ws.blkcnt);
    k_det_emit<<<dim3((unsigned)P.nblk, (unsigned)B), 256, 0, s>>>(P, ws.blkcnt, ws.rows, ws.counts);
}
// nmax: upper bound of the candidates of any image (the kernels read the true counts on the device and idle beyond them)
static void det_launch_nms(const DetParams &P, Engine::DetWs &ws, int B, int nmax, cudaStream_t s) {
    if (!(P.nms > 0.f) || nmax <= 0) return;
    const int capw = (ws.cap + 31) / 32;
    k_det_iou<<<dim3((unsigned)((capw + 127) / 128), (unsigned)std::min(nmax, 256), (unsigned)B), 128, 0, s>>>(P, ws.rows, ws.counts, ws.mask);
    int P2 = 1; while (P2 < nmax) P2 <<= 1;
    const size_t smem = (size_t)P2 * 8 + (size_t)capw * 4;
    if (smem > 48 * 1024)
        CUDA_OK(cudaFuncSetAttribute(k_det_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_det_nms<<<dim3((unsigned)P.classes, (unsigned)B), 256, smem, s>>>(P, ws.rows, ws.counts, ws.mask, P2);
}

// Synchronous form.  rows: [batch][max_rows][5 + classes]; counts[b] = candidates of image b before the max_rows cap.
// Returns 5 + classes.
int engine_detect(Engine *e, Network *net, int w, int h, float thresh, float nms, int relative, int letter,
                  float *rows, int max_rows, int *counts) {
    CUDA_OK(cudaSetDevice(e->opt.device));
    const DetParams P = det_params(e, net, e->d_final, w, h, thresh, nms, relative, letter, max_rows);
    const int B = e->batch, stride = 5 + P.classes;
    det_ws_ensure(e->det, B, P);
    cudaStream_t s = e->stream;
    det_launch_count_emit(P, e->det, B, s);
    std::vector<int> hc(B);
    CUDA_OK(cudaMemcpyAsync(hc.data(), e->det.counts, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    int nmax = 0;
    for (int b = 0; b < B; ++b) { counts[b] = hc[b]; nmax = std::max(nmax, std::min(hc[b], max_rows)); }
    det_launch_nms(P, e->det, B, nmax, s);
    for (int b = 0; b < B; ++b) {
        const int n = std::min(hc[b], max_rows);
        if (n > 0)
            CUDA_OK(cudaMemcpyAsync(rows + (size_t)b * max_rows * stride, e->det.rows + (size_t)b * max_rows * stride,
                                    (size_t)n * stride * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    CUDA_OK(cudaStreamSynchronize(s));
    CUDA_OK(cudaGetLastError());
    return stride;
}

// ---- pipelined detection path (SURVEY 8f rows 1 + 2 in the serving loop) -----------------------------------------
// One call enqueues, for one batch of 8-bit frames: H2D of the frames + the reference's resize on the copy-in stream, the
// forward on the compute stream, decode + NMS on a side stream (under the forward of the NEXT batch), and returns a ticket.
// engine_collect_detections waits for that batch and copies back exactly the candidate rows.  Host traffic per batch: the
// u8 frames in (a quarter of the float images), counts + rows out (a few hundred KB instead of the 124 MB of yolo tensors).
int engine_submit_u8(Engine *e, Network *net, const unsigned char *host_u8, int w, int h, float thresh, float nms,
                     int relative, int letter, int max_rows) {
    CUDA_OK(cudaSetDevice(e->opt.device));
    ensure_slots(e);
    const int k = e->next_slot;
    Engine::Slot &sl = e->slots[k];
    if (sl.busy) fatal_throw("submit: pipeline full (3 batches in flight) -- collect the oldest ticket first");
    const DetParams P0 = det_params(e, net, sl.d_out, w, h, thresh, nms, relative, letter, max_rows);
    const int B = e->batch, stride = 5 + P0.classes;
    det_ws_ensure(sl.det, B, P0);
    const size_t rows_bytes = (size_t)B * max_rows * stride * sizeof(float);
    if (sl.h_rows_bytes < rows_bytes) {
        if (sl.h_rows) cudaFreeHost(sl.h_rows);
        CUDA_OK(cudaHostAlloc(&sl.h_rows, rows_bytes, cudaHostAllocDefault));
        sl.h_rows_bytes = rows_bytes;
    }
    if (!sl.h_counts) CUDA_OK(cudaHostAlloc(&sl.h_counts, (size_t)B * sizeof(int), cudaHostAllocDefault));
    const size_t bytes = (size_t)B * w * h * net->c;
    if (bytes > sl.u8_bytes) {
        if (sl.d_u8) cudaFree(sl.d_u8);
        CUDA_OK(cudaMalloc(&sl.d_u8, bytes));
        sl.u8_bytes = bytes;
    }
    e->next_slot = (k + 1) % (int)e->slots.size();
    // the previous forward that read d_in[k] must have finished before it is overwritten
    CUDA_OK(cudaStreamWaitEvent(e->s_in, sl.ev_comp, 0));
    CUDA_OK(cudaMemcpyAsync(sl.d_u8, host_u8, bytes, cudaMemcpyHostToDevice, e->s_in));
    const bool direct = e->first_op_u8 && w == net->w && h == net->h && net->c == 3;   // frames of the network size: no staging
    if (!direct) {
        const long total = (long)B * net->c * net->h * net->w;
        k_resize_u8_to_nchw<<<grid_for(total), 256, 0, e->s_in>>>(sl.d_u8, B, w, h, net->c, sl.d_in, net->w, net->h);
    }
    CUDA_OK(cudaEventRecord(sl.ev_in, e->s_in));
    CUDA_OK(cudaStreamWaitEvent(e->stream, sl.ev_in, 0));
    engine_forward_impl(e, sl.d_in, direct ? sl.d_u8 : nullptr, e->stream);
    // Candidate selection + box decode (k_det_count / k_det_emit: they read the objectness planes and, for the few candidates,
    // their class scores) run right behind the forward on the compute stream, straight on the engine's yolo tensors -- the next
    // forward overwrites those, so this is the only part that must not slip.  What follows (IoU matrix + per-class NMS) works on
    // the slot's own candidate rows and goes to the side stream, where it overlaps the next batch's forward.  (Copying the
    // 124 MB of yolo tensors into the slot first, as the raw-tensor path does, cost more than the decode itself.)
    CUDA_OK(cudaStreamWaitEvent(e->stream, sl.ev_det, 0));     // the slot's previous NMS / counts copy are done with its workspace
    DetParams P1 = P0;
    {
        int k2 = 0;
        for (size_t i = 0; i < net->layers.size(); ++i)
            if (net->layers[i].type == YB_YOLO || net->layers[i].type == YB_REGION) P1.L[k2++].p = e->d_final[i];
    }
    det_launch_count_emit(P1, sl.det, B, e->stream);
    CUDA_OK(cudaEventRecord(sl.ev_comp, e->stream));
    CUDA_OK(cudaStreamWaitEvent(e->s_det, sl.ev_comp, 0));
    CUDA_OK(cudaMemcpyAsync(sl.h_counts, sl.det.counts, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, e->s_det));
    det_launch_nms(P0, sl.det, B, max_rows, e->s_det);   // no host round trip: grids sized for the cap, kernels read the counts
    CUDA_OK(cudaEventRecord(sl.ev_det, e->s_det));
    CUDA_OK(cudaGetLastError());
    sl.busy = true; sl.mode = 1;
    return k;
}

// rows: pinned [batch][max_rows][5 + classes] (valid until the slot is reused), counts[batch]; returns 5 + classes.
int engine_collect_detections(Engine *e, int ticket, const float **rows, const int **counts, size_t *d2h_bytes) {
    if (ticket < 0 || ticket >= (int)e->slots.size() || !e->slots[ticket].busy || e->slots[ticket].mode != 1)
        fatal_throw("collect_detections: bad ticket");
    CUDA_OK(cudaSetDevice(e->opt.device));
    Engine::Slot &sl = e->slots[ticket];
    CUDA_OK(cudaEventSynchronize(sl.ev_det));
    const int B = e->batch, cap = sl.det.cap, stride = sl.det.stride;
    size_t moved = (size_t)B * sizeof(int);
    for (int b = 0; b < B; ++b) {
        const int n = std::min(sl.h_counts[b], cap);
        if (n > 0) {
            CUDA_OK(cudaMemcpyAsync(sl.h_rows + (size_t)b * cap * stride, sl.det.rows + (size_t)b * cap * stride,
                                    (size_t)n * stride * sizeof(float), cudaMemcpyDeviceToHost, e->s_out));
            moved += (size_t)n * stride * sizeof(float);
        }
    }
    CUDA_OK(cudaStreamSynchronize(e->s_out));
    if (rows) *rows = sl.h_rows;
    if (counts) *counts = sl.h_counts;
    if (d2h_bytes) *d2h_bytes = moved;
    sl.busy = false;
    return stride;
}

int engine_num_launches(Engine *e) { return (int)e->ops.size(); }
long engine_info(Engine *e, const char *key) {
    if (!strcmp(key, "launches")) return (long)e->ops.size();
    if (!strcmp(key, "tc_layers")) return e->n_tc;
    if (!strcmp(key, "ksplit_layers")) return e->n_ksplit;
    return -1;
}

int engine_profile(Engine *e, const void *d_input, int *layer_idx, int *op_kind, float *ms, int max) {
    CUDA_OK(cudaSetDevice(e->opt.device));
    cudaStream_t s = e->stream;
    const float *din = d_input ? reinterpret_cast<const float *>(d_input) : e->d_input;
    const int n = (int)e->ops.size();
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto &x : ev) CUDA_OK(cudaEventCreate(&x));
    for (int rep = 0; rep < 2; ++rep) {   // second pass is the measured one
        CUDA_OK(cudaEventRecord(ev[0], s));
        for (int k = 0; k < n; ++k) {
            if (k == 0) launch_input(e, din, s);
            else e->ops[k].launch(s);
            CUDA_OK(cudaEventRecord(ev[k + 1], s));
        }
        CUDA_OK(cudaStreamSynchronize(s));
    }
    for (int k = 0; k < n && k < max; ++k) {
        layer_idx[k] = e->ops[k].layer;
        op_kind[k] = e->ops[k].kind;
        CUDA_OK(cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
    }
    for (auto &x : ev) cudaEventDestroy(x);
    return n;
}

}  // namespace yb

extern "C" void *yb_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void yb_free_pinned(void *p) { if (p) cudaFreeHost(p); }
