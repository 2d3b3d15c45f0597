// yb_conv_tc.cuh -- interface of the tensor-core (tcgen05 / TMEM / TMA) implicit-GEMM convolutions
// (implemented in yb_conv_tc.cu).
#pragma once
#include <cuda_runtime.h>

#include "yb_kernels.cuh"
#include "yb_model.h"

namespace yb {

// non-zero if the bf16 tcgen05 kernel takes this layer (input view `in`, bf16 or f32 output)
int tc_conv_supported(const Layer &l, const TV &in, const TV &out, bool out_bf16);
// builds the per-layer launch state (TMA tensor maps, tile schedule); throws yb::Error on failure.
// wide_rows: prefer wide pixel tiles (the plan will get tc_plan_fuse_yolo: NCHW plane stores in the epilogue)
// no_halo: keep the one-TMA-box-per-tap schedule for 3x3 layers (the K-split tail only exists there)
void *tc_make_plan(const Layer &l, const TV &in, const TV &out, bool out_bf16, const TV &res, bool res_bf16,
                   int act2, const void *d_weights_bf16, int ldn, const float *d_bias, int wide_rows = 0, int no_halo = 0);
// kind::tf32 variant for the FP32 detection heads of the exact (INT8 / XNOR) networks: f32 in, f32 [ldn][K] weights, f32 out
int tc_tf32_supported(const Layer &l, const TV &in, const TV &out);
void *tc_make_plan_tf32(const Layer &l, const TV &in, const TV &out, const void *d_weights_f32, int ldn, const float *d_bias,
                        int wide_rows);
// INT8 (kind::i8) variant
int tc_i8_supported(const Layer &l, const TV &q, const TV &out);
void *tc_make_plan_i8(const Layer &l, const TV &q, const TV &out, const void *d_weights_s8, int ldn, const float *d_bias,
                      float alpha1, int *acc_out, int want_halo = 0);
// XNOR layer as +-1 s8 on kind::i8 (q: s8 activation with -1 borders)
void *tc_make_plan_xnor(const Layer &l, const TV &q, const TV &out, const void *d_weights_pm1, int ldn, const float *d_bias,
                        const float *d_mean, int *counts_out, int want_halo = 0);
// fuse the following 2x2/2 max-pool + the next integer layer's input conversion (1: s8 quantised, 2: +-1 bytes) into an integer plan
int tc_plan_fuse_pool(void *plan, int mode, float mult, const TV &qnext);
// fuse the following [yolo] layer into the (f32-output) plan: logistic + NCHW store in the epilogue
void tc_plan_fuse_yolo(void *plan, float *d_yolo_nchw, int classes);
// K-split of the tail wave of a bf16 plan (wave quantisation): `ws` (tc_ksplit_ws_bytes) and `flags`
// (tc_ksplit_flag_bytes, zeroed) belong to the caller and may be shared by all plans that run on one stream.
// Returns 1 if the plan's schedule was changed.
size_t tc_ksplit_ws_bytes(int sms);
size_t tc_ksplit_flag_bytes(int sms);
int tc_plan_enable_ksplit(void *plan, float *ws, unsigned *flags);
// tensor-core stem (3-channel 3x3 from the caller's NCHW f32 image, bf16 NHWC out)
int tc_stem_supported(const Layer &l, const TV &out);
void *tc_stem_make_plan(const Layer &l, const TV &out, const void *d_w_32x32_bf16, const float *d_bias);
void tc_stem_launch(void *plan, const float *d_in_nchw, cudaStream_t s);
void tc_stem_launch_u8(void *plan, const unsigned char *d_in_hwc, cudaStream_t s);   // frames already of the network size
void tc_stem_free_plan(void *plan);
int tc_plan_cta_group(void *plan);   // 1 or 2 (k_conv_tc<2>, CTA pairs)
void tc_launch(void *plan, cudaStream_t s);
void tc_free_plan(void *plan);

}  // namespace yb
