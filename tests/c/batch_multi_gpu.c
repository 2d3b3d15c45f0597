/*
 * batch_multi_gpu.c -- plain C host program (no Python, no torch, no CUDA headers): drives the engine through the C ABI of
 * include/yolo2_light_b200.h the way the reference's main.c:160-219 drives its own code, on 1 GPU and then on 2 replicas
 * (2 GPUs when the box has them, else two replicas on GPU 0), and checks that image k of the multi-GPU batch call is
 * bit-identical to what the single-GPU predict returns for it.  SURVEY 8b "Batch extension", 8e.
 *
 *   usage: batch_multi_gpu <cfg> <weights> <batch> <nimg> [quantized]        exit code 0 = pass
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "yolo2_light_b200.h"

static float frand(unsigned *st) { *st = *st * 1664525u + 1013904223u; return (float)((*st >> 8) & 0xFFFF) / 65536.f; }

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s cfg weights batch nimg [quantized]\n", argv[0]); return 2; }
    const int batch = atoi(argv[3]), nimg = atoi(argv[4]), q = argc > 5 ? atoi(argv[5]) : 0;
    yb_set_abort_on_error(0);
    /* the reference's preparation sequence, main.c:160-171 */
    yb_network *net = yb_parse_network_cfg(argv[1], batch, q);
    if (!net) { fprintf(stderr, "parse: %s\n", yb_last_error()); return 1; }
    if (yb_load_weights_upto(net, argv[2], 1 << 30) != 0) { fprintf(stderr, "weights: %s\n", yb_last_error()); return 1; }
    yb_fuse_conv_batchnorm(net);
    yb_calculate_binary_weights(net);
    if (q) yb_quantinization_and_get_multipliers(net);
    int dims[8];
    yb_network_dims(net, dims);
    const int nl = dims[0], per_img = dims[5];
    float *images = (float *)malloc(sizeof(float) * (size_t)nimg * per_img);
    unsigned st = 7u;
    for (size_t i = 0; i < (size_t)nimg * per_img; ++i) images[i] = frand(&st);

    /* ---- single GPU, batch by batch: the expected per-image results of every yolo / region layer --------------------- */
    float **expect = (float **)calloc(nl, sizeof(float *));
    int *outs = (int *)calloc(nl, sizeof(int));
    float *padded = (float *)calloc((size_t)batch * per_img, sizeof(float));
    for (int first = 0; first < nimg; first += batch) {
        const int cnt = nimg - first < batch ? nimg - first : batch;
        memset(padded, 0, sizeof(float) * (size_t)batch * per_img);
        memcpy(padded, images + (size_t)first * per_img, sizeof(float) * (size_t)cnt * per_img);
        float *r = q ? yb_network_predict_quantized(net, padded) : yb_network_predict(net, padded);
        if (!r) { fprintf(stderr, "predict: %s\n", yb_last_error()); return 1; }
        for (int i = 0; i < nl; ++i) {
            int count = 0;
            const float *o = yb_network_layer_output(net, i, &count);
            if (!o) continue;
            outs[i] = count / batch;
            if (!expect[i]) expect[i] = (float *)malloc(sizeof(float) * (size_t)nimg * outs[i]);
            memcpy(expect[i] + (size_t)first * outs[i], o, sizeof(float) * (size_t)cnt * outs[i]);
        }
    }

    /* ---- two replicas: GPUs {0, 1} when there are two, else {0, 0} ------------------------------------------------------- */
    int devs[2] = {0, 1};
    int two_gpus = yb_network_set_devices(net, devs, 2) == 0;
    if (!two_gpus) { devs[1] = 0; if (yb_network_set_devices(net, devs, 2) != 0) { fprintf(stderr, "set_devices: %s\n", yb_last_error()); return 1; } }
    if (yb_network_predict_batch(net, images, nimg, 2, q) != 0) { fprintf(stderr, "predict_batch: %s\n", yb_last_error()); return 1; }
    int checked = 0;
    for (int i = 0; i < nl; ++i) {
        int per = 0;
        const float *got = yb_network_batch_output(net, i, &per);
        if (!got) continue;
        if (!expect[i] || per != outs[i]) { fprintf(stderr, "layer %d: missing / mis-sized output (%d vs %d)\n", i, per, outs[i]); return 1; }
        if (memcmp(got, expect[i], sizeof(float) * (size_t)nimg * per) != 0) {
            for (size_t k = 0; k < (size_t)nimg * per; ++k)
                if (memcmp(&got[k], &expect[i][k], 4)) { fprintf(stderr, "layer %d: image %zu element %zu differs: %g vs %g\n", i, k / per, k % per, got[k], expect[i][k]); break; }
            return 1;
        }
        ++checked;
    }
    printf("batch_multi_gpu: %d images, batch %d, %s, replication=%s, %d output layers bit-identical to the 1-GPU path\n", nimg, batch,
           two_gpus ? "GPUs {0,1}" : "two replicas on GPU 0", yb_network_replication(net), checked);
    yb_free_network(net);
    return checked > 0 ? 0 : 1;
}
