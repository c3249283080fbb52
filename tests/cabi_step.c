/* A training loop of BASELINE config 1 (README CNN: Conv2D(8, 2x2, VALID, relu) -> MaxPooling2D(2x2) ->
 * Flatten -> Dense(10, softmax), SGD lr 0.01, batch 256) written against NOTHING but the C ABI of
 * include/shinnosuke_b200.h: no Python, no PyTorch.  It is what a reference maintainer's binding has
 * to do for models.py:70-92 (forward sweep, loss.backward, backward sweep, optimizer.update):
 *   step 1 runs eagerly, step 2 is captured into a CUDA graph (sn_graph_begin / sn_graph_end) and
 *   launched, step 3 replays that graph on the next minibatch.
 * tests/test_gpu_cabi_step.py compiles this file with gcc, feeds it the data and seed-0 weights of the
 * live-reference fixture tests/golden/traj_cfg1.npz (KAT D) and compares the printed
 * losses / accuracies / final filters with the reference's.
 *
 * usage: cabi_step <dir>      (reads x.bin y.bin conv_w.bin conv_b.bin dense_w.bin, little-endian float32)
 * Layouts: x (768,28,28,1) NHWC == the reference's NCHW for one channel; conv_w [F][kh][kw][C]; dense_w
 * [n_out][13*13*8] with the inputs in the device's (i,j,c) order (the harness permutes the reference's
 * (c,i,j) rows): Flatten is then a view of the pooled NHWC map. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/shinnosuke_b200.h"

#define CHECK(call)                                                                    \
    do {                                                                               \
        int e_ = (call);                                                               \
        if (e_ != SN_OK) {                                                             \
            fprintf(stderr, "%s failed (%d): %s\n", #call, e_, sn_last_error());       \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

enum { N = 256, STEPS = 3, H = 28, W = 28, F = 8, K = 2, OH = 27, OW = 27, PH = 13, PW = 13, NIN = 13 * 13 * 8, NOUT = 10 };

static float *read_bin(const char *dir, const char *name, size_t count) {
    char path[1024];
    snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    float *p = (float *)malloc(count * sizeof(float));
    if (fread(p, sizeof(float), count, f) != count) { fprintf(stderr, "%s: short read\n", path); exit(2); }
    fclose(f);
    return p;
}
static size_t pad32(size_t n) { return (n + 31) / 32 * 32; }

/* device buffers of the step */
static float *x, *y1h, *conv_out, *pool_out, *probs, *dz, *dpool, *dconv, *metrics, *arena_w, *arena_g;
static unsigned char *code;
static size_t o_cw, o_cb, o_dw, o_db, n_arena;
static void *st;

static int step_body(void) {
    /* forward sweep (models.py:124-129) */
    CHECK(sn_conv2d_fprop(x, arena_w + o_cw, arena_w + o_cb, conv_out, N, 1, H, W, F, K, K, 1, 0, 0, SN_ACT_RELU, SN_PREC_FP32, st));
    CHECK(sn_maxpool2d_fwd(conv_out, pool_out, code, N, OH, OW, F, 2, 2, SN_POOL_FIRST_ARGMAX, st));
    CHECK(sn_dense_head_fprop(pool_out, arena_w + o_dw, arena_w + o_db, probs, N, NIN, NOUT, 1, st));
    /* loss.backward + loss / accuracy sums (utils/Objectives.py:72-91) */
    CHECK(sn_memset(metrics, 0, 2 * sizeof(float), st));
    CHECK(sn_scce_bwd(probs, y1h, dz, metrics, N, NOUT, 0, st));
    /* backward sweep (models.py:82-83): Dense, MaxPooling2D, Relu, Conv2D */
    CHECK(sn_dense_head_bwd(pool_out, dz, arena_w + o_dw, arena_g + o_dw, arena_g + o_db, dpool, N, NIN, NOUT, 0, st));
    CHECK(sn_maxpool2d_bwd(dpool, code, dconv, N, OH, OW, F, 2, 2, SN_POOL_FIRST_ARGMAX, 0, st));
    CHECK(sn_relu_bwd(dconv, conv_out, (long long)N * OH * OW * F, st));
    CHECK(sn_conv2d_wgrad(x, dconv, arena_g + o_cw, arena_g + o_cb, N, 1, H, W, F, K, K, 1, 0, 0, 1, SN_PREC_FP32, st));
    /* optimizer.update (utils/Optimizers.py:29-32): one kernel over the flat arena, gradients zeroed */
    CHECK(sn_sgd_update(arena_w, arena_g, (long long)n_arena, 0.01f, 1.0f, st));
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <dir>\n", argv[0]); return 2; }
    const char *dir = argv[1];
    float *hx = read_bin(dir, "x.bin", (size_t)STEPS * N * H * W);
    float *hy = read_bin(dir, "y.bin", (size_t)STEPS * N * NOUT);
    float *hcw = read_bin(dir, "conv_w.bin", F * K * K), *hcb = read_bin(dir, "conv_b.bin", F);
    float *hdw = read_bin(dir, "dense_w.bin", (size_t)NOUT * NIN);
    int ndev = 0;
    CHECK(sn_device_count(&ndev));
    if (ndev < 1) { fprintf(stderr, "no CUDA device\n"); return 3; }
    CHECK(sn_set_device(0));
    CHECK(sn_stream_create(&st));
    o_cw = 0; o_cb = o_cw + pad32(F * K * K); o_dw = o_cb + pad32(F); o_db = o_dw + pad32((size_t)NOUT * NIN);
    n_arena = o_db + pad32(NOUT);
#define DMALLOC(p, count) CHECK(sn_malloc((void **)&(p), (count) * sizeof(*(p))))
    DMALLOC(arena_w, n_arena); DMALLOC(arena_g, n_arena);
    DMALLOC(x, (size_t)N * H * W); DMALLOC(y1h, (size_t)N * NOUT);
    DMALLOC(conv_out, (size_t)N * OH * OW * F); DMALLOC(dconv, (size_t)N * OH * OW * F);
    DMALLOC(pool_out, (size_t)N * NIN); DMALLOC(dpool, (size_t)N * NIN); DMALLOC(code, (size_t)N * NIN);
    DMALLOC(probs, (size_t)N * NOUT); DMALLOC(dz, (size_t)N * NOUT); DMALLOC(metrics, 2);
    CHECK(sn_memset(arena_w, 0, n_arena * sizeof(float), st));
    CHECK(sn_memset(arena_g, 0, n_arena * sizeof(float), st));
    CHECK(sn_memcpy_h2d(arena_w + o_cw, hcw, F * K * K * sizeof(float), st));
    CHECK(sn_memcpy_h2d(arena_w + o_cb, hcb, F * sizeof(float), st));
    CHECK(sn_memcpy_h2d(arena_w + o_dw, hdw, (size_t)NOUT * NIN * sizeof(float), st));      /* Dense bias starts at zero (layers/FC.py:108) */
    CHECK(sn_stream_sync(st));

    void *graph = NULL;
    sn_launch_count_reset();
    for (int s = 0; s < STEPS; ++s) {
        CHECK(sn_memcpy_h2d(x, hx + (size_t)s * N * H * W, (size_t)N * H * W * sizeof(float), st));
        CHECK(sn_memcpy_h2d(y1h, hy + (size_t)s * N * NOUT, (size_t)N * NOUT * sizeof(float), st));
        if (s == 0) {
            if (step_body()) return 1;                       /* eager: loads every kernel */
        } else {
            if (!graph) {
                CHECK(sn_graph_begin(st));
                if (step_body()) return 1;                   /* recorded, not run */
                CHECK(sn_graph_end(st, &graph));
            }
            CHECK(sn_graph_launch(graph, st));
        }
        float m[2];
        CHECK(sn_memcpy_d2h(m, metrics, sizeof m, st));
        CHECK(sn_stream_sync(st));
        printf("step %d loss %.9f acc %.9f\n", s, m[0] / N, m[1] / N);
    }
    float cw[F * K * K], cb[F];
    CHECK(sn_memcpy_d2h(cw, arena_w + o_cw, sizeof cw, st));
    CHECK(sn_memcpy_d2h(cb, arena_w + o_cb, sizeof cb, st));
    CHECK(sn_stream_sync(st));
    printf("conv_w");
    for (int i = 0; i < F * K * K; ++i) printf(" %.9g", cw[i]);
    printf("\nconv_b");
    for (int i = 0; i < F; ++i) printf(" %.9g", cb[i]);
    printf("\nlaunches %lld\n", sn_launch_count());
    CHECK(sn_graph_destroy(graph));
    CHECK(sn_stream_destroy(st));
    return 0;
}
