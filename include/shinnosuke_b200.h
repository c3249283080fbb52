/*
 * shinnosuke_b200 — C ABI of the B200-native training hot path.
 *
 * The reference (E1eveNn/shinnosuke) is pure Python + NumPy and has no FFI; the
 * "interface each entry point replaces" is therefore the NumPy code of the
 * reference method cited next to it (paths relative to the reference tree).
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a non-zero code otherwise; the
 *     message is available from sn_last_error() (thread local).  Nothing throws
 *     or aborts.
 *   - all work is asynchronous on the caller supplied stream (a cudaStream_t
 *     passed as void*; NULL = the legacy default stream).  No hidden host sync.
 *   - device tensors are plain pointers to float (fp32) unless stated.
 *   - activation maps are NHWC on the device ("pixel rows x channel columns"):
 *     x[n][h][w][c].  Convolution filters are [F][kh][kw][C] ("KRSC"), so the
 *     contraction index is ordered (u, v, c).  Dense layers are the H=W=kh=kw=1
 *     case: X[batch][n_in], W[n_out][n_in].  sn_nchw_to_nhwc / sn_nhwc_to_nchw
 *     convert at the API boundary, where the reference's NCHW order is visible.
 *   - "accumulate" arguments: 0 = overwrite the destination, 1 = add into it
 *     (the reference's  `grads += ...`).
 */
#ifndef SHINNOSUKE_B200_H
#define SHINNOSUKE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SN_ABI_VERSION 1

#if defined(__GNUC__)
#define SN_API __attribute__((visibility("default")))
#else
#define SN_API
#endif

/* error codes (CUDA runtime errors are returned as their positive cudaError_t) */
#define SN_OK                 0
#define SN_ERR_INVALID_ARG   -1
#define SN_ERR_UNSUPPORTED   -2
#define SN_ERR_NO_DEVICE     -3

/* activation fused into a forward epilogue (utils/Activator.py) */
#define SN_ACT_LINEAR 0
#define SN_ACT_RELU   1

/* arithmetic tier of the contraction kernels */
#define SN_PREC_FP32 0   /* CUDA-core FFMA, fp32 operands and accumulators (1e-5 tier)      */
#define SN_PREC_TF32 1   /* tcgen05.mma kind::tf32, fp32 storage, fp32 TMEM accumulators    */
#define SN_PREC_BF16 2   /* tcgen05.mma kind::f16 on bf16 copies of the operands: the *_bf16
                            entry points below; the fp32-pointer entry points treat it as TF32 */

/* max-pool backward semantics (layers/Convolution.py:266-269) */
#define SN_POOL_TIE_SPLIT    0   /* 'reshape' mode: ties share the gradient equally         */
#define SN_POOL_FIRST_ARGMAX 1   /* 'im2col'  mode: first maximum in (u,v) order wins       */

/* ---------------------------------------------------------------- runtime */
SN_API int         sn_abi_version(void);
SN_API const char *sn_last_error(void);
SN_API int sn_device_count(int *count);
SN_API int sn_set_device(int device);
SN_API int sn_device_props(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *global_mem);
SN_API int sn_malloc(void **ptr, size_t bytes);
SN_API int sn_free(void *ptr);
SN_API int sn_malloc_host(void **ptr, size_t bytes);          /* pinned */
SN_API int sn_free_host(void *ptr);
/* page-lock / unlock caller-owned host memory in place (so that sn_memcpy_h2d from it is truly
 * asynchronous without a staging copy) */
SN_API int sn_host_register(void *ptr, size_t bytes);
SN_API int sn_host_unregister(void *ptr);
SN_API int sn_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
SN_API int sn_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
SN_API int sn_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
SN_API int sn_memset(void *dst, int byte, size_t bytes, void *stream);
SN_API int sn_stream_create(void **stream);
SN_API int sn_stream_destroy(void *stream);
SN_API int sn_stream_sync(void *stream);
/* CUDA-event timing on the launching stream (bench.py's per-kernel clock) */
SN_API int sn_event_create(void **event);
SN_API int sn_event_destroy(void *event);
SN_API int sn_event_record(void *event, void *stream);
SN_API int sn_event_elapsed_ms(void *start, void *stop, float *ms);   /* synchronises on `stop` */
/* number of kernels this library launched since the last reset (bench "gpu_launches") */
SN_API long long sn_launch_count(void);
SN_API void      sn_launch_count_reset(void);
/* name of the kernel the most recent launch of this library started (tests use it to see which kernel family a
 * shape was routed to, e.g. that a ragged-K Dense weight gradient stays on the tensor pipe) */
SN_API const char *sn_last_kernel(void);
/* how many of those were CTA-pair (cta_group::2) launches of the implicit-GEMM kernel: the pair tiles
 * are chosen by shape (SN_IGEMM_PAIR=0 never, =2 whenever eligible), tests use this to see which ran */
SN_API long long sn_igemm_pair_launches(void);
/* override the choice in-process: 0 never, 1 by shape (default), 2 whenever eligible, -1 re-read SN_IGEMM_PAIR */
SN_API int sn_igemm_set_pair_mode(int mode);

/* ------------------------------------------------------------ CUDA graphs */
/* The training step (reference models.py:70-92: forward sweep, loss, backward sweep, update) is ~36
 * short kernels; it is captured ONCE per batch shape between sn_graph_begin / sn_graph_end on a
 * stream created by sn_stream_create (every sn_* call on that stream in between is recorded, nothing
 * runs) and replayed with sn_graph_launch.  Capture is thread-local: other host threads may keep
 * using the device. */
SN_API int sn_graph_begin(void *stream);
SN_API int sn_graph_end(void *stream, void **graph_exec);
SN_API int sn_graph_launch(void *graph_exec, void *stream);
SN_API int sn_graph_destroy(void *graph_exec);
SN_API int sn_stream_wait_event(void *stream, void *event);
SN_API int sn_device_sync(void);

/* ------------------------------------------------- data-parallel exchange */
/* models.py fit() sharded over the GPUs of one NVLink / NVSwitch box: every replica
 * runs forward + backward on its rows of the minibatch; the gradient arena (and the per-channel sums of
 * every BatchNormalization, layers/Normalization.py:116-128) are summed over the replicas by ONE
 * hand-written kernel each, over peer-mapped memory (csrc/comm.cu) — no NCCL on the data path.
 * Results are bit-identical on every rank (fixed summation order).
 *
 * A communicator owns a "window" of device memory that every peer maps.  user_bytes of it are the
 * caller's: a tensor that lives there (the gradient arena) is all-reduced in place, zero-copy.
 *   one process per GPU : sn_comm_create on each rank, exchange the 64-byte handles of
 *                         sn_comm_ipc_handle through any host channel, sn_comm_connect(all handles);
 *   one process, G GPUs : sn_comm_init_all(G, user_bytes, comms[G]) (device r = rank r).
 * At most 8 ranks.  All ranks must issue the same sequence of sn_allreduce_* calls (same n). */
SN_API int sn_comm_create(int rank, int world, size_t user_bytes, void **comm);
SN_API int sn_comm_window(void *comm, void **ptr, size_t *bytes);
SN_API int sn_comm_ipc_handle(void *comm, void *handle64);
SN_API int sn_comm_connect(void *comm, const void *handles /* world x 64 bytes, rank order */);
SN_API int sn_comm_init_all(int ndev, size_t user_bytes, void **comms);
SN_API int sn_comm_destroy(void *comm);
/* Waits inside the exchange kernels give up after 20 s (a replica that died must not hang the others' GPUs): the
 * number of such give-ups of this communicator since the last reset; non-zero means the results are invalid.
 * Synchronises the device. */
SN_API int sn_comm_timeouts(void *comm, long long *count, int reset);
/* buf (+)= sum over ranks, in place, asynchronous on `stream` (graph-capturable).  Tensors inside the
 * window larger than 64 KB: two-shot (pull own slice from every peer, push the sum to every peer);
 * tensors of at most 256 KB anywhere: one-shot through per-rank staging slots; anything else is
 * staged through the window. */
SN_API int sn_allreduce_sum_f32(void *comm, float *buf, long long n, void *stream);
SN_API int sn_allreduce_sum_f64(void *comm, double *buf, long long n, void *stream);
/* The data-parallel tail of a training step in ONE kernel: g[0, n) = sum over ranks (as above, g inside the
 * window), then w[i] -= lr * g[i]; g[i] = 0 for i < n_params (StochasticGradientDescent.update,
 * utils/Optimizers.py:29-32) applied by each CTA to the elements whose sums it has just seen arrive.  The
 * words g[n_params, n) (the step's loss / accuracy sums) are reduced and left in place. */
SN_API int sn_allreduce_sgd_update(void *comm, float *g, float *w, long long n, long long n_params, float lr, void *stream);

/* ----------------------------------------------------------------- layout */
/* API-boundary transposes between the reference's NCHW and the device's NHWC. */
SN_API int sn_nchw_to_nhwc(const float *src, float *dst, int N, int C, int H, int W, void *stream);
SN_API int sn_nhwc_to_nchw(const float *src, float *dst, int N, int C, int H, int W, int accumulate, void *stream);
SN_API int sn_cast_f64_to_f32(const double *src, float *dst, long long n, void *stream);
SN_API int sn_cast_f32_to_f64(const float *src, double *dst, long long n, void *stream);
SN_API int sn_axpy(float alpha, const float *x, float *y, long long n, void *stream);     /* y += alpha*x */
/* Add.forward, layers/Base.py:378-385: out = a + b (same shapes; residual connections) */
SN_API int sn_add(const float *a, const float *b, float *out, long long n, void *stream);
/* out = relu(a + b), blocked elements written as -0.0f: Add followed by Activation('relu') (config 4's residual join) */
SN_API int sn_add_relu(const float *a, const float *b, float *out, long long n, void *stream);

/* ---------------------------------------------------------------- Conv2D */
/* Conv2D.forward, layers/Convolution.py:131-161 (true zero padding):
 *   y[n,i,j,f] = act(b[f] + sum_{u,v,c} w[f,u,v,c] * x[n, i*s+u-ph, j*s+v-pw, c])
 * With act == SN_ACT_RELU negative pre-activations are stored as -0.0f so the
 * backward mask "pre-activation < 0" (utils/Activator.py:25) survives in y. */
SN_API int sn_conv2d_fprop(const float *x, const float *w, const float *bias, float *y,
                    int N, int C, int H, int W, int F, int kh, int kw,
                    int stride, int pad_h, int pad_w, int act, int precision, void *stream);
/* sn_conv2d_fprop that also leaves the per-channel sums of its output (sum y, sum y^2: the batch
 * statistics a following BatchNormalization needs, layers/Normalization.py:116-118) in
 * `bn_scratch` (>= 20*F doubles), saving that layer its own read of y.  The sums are ADDED: the
 * scratch must be all zero on entry, which is how cudaMemset or any of the sn_batchnorm_* /
 * sn_bn_pool_* calls on a C = 4*2^k map leave it (their finalise step re-zeroes what it consumed).
 * Only the tensor-core kernels do this: ask sn_conv2d_fprop_stats_supported (bf16_operands = 1
 * for sn_conv2d_fprop_bf16_stats) first.  Follow with sn_batchnorm_finalize_stats. */
SN_API int sn_conv2d_fprop_stats_supported(int precision, int bf16_operands, int N, int C, int H, int W, int F,
                                    int kh, int kw, int stride, int pad_h, int pad_w);
SN_API int sn_conv2d_fprop_stats(const float *x, const float *w, const float *bias, float *y,
                          int N, int C, int H, int W, int F, int kh, int kw,
                          int stride, int pad_h, int pad_w, int act, int precision, double *bn_scratch, void *stream);
/* Conv2D.backward dX, layers/Convolution.py:199-202 + utils/ConvCol.py:24-42 (col2im-free).
 * The tensor-core tiers run dgrad as a forward convolution of dY with the filters flipped and
 * transposed; `workspace` holds that copy (sn_conv2d_dgrad_workspace() bytes, device memory,
 * rewritten by every call).  workspace == NULL selects the CUDA-core kernel. */
SN_API size_t sn_conv2d_dgrad_workspace(int C, int F, int kh, int kw, int precision);
SN_API int sn_conv2d_dgrad(const float *dy, const float *w, float *dx,
                    int N, int C, int H, int W, int F, int kh, int kw,
                    int stride, int pad_h, int pad_w, int accumulate, int precision,
                    void *workspace, void *stream);
/* Conv2D.backward dW (+= dYr . cols^T, :197) and db (+= sum dY, :192); db may be NULL */
SN_API int sn_conv2d_wgrad(const float *x, const float *dy, float *dw, float *db,
                    int N, int C, int H, int W, int F, int kh, int kw,
                    int stride, int pad_h, int pad_w, int accumulate, int precision, void *stream);
/* Relu._backward, utils/Activator.py:24-28: dy[i] = 0 where the pre-activation was < 0
 * (sign bit of y, see sn_conv2d_fprop); in place. */
SN_API int sn_relu_bwd(float *dy, const float *y, long long n, void *stream);
/* stand-alone Relu forward with the -0.0f convention (Activation layer, layers/Activation.py) */
SN_API int sn_relu_fwd(const float *x, float *y, long long n, void *stream);

/* ZeroPadding2D.forward / backward, layers/Convolution.py:232-251: y (N,H+2ph,W+2pw,C) is x
 * surrounded by zeros; backward crops dy to dx (N,H,W,C). */
SN_API int sn_pad2d_fwd(const float *x, float *y, int N, int H, int W, int C, int pad_h, int pad_w, void *stream);
SN_API int sn_pad2d_bwd(const float *dy, float *dx, int N, int H, int W, int C, int pad_h, int pad_w,
                 int accumulate, void *stream);

/* 3xTF32: a near-fp32 tier on the tensor pipe (SURVEY.md 8(c) "FFMA or 3xTF32").
 * x*w = xh*wh + xh*wl + xl*wh + O(2^-22) with xh = the 10-mantissa-bit truncation the tensor core
 * applies to an fp32 operand anyway and xl = x - xh rounded to 10 bits (sn_split_tf32_lo).  Each entry
 * point runs the tf32 kernel three times, accumulating in fp32.  Every operator is within 1e-5 of
 * float64; because the tensor core's fp32 accumulation is not FFMA's round-to-nearest, training
 * trajectories drift to ~3e-5, so this is precision='tf32x3', not the 'fp32' tier.  op: 0 = fprop,
 * 1 = dgrad, 2 = wgrad; unsupported shapes (sn_conv2d_x3_supported == 0) use the fp32-pointer entry
 * points with SN_PREC_FP32.  dgrad workspace: 2*C*F*kh*kw floats. */
SN_API int sn_split_tf32_lo(const float *src, float *dst_lo, long long n, void *stream);
SN_API int sn_conv2d_x3_supported(int op, int N, int C, int H, int W, int F, int kh, int kw,
                           int stride, int pad_h, int pad_w);
SN_API int sn_conv2d_fprop_x3(const float *x, const float *x_lo, const float *w, const float *w_lo, const float *bias,
                       float *y, int N, int C, int H, int W, int F, int kh, int kw,
                       int stride, int pad_h, int pad_w, int act, void *stream);
SN_API int sn_conv2d_dgrad_x3(const float *dy, const float *dy_lo, const float *w, float *dx,
                       int N, int C, int H, int W, int F, int kh, int kw,
                       int stride, int pad_h, int pad_w, int accumulate, void *workspace, void *stream);
SN_API int sn_conv2d_wgrad_x3(const float *x, const float *x_lo, const float *dy, const float *dy_lo, float *dw, float *db,
                       int N, int C, int H, int W, int F, int kh, int kw,
                       int stride, int pad_h, int pad_w, void *stream);

/* bf16 tier (north_star: "bf16 inputs and fp32 accumulation", 1e-2 against float64).  The
 * operands are bf16 COPIES made with sn_cast_f32_to_bf16 (x, dY) / sn_flip_transpose_filters_bf16
 * (filters for dgrad); accumulators, bias, activation, outputs and weight gradients stay fp32.
 * op: 0 = fprop, 1 = dgrad, 2 = wgrad.  Shapes the tensor-core kernels do not take report 0 from
 * sn_conv2d_bf16_supported and must use the fp32-pointer entry points. */
SN_API int sn_cast_f32_to_bf16(const float *src, void *dst_bf16, long long n, void *stream);
SN_API int sn_flip_transpose_filters_bf16(const float *w, void *wt_bf16, int F, int C, int kh, int kw, void *stream);
/* All bf16 weight operands of a training step in one launch: shadow_bf16[0, n) = bf16(arena[0, n))
 * (Conv2D.forward's filters are slices of it), plus, for each of `ntasks` convolutions, the
 * flipped + transposed dgrad operand of sn_flip_transpose_filters_bf16.  tasks_dev: device array of
 * 8 int64 per task {element offset of W in the arena, device address of Wt, F, C, kh, kw, first
 * unit, units} where a unit is one 32x32 (f, c) tile of one tap: units = kh*kw*ceil(F/32)*ceil(C/32),
 * first unit = running sum; flip_units = total. */
SN_API int sn_conv_weights_prepare_bf16(const float *arena, void *shadow_bf16, long long n, const long long *tasks_dev,
                                 int ntasks, int flip_units, void *stream);
SN_API int sn_conv2d_bf16_supported(int op, int N, int C, int H, int W, int F, int kh, int kw,
                             int stride, int pad_h, int pad_w);
SN_API int sn_conv2d_fprop_bf16(const void *x_bf16, const void *w_bf16, const float *bias, float *y,
                         int N, int C, int H, int W, int F, int kh, int kw,
                         int stride, int pad_h, int pad_w, int act, void *stream);
SN_API int sn_conv2d_fprop_bf16_stats(const void *x_bf16, const void *w_bf16, const float *bias, float *y,
                               int N, int C, int H, int W, int F, int kh, int kw,
                               int stride, int pad_h, int pad_w, int act, double *bn_scratch, void *stream);
/* wt_bf16: filters flipped and transposed by sn_flip_transpose_filters_bf16 */
/* bf16 tier, a convolution whose only consumer is a fused BatchNormalization (+ 2x2 pool): the output map
 * itself is stored in bf16 (NHWC, written once and read twice: half the bytes three times) and its
 * per-channel sums — of the values as stored — are added to bn_scratch like sn_conv2d_fprop_stats does.
 * bf16_operands = 1: x and w are bf16 copies (sn_conv2d_fprop_bf16's kernels); 0: the tiny-C stem kernel
 * on the fp32 image and filters (TF32 products).  F % 8 == 0, 32 < F <= 512. */
SN_API int sn_conv2d_fprop_ybf16_supported(int bf16_operands, int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w);
SN_API int sn_conv2d_fprop_bf16_stats_ybf16(const void *x_bf16, const void *w_bf16, const float *bias, void *y_bf16,
                                     int N, int C, int H, int W, int F, int kh, int kw,
                                     int stride, int pad_h, int pad_w, int act, double *bn_scratch, void *stream);
SN_API int sn_conv2d_fprop_stats_ybf16(const float *x, const float *w, const float *bias, void *y_bf16,
                                int N, int C, int H, int W, int F, int kh, int kw,
                                int stride, int pad_h, int pad_w, int act, double *bn_scratch, void *stream);
/* Tiny-C convolutions (C <= 4, kh*kw <= 9, stride 1: the RGB stem) on the tensor pipe with the column matrix
 * gathered by the TMA unit (cp.async.bulk.tensor im2col, csrc/conv_stem_im2col.cu); tf32 / bf16 tiers.
 * The input must be held as NHWC4 (channels padded to 4): sn_conv2d_stem_prepare writes that copy
 * (N*H*W*4 floats) from the NHWC map; fprop then reads x4.  y is fp32 or (y_bf16 != 0) bf16,
 * bn_scratch (may be NULL) receives the per-channel sums of y like sn_conv2d_fprop_stats. */
SN_API int sn_conv2d_stem_supported(int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w);
SN_API int sn_conv2d_stem_prepare(const float *x, float *x4, int N, int C, int H, int W, void *stream);
SN_API int sn_conv2d_stem_fprop(const float *x4, const float *w, const float *bias, void *y,
                         int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w,
                         int act, int y_bf16, double *bn_scratch, void *stream);
/* The stem's weight gradient in the bf16 tier (Conv2D.backward, layers/Convolution.py:176-196, for C*kh*kw < 32):
 * x is the fp32 NHWC input, dy_bf16 the bf16 copy of the gradient map (what sn_bn_pool_bwd_xbf16 writes, so the
 * fp32 map need not exist); dw += and, when db != NULL, db += column sums of dy (csrc/conv_stem_tc.cu). */
SN_API int sn_conv2d_stem_wgrad_dybf16_supported(int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w);
SN_API int sn_conv2d_stem_wgrad_dybf16(const float *x, const void *dy_bf16, float *dw, float *db,
                                int N, int C, int H, int W, int F, int kh, int kw,
                                int stride, int pad_h, int pad_w, void *stream);
SN_API int sn_conv2d_dgrad_bf16(const void *dy_bf16, const void *wt_bf16, float *dx,
                         int N, int C, int H, int W, int F, int kh, int kw,
                         int stride, int pad_h, int pad_w, int accumulate, void *stream);
/* dw += (always accumulates); db += column sums of the fp32 dY when both are non-NULL */
SN_API int sn_conv2d_wgrad_bf16(const void *x_bf16, const void *dy_bf16, const float *dy_f32, float *dw, float *db,
                         int N, int C, int H, int W, int F, int kh, int kw,
                         int stride, int pad_h, int pad_w, void *stream);

/* ----------------------------------------------------------------- Dense */
/* Dense.forward, layers/FC.py:124-134: y = act(x . W + b), x[M][K], w[Nout][K] */
SN_API int sn_dense_fprop(const float *x, const float *w, const float *bias, float *y,
                   int M, int K, int Nout, int act, int precision, void *stream);
/* Dense.backward dX = dZ . W^T, layers/FC.py:147 */
SN_API int sn_dense_dgrad(const float *dy, const float *w, float *dx,
                   int M, int K, int Nout, int accumulate, int precision,
                   void *workspace /* sn_conv2d_dgrad_workspace(K, Nout, 1, 1, precision) */, void *stream);
/* Dense.backward dW += X^T . dZ (:144), db += sum_rows dZ (:146) */
SN_API int sn_dense_wgrad(const float *x, const float *dy, float *dw, float *db,
                   int M, int K, int Nout, int accumulate, int precision, void *stream);

/* The classifier head: Dense with 1 <= n_out <= 16 (every BASELINE config ends in Dense(10, softmax)),
 * far too thin for GEMM tiles.  fprop = x.W + b with an optional row softmax (layers/FC.py:124-134,
 * utils/Activator.py:108-110) in one launch; bwd = dW += x^T.dZ, db += sum dZ, dX (=|+=) dZ.W^T
 * (layers/FC.py:140-153) in one launch that reads x once.  Any of dw / db / dx may be NULL. */
SN_API int sn_dense_head_supported(int M, int K, int Nout);
SN_API int sn_dense_head_fprop(const float *x, const float *w, const float *bias, float *out,
                        int M, int K, int Nout, int softmax, void *stream);
SN_API int sn_dense_head_bwd(const float *x, const float *dz, const float *w, float *dw, float *db, float *dx,
                      int M, int K, int Nout, int accumulate_dx, void *stream);


/* ---------------------------------------------------------- MaxPooling2D */
/* MaxPooling2D.forward, layers/Convolution.py:287-324.  x NHWC (N,H,W,C) -> y (N,oh,ow,C),
 * oh = (H-pool)/stride+1 (floor).  `code` (one byte per output element, order (n,i,j,c)) is
 *   SN_POOL_TIE_SPLIT   : bit (u*pool+v) set where x == max   (pool*pool <= 8; for 9..16 window
 *                         elements the code is TWO bytes per output element, little endian)
 *   SN_POOL_FIRST_ARGMAX: u*pool+v of the first maximum               (== the reference's
 *                         int64 argmax_index, same order, one byte each)
 * code may be NULL when not training. */
SN_API int sn_maxpool2d_fwd(const float *x, float *y, uint8_t *code,
                     int N, int H, int W, int C, int pool, int stride, int mode, void *stream);
/* MaxPooling2D.backward, layers/Convolution.py:326-362 */
SN_API int sn_maxpool2d_bwd(const float *dy, const uint8_t *code, float *dx,
                     int N, int H, int W, int C, int pool, int stride, int mode,
                     int accumulate, void *stream);

/* --------------------------------------------------------------- Flatten */
/* Flatten.forward/backward, layers/FC.py:39-58: the reference flattens NCHW in (c,i,j)
 * order; on the device that is the NHWC -> NCHW transpose (and back). */
/* (sn_nhwc_to_nchw / sn_nchw_to_nhwc above) */

/* ---------------------------------------------------- BatchNormalization */
/* BatchNormalization.forward (training), layers/Normalization.py:116-128, channel = last
 * (fastest) index of x[rows][C].  Writes y, save_mean[C], save_invstd[C] and updates the moving
 * statistics in place.  scratch: >= 20*C doubles of device memory, zeroed by the call. */
SN_API int sn_batchnorm_fwd_train(const float *x, const float *gamma, const float *beta, float *y,
                           float *save_mean, float *save_invstd, float *moving_mean, float *moving_var,
                           double *scratch, long long rows, int C, float eps, float momentum, void *stream);
/* inference branch, :130-132 */
SN_API int sn_batchnorm_fwd_infer(const float *x, const float *gamma, const float *beta, float *y,
                           const float *moving_mean, const float *moving_var,
                           long long rows, int C, float eps, void *stream);
/* BatchNormalization.backward, :186-230. dgamma/dbeta accumulate (+=).  If relu_y != NULL the
 * ReLU mask of the producing layer (sign bit of relu_y == x) is applied to dx on the way out. */
SN_API int sn_batchnorm_bwd(const float *dy, const float *x, const float *gamma,
                     const float *save_mean, const float *save_invstd,
                     float *dx, float *dgamma, float *dbeta, double *scratch,
                     long long rows, int C, int accumulate_dx, int mask_relu, void *stream);

/* Pass 1 of sn_batchnorm_fwd_train on its own: per-channel batch statistics of x[rows][C] ->
 * save_mean / save_invstd, moving statistics updated, and the apply-pass constants left in
 * `scratch` (>= 20*C doubles, zeroed by the call).  Needs C = 4*2^k <= 1024. */
SN_API int sn_batchnorm_stats(const float *x, const float *gamma, const float *beta, float *save_mean,
                       float *save_invstd, float *moving_mean, float *moving_var, double *scratch,
                       long long rows, int C, float eps, float momentum, void *stream);

/* The two halves of the training forward when the producing convolution already reduced the
 * statistics (sn_conv2d_fprop_stats): finalize turns the sums in `scratch` into save_mean /
 * save_invstd, updates the moving statistics and leaves the apply constants in `scratch`; apply
 * is y = gamma*(x-mean)*invstd + beta.  Both need C = 4*2^k <= 1024. */
SN_API int sn_batchnorm_finalize_stats(const float *gamma, const float *beta, float *save_mean, float *save_invstd,
                                float *moving_mean, float *moving_var, double *scratch,
                                long long rows, int C, float eps, float momentum, void *stream);
/* relu != 0: the Activation('relu') that follows the layer is applied on the way out (blocked elements = -0.0f) */
SN_API int sn_batchnorm_fwd_apply(const float *x, float *y, const double *scratch, long long rows, int C, int relu, void *stream);

/* BatchNormalization immediately followed by MaxPooling2D((2,2)) (stride 2), fused: the normalised
 * map is never written.  x: conv output (N,H,W,C); y_pool: (N,H/2,W/2,C); code: one byte per
 * pooled element (same meaning as sn_maxpool2d_fwd).  Requires sn_bn_pool_supported().
 * y_pool_bf16 (may be NULL): also write the bf16 copy the next convolution of the bf16 tier reads. */
SN_API int sn_bn_pool_supported(int N, int H, int W, int C);
SN_API int sn_bn_pool_fwd_train(const float *x, const float *gamma, const float *beta, float *y_pool, void *y_pool_bf16, uint8_t *code,
                         float *save_mean, float *save_invstd, float *moving_mean, float *moving_var,
                         double *scratch, int N, int H, int W, int C, float eps, float momentum,
                         int pool_mode, void *stream);
/* pass 2 of sn_bn_pool_fwd_train alone (constants already in scratch: sn_batchnorm_stats or
 * sn_batchnorm_finalize_stats ran) */
SN_API int sn_bn_pool_fwd_apply(const float *x, float *y_pool, void *y_pool_bf16, uint8_t *code, const double *scratch,
                         int N, int H, int W, int C, int pool_mode, void *stream);
/* backward of the fused pair: dy_pool (N,H/2,W/2,C) -> dx (N,H,W,C) (overwritten), dgamma/dbeta +=.
 * y_pool / beta (both or neither; may be NULL): the forward's pooled output and the shift it was
 * computed with.  With them the two reductions of BatchNormalization.backward (:196-203) read the
 * pooled tensors only, since every pool winner satisfies xhat = (y_pool - beta)/gamma.
 * Outputs for the producing Conv2D's backward (layers/Convolution.py:186-197), each optional but
 * at least one of dx / dx_bf16: dx (fp32), dx_bf16 (the bf16 operand copy of the same values),
 * dx_colsum[C] += sum over rows of dx (that layer's bias gradient, :192; needs scratch >= 28*C doubles whose last 8*C are zero on first use and left zero). */
SN_API int sn_bn_pool_bwd(const float *dy_pool, const uint8_t *code, const float *x, const float *y_pool,
                   const float *gamma, const float *beta, const float *save_mean, const float *save_invstd,
                   float *dx, void *dx_bf16, float *dx_colsum, float *dgamma, float *dbeta, double *scratch,
                   int N, int H, int W, int C, int pool_mode, int mask_relu, void *stream);

/* Cross-replica BatchNormalization (SyncBN) under data parallelism: the statistics of
 * layers/Normalization.py:116-118 and the two sums of :196-203 are over the WHOLE minibatch, which
 * is sharded.  Each direction runs in halves around an all-reduce the caller performs on
 * scratch[0, 16*C) (float64): *_local leaves this replica's sums there (dgamma / dbeta, which are
 * per-replica contributions, are accumulated at once), the caller all-reduces, *_finalize_* /
 * sn_batchnorm_finalize_stats turns the global sums into the apply constants with the GLOBAL row
 * count, *_apply writes the result.  C = 4*2^k <= 1024. */
SN_API int sn_batchnorm_local_stats(const float *x, double *scratch, long long rows, int C, void *stream);
SN_API int sn_batchnorm_bwd_local(const float *dy, const float *x, const float *save_mean, const float *save_invstd,
                           float *dgamma, float *dbeta, double *scratch, long long rows, int C, void *stream);
SN_API int sn_batchnorm_finalize_bwd(const float *gamma, const float *save_mean, const float *save_invstd,
                              double *scratch, long long rows, int C, void *stream);
/* Data-parallel finalize (exact BatchNormalization over the WHOLE minibatch, layers/Normalization.py:116-128,
 * with the minibatch sharded over the replicas of `comm`): ONE kernel folds the local sums in scratch,
 * exchanges them with every replica over peer-mapped memory, adds them in rank order and finalises with
 * the global row count — instead of sn_allreduce_sum_f64 + sn_batchnorm_finalize_*.  Same scratch
 * contract as the local versions.  All replicas must call it in the same order. */
SN_API int sn_batchnorm_finalize_stats_dp(void *comm, const float *gamma, const float *beta, float *save_mean, float *save_invstd,
                                   float *moving_mean, float *moving_var, double *scratch, long long rows_global, int C,
                                   float eps, float momentum, void *stream);
SN_API int sn_batchnorm_finalize_bwd_dp(void *comm, const float *gamma, const float *save_mean, const float *save_invstd,
                                 double *scratch, long long rows_global, int C, void *stream);
/* The exchange split in two, so that independent work can run between the halves and absorb both the NVLink
 * latency and the skew between the replicas (the backward pass puts the next layer's weight-gradient kernel
 * there): _push folds the local sums and sends them, _pushed polls, adds and finalises.  No other exchange on
 * this communicator may be issued between the two. */
SN_API int sn_batchnorm_dp_push(void *comm, double *scratch, int C, void *stream);
SN_API int sn_batchnorm_finalize_bwd_dp_pushed(void *comm, const float *gamma, const float *save_mean, const float *save_invstd,
                                        double *scratch, long long rows_global, int C, void *stream);
SN_API int sn_batchnorm_bwd_apply(const float *dy, const float *x, float *dx, const double *scratch,
                           long long rows, int C, int accumulate_dx, int mask_relu, void *stream);
SN_API int sn_bn_pool_bwd_local(const float *dy_pool, const uint8_t *code, const float *x, const float *y_pool,
                         const float *gamma, const float *beta, const float *save_mean, const float *save_invstd,
                         float *dgamma, float *dbeta, double *scratch, int N, int H, int W, int C,
                         int pool_mode, void *stream);
SN_API int sn_bn_pool_bwd_apply(const float *dy_pool, const uint8_t *code, const float *x, float *dx, void *dx_bf16,
                         float *dx_colsum, double *scratch, int N, int H, int W, int C,
                         int pool_mode, int mask_relu, void *stream);
/* the same kernels reading the full-resolution map x as bf16 (the *_ybf16 convolution outputs above);
 * everything else — pooled tensors, gradients, constants — stays fp32, a blocked ReLU element is -0.0 */
SN_API int sn_bn_pool_fwd_apply_xbf16(const void *x_bf16, float *y_pool, void *y_pool_bf16, uint8_t *code, const double *scratch,
                               int N, int H, int W, int C, int pool_mode, void *stream);
SN_API int sn_bn_pool_bwd_xbf16(const float *dy_pool, const uint8_t *code, const void *x_bf16, const float *y_pool, const float *gamma,
                         const float *beta, const float *save_mean, const float *save_invstd, float *dx, void *dx_bf16,
                         float *dx_colsum, float *dgamma, float *dbeta, double *scratch, int N, int H, int W, int C,
                         int pool_mode, int mask_relu, void *stream);
SN_API int sn_bn_pool_bwd_local_xbf16(const float *dy_pool, const uint8_t *code, const void *x_bf16, const float *y_pool,
                               const float *gamma, const float *beta, const float *save_mean, const float *save_invstd,
                               float *dgamma, float *dbeta, double *scratch, int N, int H, int W, int C, int pool_mode, void *stream);
SN_API int sn_bn_pool_bwd_apply_xbf16(const float *dy_pool, const uint8_t *code, const void *x_bf16, float *dx, void *dx_bf16,
                               float *dx_colsum, double *scratch, int N, int H, int W, int C, int pool_mode, int mask_relu, void *stream);

/* ------------------------------------------------------ Embedding / LSTM */
/* Embedding.forward, layers/Embeddings.py:56-63: out[i][:] = W[idx[i]][:].  Indices arrive as fp32
 * VALUES (the minibatch pipeline moves float32; exact below 2^24); out-of-range indices are clamped. */
SN_API int sn_embedding_fwd(const float *w, const float *idx, float *out, long long n_idx, int vocab, int dim, void *stream);
/* Embedding.backward, :66-76: np.add.at(W.grads, idx, grads) — dw[idx[i]][:] += dy[i][:] */
SN_API int sn_embedding_bwd(const float *idx, const float *dy, float *dw, long long n_idx, int vocab, int dim, void *stream);
/* ids follow NumPy's W[idx]: [-vocab, 0) wraps; ids outside [-vocab, vocab) (IndexError in the reference,
 * layers/Embeddings.py:60) are clamped and COUNTED on the device; this reads (and optionally clears) the count.
 * Synchronises the device. */
SN_API int sn_embedding_oob_count(long long *count, int reset);
/* dst[b][a][:] (=|+=) src[a][b][:] for rows of D floats: batch-major (N,T,D) <-> time-major (T,N,D) */
SN_API int sn_swap_leading_axes(const float *src, float *dst, int A, int B, int D, int accumulate, void *stream);
/* One timestep of LSTMCell._forward, layers/Recurrent.py:293-309.  zx[r][0..4H) = x_t . Wx + b (row
 * stride zx_stride floats; the input half of the reference's fused [h, x] . W product, computed for
 * all timesteps by one Dense GEMM), wh[4H][H] = the recurrent rows of W transposed, gate columns
 * (f, u, c~, o).  Writes the gate activations acts[N][4H], c[N][H], h[N][H] and, when h_seq != NULL,
 * h again at h_seq[r * hseq_stride + j] (the (N,T,H) output of return_sequences=True).
 * h_prev / c_prev may be NULL (zero initial state).  sigmoid / tanh as utils/Activator.py:58,82. */
SN_API int sn_lstm_step_fwd(const float *zx, long long zx_stride, const float *wh, const float *h_prev, const float *c_prev,
                     float *acts, float *c, float *h, float *h_seq, long long hseq_stride, int N, int H, void *stream);
/* One timestep of LSTMCell._backward, :326-366: dh = dh_ext[r * dh_stride + j] (upstream gradient of
 * this step, may be NULL) + dz_next . Wh^T (dz of step t+1, NULL at the last step); dz[N][4H] = the
 * gate pre-activation gradients in the FORWARD column order (f, u, c~, o) — the reference
 * concatenates them as (o, c~, u, f), which does not match its own W; dc[N][H]
 * holds dc_{t+1} on entry (ignored when first != 0) and dc_t * f on return. */
SN_API int sn_lstm_step_bwd(const float *dh_ext, long long dh_stride, const float *dz_next, const float *wh, const float *acts,
                     const float *c, const float *c_prev, float *dc, float *dz, int N, int H, int first, void *stream);

/* The whole recurrence in ONE launch (H = 64 or 128): clusters of H/32 CTAs own 8 batch rows, keep
 * their slice of Wh in registers for all T steps and exchange h_t (forward) / dz_t (backward) through
 * distributed shared memory, one cluster barrier per step.  Time-major tensors: zx, acts, dz
 * [T][N][4H]; c_all, h_all [T+1][N][H] (slot 0 = initial state, read only through h0 / c0, which may
 * be NULL for zeros); h_seq (N,T,H) optional.  Backward: dh_seq (N,T,H) or dh_last (N,H). */
SN_API int sn_lstm_seq_supported(int N, int H);
SN_API int sn_lstm_seq_fwd(const float *zx, const float *wh, const float *h0, const float *c0, float *acts, float *c_all,
                    float *h_all, float *h_seq, int T, int N, int H, void *stream);
SN_API int sn_lstm_seq_bwd(const float *dh_last, const float *dh_seq, const float *wh, const float *acts, const float *c_all,
                    int c0_is_zero, float *dz, int T, int N, int H, void *stream);
/* The same recurrence (layers/Recurrent.py:293-309 forward, :326-366 backward) with the per-step product
 * h_{t-1}.Wh^T / dz_{t+1}.Wh on the tensor pipe (H = 128, bf16 tier: bf16 operands, fp32 accumulation and
 * gate arithmetic): Wh resident in shared memory as the K-major UMMA A operand, 8 batch rows per CTA as the
 * B operand, D transposed in tensor memory (lane = hidden unit), csrc/recurrent_tc.cu.  Same arguments. */
SN_API int sn_lstm_seq_tc_supported(int N, int H);
SN_API int sn_lstm_seq_fwd_tc(const float *zx, const float *wh, const float *h0, const float *c0, float *acts, float *c_all,
                       float *h_all, float *h_seq, int T, int N, int H, void *stream);
SN_API int sn_lstm_seq_bwd_tc(const float *dh_last, const float *dh_seq, const float *wh, const float *acts, const float *c_all,
                       int c0_is_zero, float *dz, int T, int N, int H, void *stream);

/* ------------------------------------------------------------------ loss */
/* Softmax._forward, utils/Activator.py:108-110 (row-wise; the reference's global shift cancels) */
SN_API int sn_softmax_fwd(const float *logits, float *probs, int rows, int classes, void *stream);
/* SparseCategoricalCrossEntropy.backward / calc_loss / calc_acc, utils/Objectives.py:72-91
 * (one-hot labels).  dlogits = (probs - onehot)/denom_rows, where denom_rows is the reference's
 * N = prod(shape[:-1]) — the GLOBAL minibatch rows when `rows` is one data-parallel shard of it
 * (pass 0 for denom_rows == rows);  metrics[0] += sum(-y log p), metrics[1] += #(argmax p ==
 * argmax y).  metrics must be zeroed by the caller; may be NULL.  A probability that underflowed fp32
 * to 0 enters the loss as FLT_MIN (87.3 per sample) instead of log(0). */
SN_API int sn_scce_bwd(const float *probs, const float *onehot, float *dlogits, float *metrics,
                int rows, int classes, int denom_rows, void *stream);

/* ------------------------------------------------------------- optimizer */
/* StochasticGradientDescent.update, utils/Optimizers.py:29-32 over a flat arena:
 *   w -= lr * grad_scale * g ;  g = 0 */
SN_API int sn_sgd_update(float *w, float *g, long long n, float lr, float grad_scale, void *stream);
/* Momentum.update, utils/Optimizers.py:46-55:  v = rho*v + (1-rho)*g ; w -= lr*v ; g kept.
 * If gstep != NULL (data-parallel): g += grad_scale*gstep ; gstep = 0 first. */
SN_API int sn_momentum_update(float *w, float *g, float *v, float *gstep, long long n,
                       float lr, float rho, float grad_scale, void *stream);
/* RMSprop (kind 0), AdaGrad (1), AdaDelta (2), Adam (3): utils/Optimizers.py:61-157 over a flat
 * arena.  s1 / s2: state buffers of n floats (s2 for AdaDelta and Adam), zero before the first
 * step.  p1, p2 = rho / (beta1, beta2).  None of them zeroes g (as in the reference); gstep as in
 * sn_momentum_update.  tick (Adam only): 3 floats of device memory, zero before the first step,
 * holding the reference's iteration counter (which advances by two per update) and the two bias
 * corrections, so that a captured graph of the step stays correct when replayed. */
SN_API int sn_adaptive_update(int kind, float *w, float *g, float *s1, float *s2, float *gstep, float *tick,
                       long long n, float lr, float p1, float p2, float eps, float grad_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SHINNOSUKE_B200_H */
