// extern "C" entry points of Conv2D and Dense; picks the kernel family for the requested tier.
#include "common.cuh"
#include <stdlib.h>
using namespace sn;

namespace sn {
int conv_fprop_simt(const float *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv_dgrad_simt(const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv_wgrad_simt(const float *, const float *, float *, float *, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
bool conv_tc_eligible(int N, int C, int OH, int OW, int F, int stride, bool bf16 = false);
bool wgrad_tc_eligible(int N, int C, int OH, int OW, int F, int stride);
int conv_igemm_tc(const void *x, const void *w, const float *bias, void *out, int N, int C, int IH, int IW, int F, int kh,
                  int kw, int OH, int OW, int pad_h, int pad_w, int act, int accumulate, bool bf16, double *stats, cudaStream_t st,
                  bool out_bf16 = false);
int flip_transpose_filters_bf16(const float *w, void *wt, int F, int C, int kh, int kw, cudaStream_t st);
int cast_f32_to_bf16(const float *src, void *dst, long long n, cudaStream_t st);
int split_tf32_lo(const float *src, float *lo, long long n, cudaStream_t st);
int prepare_weights_bf16(const float *arena, void *shadow, long long n, const long long *tasks, int ntasks, int flip_units, cudaStream_t st);
int conv_wgrad_tc(const float *x, const float *dy, float *dw, int N, int C, int IH, int IW, int F, int kh, int kw, int OH,
                  int OW, int pad_h, int pad_w, cudaStream_t st);
int flip_transpose_filters(const float *w, float *wt, int F, int C, int kh, int kw, cudaStream_t st);
bool wgrad_halo_eligible(int N, int C, int OH, int OW, int F, int kh, int kw, int stride, bool bf16 = false);
int conv_wgrad_halo_tc(const void *x, const void *dy, float *dw, int N, int C, int IH, int IW, int F, int kh, int kw, int OH,
                       int OW, int pad_h, int pad_w, bool bf16, cudaStream_t st);
int colsum_accumulate(const float *dy, float *db, long long rows, int F, cudaStream_t st);
bool stem_tc_eligible(int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw);
int conv_stem_fprop_tc(const float *x, const float *w, const float *bias, void *y, int N, int C, int H, int W, int F, int kh,
                       int kw, int stride, int ph, int pw, int act, double *stats, cudaStream_t st, bool out_bf16 = false);
int conv_stem_wgrad_tc(const float *x, const void *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh, int kw,
                       int stride, int ph, int pw, cudaStream_t st, bool dy_bf16 = false);
bool stem_im2col_eligible(int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw);
int pad_channels4(const float *x, float *x4, long long npix, int C, cudaStream_t st);
int conv_stem_im2col_fprop(const float *x4, const float *w, const float *bias, void *y, int N, int C, int H, int W, int F, int kh, int kw,
                           int stride, int ph, int pw, int act, double *stats, cudaStream_t st, bool out_bf16);
bool tiny_wgrad_eligible(int C, int kh, int kw, int F);
int conv_wgrad_tiny(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                    int kw, int stride, int ph, int pw, cudaStream_t st);
bool direct_fprop_eligible(int C, int kh, int kw, int F);
bool direct_wgrad_eligible(int C, int kh, int kw, int F);
int conv_fprop_direct(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                      int kh, int kw, int stride, int ph, int pw, int act, cudaStream_t st);
int conv_wgrad_direct(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                      int kw, int stride, int ph, int pw, cudaStream_t st);
}

static int check_prec(int precision) {
    if (precision != SN_PREC_FP32 && precision != SN_PREC_TF32 && precision != SN_PREC_BF16) {
        set_error("unknown precision tier %d", precision);
        return SN_ERR_INVALID_ARG;
    }
    return SN_OK;
}

// Which forward kernels can also produce the BatchNorm sums of their output (TMA-store epilogues)
static bool igemm_stats_ok(int F) { return F > 16 && F <= 512 && (F & 3) == 0; }

static int fprop_impl(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                      int kh, int kw, int stride, int pad_h, int pad_w, int act, int precision, double *stats, void *stream) {
    SN_REQUIRE(x && w && y, "null tensor");
    SN_REQUIRE(act == SN_ACT_LINEAR || act == SN_ACT_RELU, "unknown activation");
    int e = check_prec(precision);
    if (e) return e;
    const bool geom_ok = N > 0 && C > 0 && F > 0 && kh > 0 && kw > 0 && stride > 0 && pad_h >= 0 && pad_w >= 0 &&
                         H + 2 * pad_h >= kh && W + 2 * pad_w >= kw;
    if (geom_ok && precision != SN_PREC_FP32 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)bias) & 15) == 0 &&
        stem_tc_eligible(N, C, H, W, F, kh, kw, stride, pad_h, pad_w))
        return conv_stem_fprop_tc(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, stats, as_stream(stream));
    if (!stats && geom_ok && (((uintptr_t)y) & 15) == 0 && direct_fprop_eligible(C, kh, kw, F))
        return conv_fprop_direct(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, as_stream(stream));
    if (geom_ok && precision != SN_PREC_FP32) {
        const int oh = (H + 2 * pad_h - kh) / stride + 1, ow = (W + 2 * pad_w - kw) / stride + 1;
        if (conv_tc_eligible(N, C, oh, ow, F, stride))
            return conv_igemm_tc(x, w, bias, y, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, act, 0, false, stats, as_stream(stream));
    }
    SN_REQUIRE(!stats, "no BatchNorm statistics from this kernel family (ask sn_conv2d_fprop_stats_supported first)");
    return conv_fprop_simt(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, as_stream(stream));
}

extern "C" {

int sn_conv2d_fprop(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                    int kh, int kw, int stride, int pad_h, int pad_w, int act, int precision, void *stream) {
    return fprop_impl(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, precision, nullptr, stream);
}

int sn_conv2d_fprop_stats(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                          int kh, int kw, int stride, int pad_h, int pad_w, int act, int precision, double *bn_scratch,
                          void *stream) {
    SN_REQUIRE(bn_scratch, "null statistics buffer");
    return fprop_impl(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, precision, bn_scratch, stream);
}

int sn_conv2d_fprop_stats_supported(int precision, int bf16_operands, int N, int C, int H, int W, int F, int kh, int kw,
                                    int stride, int pad_h, int pad_w) {
    if (N <= 0 || C <= 0 || F <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad_h < 0 || pad_w < 0 || H + 2 * pad_h < kh ||
        W + 2 * pad_w < kw)
        return 0;
    if (bf16_operands) return (sn_conv2d_bf16_supported(0, N, C, H, W, F, kh, kw, stride, pad_h, pad_w) && igemm_stats_ok(F)) ? 1 : 0;
    if (precision != SN_PREC_TF32 && precision != SN_PREC_BF16) return 0;
    const int oh = (H + 2 * pad_h - kh) / stride + 1, ow = (W + 2 * pad_w - kw) / stride + 1;
    if (stem_tc_eligible(N, C, H, W, F, kh, kw, stride, pad_h, pad_w)) return 1;
    if (direct_fprop_eligible(C, kh, kw, F)) return 0;
    return (conv_tc_eligible(N, C, oh, ow, F, stride) && igemm_stats_ok(F)) ? 1 : 0;
}

size_t sn_conv2d_dgrad_workspace(int C, int F, int kh, int kw, int precision) {
    if (precision == SN_PREC_FP32 || C <= 0 || F <= 0 || kh <= 0 || kw <= 0) return 0;
    return sizeof(float) * (size_t)C * F * kh * kw;
}

int sn_conv2d_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                    int stride, int pad_h, int pad_w, int accumulate, int precision, void *workspace, void *stream) {
    SN_REQUIRE(dy && w && dx, "null tensor");
    int e = check_prec(precision);
    if (e) return e;
    if (N > 0 && precision != SN_PREC_FP32 && workspace && stride == 1 && H + 2 * pad_h >= kh && W + 2 * pad_w >= kw &&
        kh - 1 - pad_h >= 0 && kw - 1 - pad_w >= 0) {
        // dX = forward convolution of dY (N, oh, ow, F) with Wt[c][kh-1-u][kw-1-v][f], padding k-1-p, output (H, W)
        const int oh = H + 2 * pad_h - kh + 1, ow = W + 2 * pad_w - kw + 1;
        if (conv_tc_eligible(N, F, H, W, C, 1)) {
            e = flip_transpose_filters(w, (float *)workspace, F, C, kh, kw, as_stream(stream));
            if (e) return e;
            return conv_igemm_tc(dy, (const float *)workspace, nullptr, dx, N, F, oh, ow, C, kh, kw, H, W, kh - 1 - pad_h,
                                 kw - 1 - pad_w, SN_ACT_LINEAR, accumulate, false, nullptr, as_stream(stream));
        }
    }
    return conv_dgrad_simt(dy, w, dx, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, accumulate, as_stream(stream));
}

int sn_conv2d_wgrad(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                    int kw, int stride, int pad_h, int pad_w, int accumulate, int precision, void *stream) {
    SN_REQUIRE(x && dy && dw, "null tensor");
    int e = check_prec(precision);
    if (e) return e;
    const bool geom_ok = N > 0 && C > 0 && F > 0 && kh > 0 && kw > 0 && stride > 0 && pad_h >= 0 && pad_w >= 0 &&
                         H + 2 * pad_h >= kh && W + 2 * pad_w >= kw;
    if (geom_ok && precision != SN_PREC_FP32 && (((uintptr_t)dy | (uintptr_t)x) & 15) == 0 && (C * kh * kw < 32 || !db) &&
        stem_tc_eligible(N, C, H, W, F, kh, kw, stride, pad_h, pad_w)) {
        cudaStream_t st = as_stream(stream);
        if (!accumulate) {
            SN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)F * kh * kw * C, st));
            if (db) SN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * F, st));
        }
        return conv_stem_wgrad_tc(x, dy, dw, db, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, st);
    }
    if (geom_ok && tiny_wgrad_eligible(C, kh, kw, F) && (long long)N * H * W * C < (1ll << 31) &&
        (long long)N * F * ((H + 2 * pad_h - kh) / stride + 1) * ((W + 2 * pad_w - kw) / stride + 1) < (1ll << 31)) {
        cudaStream_t st = as_stream(stream);
        if (!accumulate) {
            SN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)F * kh * kw * C, st));
            if (db) SN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * F, st));
        }
        return conv_wgrad_tiny(x, dy, dw, db, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, st);
    }
    if (geom_ok && (((uintptr_t)dy) & 15) == 0 && direct_wgrad_eligible(C, kh, kw, F)) {
        cudaStream_t st = as_stream(stream);
        if (!accumulate) {
            SN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)F * kh * kw * C, st));
            if (db) SN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * F, st));
        }
        return conv_wgrad_direct(x, dy, dw, db, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, st);
    }
    if (geom_ok && precision != SN_PREC_FP32) {
        const int oh = (H + 2 * pad_h - kh) / stride + 1, ow = (W + 2 * pad_w - kw) / stride + 1;
        const bool halo = wgrad_halo_eligible(N, C, oh, ow, F, kh, kw, stride);
        if (halo || wgrad_tc_eligible(N, C, oh, ow, F, stride)) {
            cudaStream_t st = as_stream(stream);
            if (!accumulate) {
                SN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)F * kh * kw * C, st));
                if (db) SN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * F, st));
            }
            e = halo ? conv_wgrad_halo_tc(x, dy, dw, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, false, st)
                     : conv_wgrad_tc(x, dy, dw, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, st);
            if (e || !db) return e;
            return colsum_accumulate(dy, db, (long long)N * oh * ow, F, st);
        }
    }
    return conv_wgrad_simt(x, dy, dw, db, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, accumulate, as_stream(stream));
}

// ------------------------------------------------------------------------------ 3xTF32 (near-fp32 tensor-core tier)
// x*w = xh*wh + xh*wl + xl*wh + O(2^-22): three passes of the tf32 kernels over (x, w), (x_lo, w), (x, w_lo),
// accumulating in fp32; the hi parts are the fp32 buffers themselves (the tensor core truncates).
// Measured: every operator within 1e-5 of float64, but the tensor core's fp32 accumulation is not
// FFMA's round-to-nearest (adding the lo*lo term changes nothing), so multi-step trajectories drift to
// ~3e-5 where the CUDA-core tier stays under 1e-5: the tier is therefore its own precision='tf32x3'.
int sn_split_tf32_lo(const float *src, float *dst_lo, long long n, void *stream) {
    SN_REQUIRE(src && dst_lo && n >= 0, "bad arguments");
    return split_tf32_lo(src, dst_lo, n, as_stream(stream));
}

static bool x3_geom(int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw, int &oh, int &ow) {
    if (N <= 0 || C <= 0 || F <= 0 || kh <= 0 || kw <= 0 || stride != 1 || ph < 0 || pw < 0 || H + 2 * ph < kh || W + 2 * pw < kw)
        return false;
    oh = H + 2 * ph - kh + 1; ow = W + 2 * pw - kw + 1;
    return true;
}

int sn_conv2d_x3_supported(int op, int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w) {
    int oh, ow;
    if (!x3_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow)) return 0;
    if (op == 0) return (F % 4 == 0 && conv_tc_eligible(N, C, oh, ow, F, 1)) ? 1 : 0;
    if (op == 1) return (kh - 1 - pad_h >= 0 && kw - 1 - pad_w >= 0 && C % 4 == 0 && conv_tc_eligible(N, F, H, W, C, 1)) ? 1 : 0;
    if (op == 2) return wgrad_halo_eligible(N, C, oh, ow, F, kh, kw, 1) ? 1 : 0;
    return 0;
}

int sn_conv2d_fprop_x3(const float *x, const float *x_lo, const float *w, const float *w_lo, const float *bias, float *y, int N, int C,
                       int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w, int act, void *stream) {
    SN_REQUIRE(x && x_lo && w && w_lo && y, "null tensor");
    SN_REQUIRE(act == SN_ACT_LINEAR || act == SN_ACT_RELU, "unknown activation");
    SN_REQUIRE(sn_conv2d_x3_supported(0, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the 3xTF32 kernels");
    int oh, ow;
    x3_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    cudaStream_t st = as_stream(stream);
    int e = conv_igemm_tc(x, w, nullptr, y, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, SN_ACT_LINEAR, 0, false, nullptr, st);
    if (e) return e;
    e = conv_igemm_tc(x_lo, w, nullptr, y, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, SN_ACT_LINEAR, 1, false, nullptr, st);
    if (e) return e;
    return conv_igemm_tc(x, w_lo, bias, y, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, act, 1, false, nullptr, st);
}

// workspace: 2 * C*F*kh*kw floats (the flipped + transposed filters and their lo parts)
int sn_conv2d_dgrad_x3(const float *dy, const float *dy_lo, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                       int stride, int pad_h, int pad_w, int accumulate, void *workspace, void *stream) {
    SN_REQUIRE(dy && dy_lo && w && dx && workspace, "null tensor");
    SN_REQUIRE(sn_conv2d_x3_supported(1, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the 3xTF32 kernels");
    int oh, ow;
    x3_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    cudaStream_t st = as_stream(stream);
    float *wt = (float *)workspace, *wt_lo = wt + (size_t)C * F * kh * kw;
    int e = flip_transpose_filters(w, wt, F, C, kh, kw, st);
    if (e) return e;
    e = split_tf32_lo(wt, wt_lo, (long long)C * F * kh * kw, st);
    if (e) return e;
    const int ph2 = kh - 1 - pad_h, pw2 = kw - 1 - pad_w;
    e = conv_igemm_tc(dy, wt, nullptr, dx, N, F, oh, ow, C, kh, kw, H, W, ph2, pw2, SN_ACT_LINEAR, accumulate, false, nullptr, st);
    if (e) return e;
    e = conv_igemm_tc(dy_lo, wt, nullptr, dx, N, F, oh, ow, C, kh, kw, H, W, ph2, pw2, SN_ACT_LINEAR, 1, false, nullptr, st);
    if (e) return e;
    return conv_igemm_tc(dy, wt_lo, nullptr, dx, N, F, oh, ow, C, kh, kw, H, W, ph2, pw2, SN_ACT_LINEAR, 1, false, nullptr, st);
}

// dw += (always accumulates); db += column sums of dY when non-NULL
int sn_conv2d_wgrad_x3(const float *x, const float *x_lo, const float *dy, const float *dy_lo, float *dw, float *db, int N, int C, int H,
                       int W, int F, int kh, int kw, int stride, int pad_h, int pad_w, void *stream) {
    SN_REQUIRE(x && x_lo && dy && dy_lo && dw, "null tensor");
    SN_REQUIRE(sn_conv2d_x3_supported(2, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the 3xTF32 kernels");
    int oh, ow;
    x3_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    cudaStream_t st = as_stream(stream);
    int e = conv_wgrad_halo_tc(x, dy, dw, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, false, st);
    if (e) return e;
    e = conv_wgrad_halo_tc(x_lo, dy, dw, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, false, st);
    if (e) return e;
    e = conv_wgrad_halo_tc(x, dy_lo, dw, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, false, st);
    if (e || !db) return e;
    return colsum_accumulate(dy, db, (long long)N * oh * ow, F, st);
}

// ------------------------------------------------------------------------------ bf16 tier
static bool bf16_geom(int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw, int &oh, int &ow) {
    if (N <= 0 || C <= 0 || F <= 0 || kh <= 0 || kw <= 0 || stride != 1 || ph < 0 || pw < 0 || H + 2 * ph < kh || W + 2 * pw < kw)
        return false;
    oh = H + 2 * ph - kh + 1; ow = W + 2 * pw - kw + 1;
    return true;
}

int sn_cast_f32_to_bf16(const float *src, void *dst_bf16, long long n, void *stream) {
    SN_REQUIRE(src && dst_bf16 && n >= 0, "bad arguments");
    return cast_f32_to_bf16(src, dst_bf16, n, as_stream(stream));
}
int sn_flip_transpose_filters_bf16(const float *w, void *wt_bf16, int F, int C, int kh, int kw, void *stream) {
    SN_REQUIRE(w && wt_bf16 && F > 0 && C > 0 && kh > 0 && kw > 0, "bad arguments");
    return flip_transpose_filters_bf16(w, wt_bf16, F, C, kh, kw, as_stream(stream));
}
int sn_conv2d_bf16_supported(int op, int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w) {
    int oh, ow;
    if (!bf16_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow)) return 0;
    if (op == 0) return (F % 4 == 0 && conv_tc_eligible(N, C, oh, ow, F, 1, true)) ? 1 : 0;
    if (op == 1) return (kh - 1 - pad_h >= 0 && kw - 1 - pad_w >= 0 && C % 4 == 0 && conv_tc_eligible(N, F, H, W, C, 1, true)) ? 1 : 0;
    if (op == 2) return wgrad_halo_eligible(N, C, oh, ow, F, kh, kw, 1, true) ? 1 : 0;
    return 0;
}
int sn_conv_weights_prepare_bf16(const float *arena, void *shadow_bf16, long long n, const long long *tasks_dev, int ntasks,
                                 int flip_units, void *stream) {
    SN_REQUIRE(arena && shadow_bf16 && n >= 0 && ntasks >= 0 && ntasks <= 64 && (ntasks == 0 || tasks_dev) && flip_units >= 0, "bad arguments");
    SN_REQUIRE(((((uintptr_t)arena) & 15) | (((uintptr_t)shadow_bf16) & 7)) == 0, "misaligned arena");
    return prepare_weights_bf16(arena, shadow_bf16, n, tasks_dev, ntasks, flip_units, as_stream(stream));
}
int sn_conv2d_fprop_bf16(const void *x, const void *w, const float *bias, float *y, int N, int C, int H, int W, int F, int kh,
                         int kw, int stride, int pad_h, int pad_w, int act, void *stream) {
    SN_REQUIRE(x && w && y, "null tensor");
    SN_REQUIRE(sn_conv2d_bf16_supported(0, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the bf16 kernels");
    int oh, ow;
    bf16_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    return conv_igemm_tc(x, w, bias, y, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, act, 0, true, nullptr, as_stream(stream));
}
int sn_conv2d_fprop_bf16_stats(const void *x, const void *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                               int kh, int kw, int stride, int pad_h, int pad_w, int act, double *bn_scratch, void *stream) {
    SN_REQUIRE(x && w && y && bn_scratch, "null tensor");
    SN_REQUIRE(sn_conv2d_fprop_stats_supported(SN_PREC_BF16, 1, N, C, H, W, F, kh, kw, stride, pad_h, pad_w),
               "shape not supported by the bf16 kernels with statistics");
    int oh, ow;
    bf16_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    return conv_igemm_tc(x, w, bias, y, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, act, 0, true, bn_scratch, as_stream(stream));
}
// bf16 tier, convolution in front of a fused BatchNormalization (+ pool): bf16 operands, fp32 accumulation,
// the output map stored in bf16 and its per-channel sums (of the stored values) added to bn_scratch
int sn_conv2d_fprop_bf16_stats_ybf16(const void *x, const void *w, const float *bias, void *y_bf16, int N, int C, int H, int W, int F,
                                     int kh, int kw, int stride, int pad_h, int pad_w, int act, double *bn_scratch, void *stream) {
    SN_REQUIRE(x && w && y_bf16 && bn_scratch, "null tensor");
    SN_REQUIRE(sn_conv2d_fprop_ybf16_supported(1, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported with a bf16 output map");
    int oh, ow;
    bf16_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    return conv_igemm_tc(x, w, bias, y_bf16, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, act, 0, true, bn_scratch, as_stream(stream), true);
}
// the tiny-C stem (fp32 input image and filters, TF32 products) with a bf16 output map + statistics
int sn_conv2d_fprop_stats_ybf16(const float *x, const float *w, const float *bias, void *y_bf16, int N, int C, int H, int W, int F,
                                int kh, int kw, int stride, int pad_h, int pad_w, int act, double *bn_scratch, void *stream) {
    SN_REQUIRE(x && w && y_bf16 && bn_scratch, "null tensor");
    SN_REQUIRE(sn_conv2d_fprop_ybf16_supported(0, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported with a bf16 output map");
    return conv_stem_fprop_tc(x, w, bias, y_bf16, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, bn_scratch, as_stream(stream), true);
}
int sn_conv2d_fprop_ybf16_supported(int bf16_operands, int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w) {
    if ((F & 7) || F <= 32 || F > 512) return 0;
    if (bf16_operands) return sn_conv2d_fprop_stats_supported(SN_PREC_BF16, 1, N, C, H, W, F, kh, kw, stride, pad_h, pad_w);
    if (N <= 0 || C <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad_h < 0 || pad_w < 0 || H + 2 * pad_h < kh || W + 2 * pad_w < kw) return 0;
    return stem_tc_eligible(N, C, H, W, F, kh, kw, stride, pad_h, pad_w) ? 1 : 0;
}
int sn_conv2d_dgrad_bf16(const void *dy, const void *wt, float *dx, int N, int C, int H, int W, int F, int kh, int kw, int stride,
                         int pad_h, int pad_w, int accumulate, void *stream) {
    SN_REQUIRE(dy && wt && dx, "null tensor");
    SN_REQUIRE(sn_conv2d_bf16_supported(1, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the bf16 kernels");
    int oh, ow;
    bf16_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    return conv_igemm_tc(dy, wt, nullptr, dx, N, F, oh, ow, C, kh, kw, H, W, kh - 1 - pad_h, kw - 1 - pad_w, SN_ACT_LINEAR,
                         accumulate, true, nullptr, as_stream(stream));
}
int sn_conv2d_wgrad_bf16(const void *x, const void *dy, const float *dy_f32, float *dw, float *db, int N, int C, int H, int W,
                         int F, int kh, int kw, int stride, int pad_h, int pad_w, void *stream) {
    SN_REQUIRE(x && dy && dw, "null tensor");
    SN_REQUIRE(sn_conv2d_bf16_supported(2, N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the bf16 kernels");
    int oh, ow;
    bf16_geom(N, C, H, W, F, kh, kw, stride, pad_h, pad_w, oh, ow);
    int e = conv_wgrad_halo_tc(x, dy, dw, N, C, H, W, F, kh, kw, oh, ow, pad_h, pad_w, true, as_stream(stream));
    if (e || !db || !dy_f32) return e;
    return colsum_accumulate(dy_f32, db, (long long)N * oh * ow, F, as_stream(stream));
}

// Tiny-C stem (C*kh*kw < 32), bf16 tier: the weight gradient from the fp32 input and the bf16 copy of the gradient map
// (what the fused BN+pool backward writes; the fp32 map then never exists).  dw / db accumulated into (+=).
int sn_conv2d_stem_wgrad_dybf16_supported(int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w) {
    if (N <= 0 || C <= 0 || F <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad_h < 0 || pad_w < 0 || H + 2 * pad_h < kh || W + 2 * pad_w < kw)
        return 0;
    const char *e = getenv("SN_STEM_WGRAD_BF16");
    if (e && e[0] == '0') return 0;
    return (C * kh * kw < 32 && (F & 7) == 0 && stem_tc_eligible(N, C, H, W, F, kh, kw, stride, pad_h, pad_w)) ? 1 : 0;
}
int sn_conv2d_stem_wgrad_dybf16(const float *x, const void *dy_bf16, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                                int kw, int stride, int pad_h, int pad_w, void *stream) {
    SN_REQUIRE(x && dy_bf16 && dw, "null tensor");
    SN_REQUIRE(sn_conv2d_stem_wgrad_dybf16_supported(N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the bf16 stem wgrad");
    SN_REQUIRE(((((uintptr_t)x) | ((uintptr_t)dy_bf16)) & 15) == 0, "misaligned tensor");
    return conv_stem_wgrad_tc(x, dy_bf16, dw, db, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, as_stream(stream), true);
}

// ------------------------------------------------------------------------------ tiny-C stem, TMA im2col
int sn_conv2d_stem_supported(int N, int C, int H, int W, int F, int kh, int kw, int stride, int pad_h, int pad_w) {
    if (N <= 0 || C <= 0 || F <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad_h < 0 || pad_w < 0 || H + 2 * pad_h < kh || W + 2 * pad_w < kw)
        return 0;
    const char *e = getenv("SN_STEM_IM2COL");
    if (e && e[0] == '0') return 0;
    return stem_im2col_eligible(N, C, H, W, F, kh, kw, stride, pad_h, pad_w) ? 1 : 0;
}
int sn_conv2d_stem_prepare(const float *x, float *x4, int N, int C, int H, int W, void *stream) {
    SN_REQUIRE(x && x4 && N >= 0 && C >= 1 && C <= 4 && H > 0 && W > 0, "bad arguments");
    SN_REQUIRE((((uintptr_t)x4) & 15) == 0, "x4 must be 16-byte aligned");
    return pad_channels4(x, x4, (long long)N * H * W, C, as_stream(stream));
}
int sn_conv2d_stem_fprop(const float *x4, const float *w, const float *bias, void *y, int N, int C, int H, int W, int F, int kh, int kw,
                         int stride, int pad_h, int pad_w, int act, int y_bf16, double *bn_scratch, void *stream) {
    SN_REQUIRE(x4 && w && y, "null tensor");
    SN_REQUIRE(act == SN_ACT_LINEAR || act == SN_ACT_RELU, "unknown activation");
    SN_REQUIRE(sn_conv2d_stem_supported(N, C, H, W, F, kh, kw, stride, pad_h, pad_w), "shape not supported by the im2col stem kernels");
    SN_REQUIRE(((((uintptr_t)x4) | ((uintptr_t)y) | ((uintptr_t)bias)) & 15) == 0, "misaligned tensor");
    return conv_stem_im2col_fprop(x4, w, bias, y, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, bn_scratch, as_stream(stream), y_bf16 != 0);
}

// Dense is the H = W = kh = kw = 1 convolution: x[M][K] are M "pixels" of K channels.
int sn_dense_fprop(const float *x, const float *w, const float *bias, float *y, int M, int K, int Nout, int act,
                   int precision, void *stream) {
    return sn_conv2d_fprop(x, w, bias, y, M, K, 1, 1, Nout, 1, 1, 1, 0, 0, act, precision, stream);
}
int sn_dense_dgrad(const float *dy, const float *w, float *dx, int M, int K, int Nout, int accumulate, int precision,
                   void *workspace, void *stream) {
    return sn_conv2d_dgrad(dy, w, dx, M, K, 1, 1, Nout, 1, 1, 1, 0, 0, accumulate, precision, workspace, stream);
}
int sn_dense_wgrad(const float *x, const float *dy, float *dw, float *db, int M, int K, int Nout, int accumulate,
                   int precision, void *stream) {
    return sn_conv2d_wgrad(x, dy, dw, db, M, K, 1, 1, Nout, 1, 1, 1, 0, 0, accumulate, precision, stream);
}

}  // extern "C"
