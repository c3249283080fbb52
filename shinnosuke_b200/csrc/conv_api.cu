// extern "C" entry points of Conv2D and Dense; picks the kernel family for the requested tier.
#include "common.cuh"
using namespace sn;

namespace sn {
int conv_fprop_simt(const float *, const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv_dgrad_simt(const float *, const float *, float *, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv_wgrad_simt(const float *, const float *, float *, float *, int, int, int, int, int, int, int, int, int, int, int, cudaStream_t);
}

static int check_prec(int precision) {
    if (precision != SN_PREC_FP32 && precision != SN_PREC_TF32 && precision != SN_PREC_BF16) {
        set_error("unknown precision tier %d", precision);
        return SN_ERR_INVALID_ARG;
    }
    return SN_OK;
}

extern "C" {

int sn_conv2d_fprop(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                    int kh, int kw, int stride, int pad_h, int pad_w, int act, int precision, void *stream) {
    SN_REQUIRE(x && w && y, "null tensor");
    SN_REQUIRE(act == SN_ACT_LINEAR || act == SN_ACT_RELU, "unknown activation");
    int e = check_prec(precision);
    if (e) return e;
    return conv_fprop_simt(x, w, bias, y, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, act, as_stream(stream));
}

int sn_conv2d_dgrad(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                    int stride, int pad_h, int pad_w, int accumulate, int precision, void *stream) {
    SN_REQUIRE(dy && w && dx, "null tensor");
    int e = check_prec(precision);
    if (e) return e;
    return conv_dgrad_simt(dy, w, dx, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, accumulate, as_stream(stream));
}

int sn_conv2d_wgrad(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                    int kw, int stride, int pad_h, int pad_w, int accumulate, int precision, void *stream) {
    SN_REQUIRE(x && dy && dw, "null tensor");
    int e = check_prec(precision);
    if (e) return e;
    return conv_wgrad_simt(x, dy, dw, db, N, C, H, W, F, kh, kw, stride, pad_h, pad_w, accumulate, as_stream(stream));
}

// Dense is the H = W = kh = kw = 1 convolution: x[M][K] are M "pixels" of K channels.
int sn_dense_fprop(const float *x, const float *w, const float *bias, float *y, int M, int K, int Nout, int act,
                   int precision, void *stream) {
    return sn_conv2d_fprop(x, w, bias, y, M, K, 1, 1, Nout, 1, 1, 1, 0, 0, act, precision, stream);
}
int sn_dense_dgrad(const float *dy, const float *w, float *dx, int M, int K, int Nout, int accumulate, int precision,
                   void *stream) {
    return sn_conv2d_dgrad(dy, w, dx, M, K, 1, 1, Nout, 1, 1, 1, 0, 0, accumulate, precision, stream);
}
int sn_dense_wgrad(const float *x, const float *dy, float *dw, float *db, int M, int K, int Nout, int accumulate,
                   int precision, void *stream) {
    return sn_conv2d_wgrad(x, dy, dw, db, M, K, 1, 1, Nout, 1, 1, 1, 0, 0, accumulate, precision, stream);
}

}  // extern "C"
