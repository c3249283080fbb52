// Layout transposes, casts, ReLU, loss and optimizer kernels (all HBM-bound, fp32).
#include "common.cuh"
using namespace sn;

// ------------------------------------------------------------------ transposes
// One image at a time: src viewed as [R][S] -> dst [S][R], 32x32 tiles through shared memory
// so that both the read and the write are 128-byte coalesced.
template <bool ACC>
__global__ void transpose_batched_kernel(const float *__restrict__ src, float *__restrict__ dst, int R, int S) {
    __shared__ float tile[32][33];
    const long long img = (long long)blockIdx.z * R * S;
    int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int r = r0 + i, s = s0 + threadIdx.x;
        if (r < R && s < S) tile[i][threadIdx.x] = src[img + (long long)r * S + s];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int s = s0 + i, r = r0 + threadIdx.x;
        if (r < R && s < S) {
            float v = tile[threadIdx.x][i];
            float *p = dst + img + (long long)s * R + r;
            if (ACC) *p += v; else *p = v;
        }
    }
}

static int transpose_batched(const float *src, float *dst, int batch, int R, int S, bool acc, cudaStream_t st) {
    if ((long long)batch * R * S == 0) return SN_OK;
    if ((R == 1 || S == 1) && !acc) {
        SN_CUDA(cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)batch * R * S, cudaMemcpyDeviceToDevice, st));
        return SN_OK;
    }
    dim3 block(32, 8), grid((S + 31) / 32, (R + 31) / 32, batch);
    SN_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "tensor too large for the transpose grid");
    if (acc) transpose_batched_kernel<true><<<grid, block, 0, st>>>(src, dst, R, S);
    else     transpose_batched_kernel<false><<<grid, block, 0, st>>>(src, dst, R, S);
    return after_launch("transpose_batched_kernel");
}

// ------------------------------------------------------------------ casts / axpy / relu
__global__ void cast_f64_f32_kernel(const double *__restrict__ s, float *__restrict__ d, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        d[i] = (float)s[i];
}
__global__ void cast_f32_f64_kernel(const float *__restrict__ s, double *__restrict__ d, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        d[i] = (double)s[i];
}
__global__ void axpy_kernel(float a, const float *__restrict__ x, float *__restrict__ y, long long n) {
    pdl_wait();
    long long n4 = n >> 2;
    long long stride = (long long)gridDim.x * blockDim.x;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0) {
        for (long long i = t; i < n4; i += stride) {
            float4 xv = ldg4(x + 4 * i);
            float4 yv = reinterpret_cast<float4 *>(y)[i];
            yv.x += a * xv.x; yv.y += a * xv.y; yv.z += a * xv.z; yv.w += a * xv.w;
            reinterpret_cast<float4 *>(y)[i] = yv;
        }
        for (long long i = 4 * n4 + t; i < n; i += stride) y[i] += a * x[i];
    } else {
        for (long long i = t; i < n; i += stride) y[i] += a * x[i];
    }
}
template <bool RELU>
__global__ void add_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ o, long long n) {
    pdl_wait();
    long long n4 = n >> 2;
    long long stride = (long long)gridDim.x * blockDim.x;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    auto f = [](float v) { return RELU ? relu_keep_sign(v) : v; };
    if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)o) & 15) == 0) {
        for (long long i = t; i < n4; i += stride) {
            float4 x = ldg4(a + 4 * i), y = ldg4(b + 4 * i);
            reinterpret_cast<float4 *>(o)[i] = make_float4(f(x.x + y.x), f(x.y + y.y), f(x.z + y.z), f(x.w + y.w));
        }
        for (long long i = 4 * n4 + t; i < n; i += stride) o[i] = f(a[i] + b[i]);
    } else {
        for (long long i = t; i < n; i += stride) o[i] = f(a[i] + b[i]);
    }
}
__global__ void relu_bwd_kernel(float *__restrict__ dy, const float *__restrict__ y, long long n) {
    pdl_wait();
    long long n4 = n >> 2;
    long long stride = (long long)gridDim.x * blockDim.x;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (((uintptr_t)dy & 15) == 0 && ((uintptr_t)y & 15) == 0) {
        for (long long i = t; i < n4; i += stride) {
            float4 yv = ldg4(y + 4 * i);
            float4 g = reinterpret_cast<float4 *>(dy)[i];
            if (relu_blocked(yv.x)) g.x = 0.f;
            if (relu_blocked(yv.y)) g.y = 0.f;
            if (relu_blocked(yv.z)) g.z = 0.f;
            if (relu_blocked(yv.w)) g.w = 0.f;
            reinterpret_cast<float4 *>(dy)[i] = g;
        }
        for (long long i = 4 * n4 + t; i < n; i += stride) if (relu_blocked(y[i])) dy[i] = 0.f;
    } else {
        for (long long i = t; i < n; i += stride) if (relu_blocked(y[i])) dy[i] = 0.f;
    }
}
__global__ void relu_fwd_kernel(const float *x, float *y, long long n) {
    pdl_wait();
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if ((((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
        for (long long i = t; i < n4; i += stride) {
            const float4 v = reinterpret_cast<const float4 *>(x)[i];          // (plain load: x may alias y)
            reinterpret_cast<float4 *>(y)[i] = make_float4(relu_keep_sign(v.x), relu_keep_sign(v.y), relu_keep_sign(v.z), relu_keep_sign(v.w));
        }
        for (long long i = 4 * n4 + t; i < n; i += stride) y[i] = relu_keep_sign(x[i]);
    } else {
        for (long long i = t; i < n; i += stride) y[i] = relu_keep_sign(x[i]);
    }
}

// ------------------------------------------------------------------ zero padding
// FWD: one thread per element of the padded map.  BWD: one thread per element of dx.
template <bool FWD, bool ACC>
__global__ void pad2d_kernel(const float *__restrict__ src, float *__restrict__ dst, int N, int H, int W, int C,
                             int ph, int pw) {
    const int Hp = H + 2 * ph, Wp = W + 2 * pw;
    const long long total = FWD ? (long long)N * Hp * Wp * C : (long long)N * H * W * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        long long pix = idx / C;
        if (FWD) {
            int w = (int)(pix % Wp) - pw, h = (int)((pix / Wp) % Hp) - ph;
            long long n = pix / ((long long)Wp * Hp);
            float v = 0.f;
            if ((unsigned)h < (unsigned)H && (unsigned)w < (unsigned)W) v = src[((n * H + h) * W + w) * C + c];
            dst[idx] = v;
        } else {
            int w = (int)(pix % W), h = (int)((pix / W) % H);
            long long n = pix / ((long long)W * H);
            float v = src[((n * Hp + h + ph) * Wp + w + pw) * C + c];
            dst[idx] = ACC ? dst[idx] + v : v;
        }
    }
}

// ------------------------------------------------------------------ softmax / SCCE
// One warp per row.
__global__ void softmax_fwd_kernel(const float *__restrict__ z, float *__restrict__ p, int rows, int classes) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float *zr = z + (long long)row * classes;
    float m = -INFINITY;
    for (int c = lane; c < classes; c += 32) m = fmaxf(m, zr[c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < classes; c += 32) s += expf(zr[c] - m);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float inv = 1.f / s;
    for (int c = lane; c < classes; c += 32) p[(long long)row * classes + c] = expf(zr[c] - m) * inv;
}

__global__ void scce_bwd_kernel(const float *__restrict__ p, const float *__restrict__ y, float *__restrict__ dz,
                                float *__restrict__ metrics, int rows, int classes, float inv_rows) {
    pdl_wait();
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    float loss = 0.f, hit = 0.f;
    if (row < rows) {
        const float *pr = p + (long long)row * classes;
        const float *yr = y + (long long)row * classes;
        float pbest = -INFINITY, ybest = -INFINITY;
        int pi = 0x7fffffff, yi = 0x7fffffff;
        for (int c = lane; c < classes; c += 32) {
            float pv = pr[c], yv = yr[c];
            dz[(long long)row * classes + c] = (pv - yv) * inv_rows;
            // a probability that underflowed fp32 (logit gap > 87; the float64 reference still has it) counts
            // as the smallest normal float instead of log(0) = -inf, which would poison the minibatch mean
            if (yv != 0.f) loss -= yv * logf(fmaxf(pv, 1.17549435e-38f));
            if (pv > pbest) { pbest = pv; pi = c; }
            if (yv > ybest) { ybest = yv; yi = c; }
        }
        for (int o = 16; o; o >>= 1) {
            loss += __shfl_xor_sync(0xffffffffu, loss, o);
            float ob = __shfl_xor_sync(0xffffffffu, pbest, o); int oi = __shfl_xor_sync(0xffffffffu, pi, o);
            if (ob > pbest || (ob == pbest && oi < pi)) { pbest = ob; pi = oi; }
            ob = __shfl_xor_sync(0xffffffffu, ybest, o); oi = __shfl_xor_sync(0xffffffffu, yi, o);
            if (ob > ybest || (ob == ybest && oi < yi)) { ybest = ob; yi = oi; }
        }
        hit = (pi == yi) ? 1.f : 0.f;
    }
    if (metrics) {
        __shared__ float sl[32], sh[32];
        int w = threadIdx.x >> 5;
        if (lane == 0) { sl[w] = loss; sh[w] = hit; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.f, b = 0.f;
            for (int i = 0; i < (blockDim.x >> 5); ++i) { a += sl[i]; b += sh[i]; }
            atomicAdd(metrics + 0, a);
            atomicAdd(metrics + 1, b);
        }
    }
}

// ------------------------------------------------------------------ optimizers
__global__ void sgd_kernel(float *__restrict__ w, float *__restrict__ g, long long n, float lr_scaled) {
    pdl_wait();
    long long n4 = n >> 2;
    long long stride = (long long)gridDim.x * blockDim.x;
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float4 *w4 = reinterpret_cast<float4 *>(w);
    float4 *g4 = reinterpret_cast<float4 *>(g);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = t; i < n4; i += stride) {
        float4 wv = w4[i], gv = g4[i];
        wv.x -= lr_scaled * gv.x; wv.y -= lr_scaled * gv.y; wv.z -= lr_scaled * gv.z; wv.w -= lr_scaled * gv.w;
        w4[i] = wv;
        g4[i] = zero;
    }
    for (long long i = 4 * n4 + t; i < n; i += stride) { w[i] -= lr_scaled * g[i]; g[i] = 0.f; }
}
__global__ void sgd_scalar_kernel(float *__restrict__ w, float *__restrict__ g, long long n, float lr_scaled) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        w[i] -= lr_scaled * g[i]; g[i] = 0.f;
    }
}
extern "C" {

int sn_nchw_to_nhwc(const float *src, float *dst, int N, int C, int H, int W, void *stream) {
    SN_REQUIRE(src && dst && N >= 0 && C > 0 && H > 0 && W > 0, "bad arguments");
    return transpose_batched(src, dst, N, C, H * W, false, as_stream(stream));
}
int sn_nhwc_to_nchw(const float *src, float *dst, int N, int C, int H, int W, int accumulate, void *stream) {
    SN_REQUIRE(src && dst && N >= 0 && C > 0 && H > 0 && W > 0, "bad arguments");
    return transpose_batched(src, dst, N, H * W, C, accumulate != 0, as_stream(stream));
}
int sn_cast_f64_to_f32(const double *src, float *dst, long long n, void *stream) {
    SN_REQUIRE(src && dst && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    cast_f64_f32_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(src, dst, n);
    return after_launch("cast_f64_f32_kernel");
}
int sn_cast_f32_to_f64(const float *src, double *dst, long long n, void *stream) {
    SN_REQUIRE(src && dst && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    cast_f32_f64_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(src, dst, n);
    return after_launch("cast_f32_f64_kernel");
}
int sn_axpy(float alpha, const float *x, float *y, long long n, void *stream) {
    SN_REQUIRE(x && y && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    SN_CUDA(launch_dependent(axpy_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), alpha, x, y, n));
    return after_launch("axpy_kernel");
}
int sn_add(const float *a, const float *b, float *out, long long n, void *stream) {
    SN_REQUIRE(a && b && out && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    SN_CUDA(launch_dependent(add_kernel<false>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), a, b, out, n));
    return after_launch("add_kernel");
}
int sn_add_relu(const float *a, const float *b, float *out, long long n, void *stream) {
    SN_REQUIRE(a && b && out && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    SN_CUDA(launch_dependent(add_kernel<true>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), a, b, out, n));
    return after_launch("add_kernel");
}
int sn_relu_bwd(float *dy, const float *y, long long n, void *stream) {
    SN_REQUIRE(dy && y && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    SN_CUDA(launch_dependent(relu_bwd_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), dy, y, n));
    return after_launch("relu_bwd_kernel");
}
int sn_relu_fwd(const float *x, float *y, long long n, void *stream) {
    SN_REQUIRE(x && y && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    SN_CUDA(launch_dependent(relu_fwd_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), x, y, n));
    return after_launch("relu_fwd_kernel");
}
int sn_pad2d_fwd(const float *x, float *y, int N, int H, int W, int C, int pad_h, int pad_w, void *stream) {
    SN_REQUIRE(x && y && N >= 0 && H > 0 && W > 0 && C > 0 && pad_h >= 0 && pad_w >= 0, "bad arguments");
    long long total = (long long)N * (H + 2 * pad_h) * (W + 2 * pad_w) * C;
    if (!total) return SN_OK;
    pad2d_kernel<true, false><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(x, y, N, H, W, C, pad_h, pad_w);
    return after_launch("pad2d_kernel");
}
int sn_pad2d_bwd(const float *dy, float *dx, int N, int H, int W, int C, int pad_h, int pad_w, int accumulate, void *stream) {
    SN_REQUIRE(dy && dx && N >= 0 && H > 0 && W > 0 && C > 0 && pad_h >= 0 && pad_w >= 0, "bad arguments");
    long long total = (long long)N * H * W * C;
    if (!total) return SN_OK;
    if (accumulate) pad2d_kernel<false, true><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dy, dx, N, H, W, C, pad_h, pad_w);
    else            pad2d_kernel<false, false><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dy, dx, N, H, W, C, pad_h, pad_w);
    return after_launch("pad2d_kernel");
}
int sn_softmax_fwd(const float *logits, float *probs, int rows, int classes, void *stream) {
    SN_REQUIRE(logits && probs && rows >= 0 && classes > 0, "bad arguments");
    if (!rows) return SN_OK;
    softmax_fwd_kernel<<<(rows + 7) / 8, 256, 0, as_stream(stream)>>>(logits, probs, rows, classes);
    return after_launch("softmax_fwd_kernel");
}
int sn_scce_bwd(const float *probs, const float *onehot, float *dlogits, float *metrics, int rows, int classes,
                int denom_rows, void *stream) {
    SN_REQUIRE(probs && onehot && dlogits && rows >= 0 && classes > 0 && denom_rows >= 0, "bad arguments");
    if (!rows) return SN_OK;
    float inv = 1.f / (float)(denom_rows > 0 ? denom_rows : rows);
    SN_CUDA(launch_dependent(scce_bwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, as_stream(stream), probs, onehot, dlogits, metrics, rows, classes, inv));
    return after_launch("scce_bwd_kernel");
}
int sn_sgd_update(float *w, float *g, long long n, float lr, float grad_scale, void *stream) {
    SN_REQUIRE(w && g && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    if ((((uintptr_t)w | (uintptr_t)g) & 15) == 0)
        SN_CUDA(launch_dependent(sgd_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, as_stream(stream), w, g, n, lr * grad_scale));
    else
        sgd_scalar_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w, g, n, lr * grad_scale);
    return after_launch("sgd_kernel");
}
}  // extern "C"
