// RMSprop / AdaGrad / AdaDelta / Adam over a flat parameter arena (reference:
// utils/Optimizers.py:61-157).  One launch per step; state (s1, s2) is allocated by the caller,
// same length as the arena.  Bug-compatible with the reference where it matters for parity:
//   * none of these optimizers zeroes the gradients (their base class only counts iterations), so
//     `g` keeps accumulating across steps; under data parallelism this step's all-reduced
//     gradients arrive in `gstep` and are folded into the persistent `g` first (g += scale*gstep);
//   * Adam's bias corrections use an iteration counter that advances by TWO per update
//     (1, 3, 5, ...: its own `+= 1` and the base class's).  The counter lives on the device
//     (`tick[0]`), so a captured CUDA graph of the step replays with the right correction.
#include "common.cuh"
using namespace sn;

namespace {

enum { OPT_RMSPROP = 0, OPT_ADAGRAD = 1, OPT_ADADELTA = 2, OPT_ADAM = 3 };

// tick: [0] = iterations as seen by the update (float for the pow), [1] = 1 - b1^it, [2] = 1 - b2^it
__global__ void adam_tick_kernel(float *tick, float beta1, float beta2) {
    const float it = tick[0] + 1.f;            // Adam's own `self.iterations += 1`
    tick[1] = 1.f - powf(beta1, it);
    tick[2] = 1.f - powf(beta2, it);
    tick[0] = it + 1.f;                        // the base class's increment, after the update used `it`
}

enum { OPT_MOMENTUM = 4 };

// one element of the update; s1 / s2 are the optimizer's state words of this element
template <int KIND>
__device__ __forceinline__ void opt_update(float &w, float gi, float &s1, float &s2, float lr, float p1, float p2, float eps,
                                           float c1, float c2) {
    if (KIND == OPT_RMSPROP) {
        const float s = p1 * s1 + (1.f - p1) * gi * gi;
        s1 = s;
        w -= lr * gi / sqrtf(s + eps);
    } else if (KIND == OPT_ADAGRAD) {
        const float s = s1 + gi * gi;
        s1 = s;
        w -= lr * gi / sqrtf(s + eps);
    } else if (KIND == OPT_ADADELTA) {
        const float s = p1 * s1 + (1.f - p1) * gi * gi;
        const float u = sqrtf((s2 + eps) / (s + eps)) * gi;
        s1 = s;
        w -= u;
        s2 = p1 * s2 + (1.f - p1) * u * u;
    } else if (KIND == OPT_ADAM) {
        const float v = p1 * s1 + (1.f - p1) * gi;          // first moment ("v" in the reference)
        const float m = p2 * s2 + (1.f - p2) * gi * gi;     // second moment ("m")
        s1 = v; s2 = m;
        // c1 = 1 / (1 - beta1^t), c2 = 1 / (1 - beta2^t): reciprocals taken once per launch; one square
        // root and one division per element (three IEEE divisions made this kernel issue-bound: 0.57 of HBM)
        w -= __fdividef(lr * (v * c1), sqrtf(m * c2) + eps);
    } else {                                                 // Momentum (utils/Optimizers.py:46-55): p1 = rho
        const float v = p1 * s1 + (1.f - p1) * gi;
        s1 = v;
        w -= lr * v;
    }
}

// 128-bit streams: every tensor of the update is read / written as float4 (the arena and its twins are
// 128-byte aligned), two quads per thread in flight; a scalar tail and a scalar kernel for unaligned
// stand-alone variables.  HBM traffic per element: w r+w, g r (+w and gstep r+w under data parallelism),
// one or two state words r+w = 20 B (Momentum, RMSprop, AdaGrad) or 28 B (AdaDelta, Adam).
template <int KIND, bool VEC>
__global__ void __launch_bounds__(256) adaptive_kernel(float *__restrict__ w, float *__restrict__ g, float *__restrict__ s1,
                                                        float *__restrict__ s2, float *__restrict__ gstep,
                                                        const float *__restrict__ tick, long long n, float lr, float p1,
                                                        float p2, float eps, float scale) {
    constexpr bool TWO = (KIND == OPT_ADADELTA || KIND == OPT_ADAM);
    float c1 = 1.f, c2 = 1.f;
    if (KIND == OPT_ADAM) { c1 = 1.f / tick[1]; c2 = 1.f / tick[2]; }
    const long long stride = (long long)gridDim.x * blockDim.x, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long done = 0;
    if (VEC) {
        const long long n4 = n >> 2;
        float4 *w4 = reinterpret_cast<float4 *>(w), *g4 = reinterpret_cast<float4 *>(g), *a4 = reinterpret_cast<float4 *>(s1);
        float4 *b4 = reinterpret_cast<float4 *>(s2), *q4 = reinterpret_cast<float4 *>(gstep);
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long i = t; i < n4; i += 2 * stride) {
            const long long j = i + stride;
            const bool two = j < n4;
            float4 wv[2], gv[2], av[2], bv[2], qv[2];
            wv[0] = w4[i]; gv[0] = g4[i]; av[0] = a4[i];
            if (TWO) bv[0] = b4[i];
            if (gstep) qv[0] = q4[i];
            if (two) {
                wv[1] = w4[j]; gv[1] = g4[j]; av[1] = a4[j];
                if (TWO) bv[1] = b4[j];
                if (gstep) qv[1] = q4[j];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                const long long k = u ? j : i;
                if (gstep) {
                    gv[u].x += scale * qv[u].x; gv[u].y += scale * qv[u].y; gv[u].z += scale * qv[u].z; gv[u].w += scale * qv[u].w;
                    q4[k] = zero;
                    g4[k] = gv[u];
                }
                float dummy = 0.f;
                opt_update<KIND>(wv[u].x, gv[u].x, av[u].x, TWO ? bv[u].x : dummy, lr, p1, p2, eps, c1, c2);
                opt_update<KIND>(wv[u].y, gv[u].y, av[u].y, TWO ? bv[u].y : dummy, lr, p1, p2, eps, c1, c2);
                opt_update<KIND>(wv[u].z, gv[u].z, av[u].z, TWO ? bv[u].z : dummy, lr, p1, p2, eps, c1, c2);
                opt_update<KIND>(wv[u].w, gv[u].w, av[u].w, TWO ? bv[u].w : dummy, lr, p1, p2, eps, c1, c2);
                w4[k] = wv[u];
                a4[k] = av[u];
                if (TWO) b4[k] = bv[u];
            }
        }
        done = n4 << 2;
    }
    for (long long i = done + t; i < n; i += stride) {
        float gi = g[i];
        if (gstep) { gi += scale * gstep[i]; gstep[i] = 0.f; g[i] = gi; }
        float wi = w[i], a = s1[i], b = TWO ? s2[i] : 0.f;
        opt_update<KIND>(wi, gi, a, b, lr, p1, p2, eps, c1, c2);
        w[i] = wi; s1[i] = a;
        if (TWO) s2[i] = b;
    }
}

template <int KIND>
int launch_adaptive(float *w, float *g, float *s1, float *s2, float *gstep, const float *tick, long long n, float lr, float p1,
                    float p2, float eps, float scale, cudaStream_t st) {
    const bool vec = ((((uintptr_t)w | (uintptr_t)g | (uintptr_t)s1 | (uintptr_t)s2 | (uintptr_t)gstep) & 15) == 0) && n >= 4;
    if (vec)
        adaptive_kernel<KIND, true><<<grid_for((n / 4 + 1) / 2, 256), 256, 0, st>>>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, scale);
    else
        adaptive_kernel<KIND, false><<<grid_for(n, 256), 256, 0, st>>>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, scale);
    return after_launch("adaptive_kernel");
}

}  // namespace

extern "C" {

int sn_adaptive_update(int kind, float *w, float *g, float *s1, float *s2, float *gstep, float *tick, long long n, float lr,
                       float p1, float p2, float eps, float grad_scale, void *stream) {
    SN_REQUIRE(w && g && s1 && n >= 0, "bad arguments");
    SN_REQUIRE(kind >= OPT_RMSPROP && kind <= OPT_ADAM, "unknown optimizer kind");
    SN_REQUIRE((kind != OPT_ADADELTA && kind != OPT_ADAM) || s2, "this optimizer needs a second state buffer");
    SN_REQUIRE(kind != OPT_ADAM || tick, "Adam needs its 3-float device tick buffer");
    if (!n) return SN_OK;
    cudaStream_t st = as_stream(stream);
    switch (kind) {
        case OPT_RMSPROP: return launch_adaptive<OPT_RMSPROP>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale, st);
        case OPT_ADAGRAD: return launch_adaptive<OPT_ADAGRAD>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale, st);
        case OPT_ADADELTA: return launch_adaptive<OPT_ADADELTA>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale, st);
        default:
            adam_tick_kernel<<<1, 1, 0, st>>>(tick, p1, p2);
            ++g_launch_count;
            return launch_adaptive<OPT_ADAM>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale, st);
    }
}

// Momentum (utils/Optimizers.py:46-55): v = rho*v + (1-rho)*g; w -= lr*v; the gradients are never zeroed
int sn_momentum_update(float *w, float *g, float *v, float *gstep, long long n, float lr, float rho, float grad_scale, void *stream) {
    SN_REQUIRE(w && g && v && n >= 0, "bad arguments");
    if (!n) return SN_OK;
    return launch_adaptive<OPT_MOMENTUM>(w, g, v, nullptr, gstep, nullptr, n, lr, rho, 0.f, 0.f, grad_scale, as_stream(stream));
}

}  // extern "C"
