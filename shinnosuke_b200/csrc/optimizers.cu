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

template <int KIND>
__global__ void adaptive_kernel(float *__restrict__ w, float *__restrict__ g, float *__restrict__ s1, float *__restrict__ s2,
                                float *__restrict__ gstep, const float *__restrict__ tick, long long n, float lr, float p1,
                                float p2, float eps, float scale) {
    float c1 = 1.f, c2 = 1.f;
    if (KIND == OPT_ADAM) { c1 = tick[1]; c2 = tick[2]; }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i];
        if (gstep) { gi += scale * gstep[i]; gstep[i] = 0.f; g[i] = gi; }
        if (KIND == OPT_RMSPROP) {
            const float s = p1 * s1[i] + (1.f - p1) * gi * gi;
            s1[i] = s;
            w[i] -= lr * gi / sqrtf(s + eps);
        } else if (KIND == OPT_ADAGRAD) {
            const float s = s1[i] + gi * gi;
            s1[i] = s;
            w[i] -= lr * gi / sqrtf(s + eps);
        } else if (KIND == OPT_ADADELTA) {
            const float s = p1 * s1[i] + (1.f - p1) * gi * gi;
            const float u = sqrtf((s2[i] + eps) / (s + eps)) * gi;
            s1[i] = s;
            w[i] -= u;
            s2[i] = p1 * s2[i] + (1.f - p1) * u * u;
        } else {
            const float v = p1 * s1[i] + (1.f - p1) * gi;          // first moment ("v" in the reference)
            const float m = p2 * s2[i] + (1.f - p2) * gi * gi;     // second moment ("m")
            s1[i] = v; s2[i] = m;
            w[i] -= lr * ((v / c1) / (sqrtf(m / c2) + eps));
        }
    }
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
    const int grid = grid_for(n, 256);
    switch (kind) {
        case OPT_RMSPROP: adaptive_kernel<OPT_RMSPROP><<<grid, 256, 0, st>>>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale); break;
        case OPT_ADAGRAD: adaptive_kernel<OPT_ADAGRAD><<<grid, 256, 0, st>>>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale); break;
        case OPT_ADADELTA: adaptive_kernel<OPT_ADADELTA><<<grid, 256, 0, st>>>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale); break;
        default:
            adam_tick_kernel<<<1, 1, 0, st>>>(tick, p1, p2);
            adaptive_kernel<OPT_ADAM><<<grid, 256, 0, st>>>(w, g, s1, s2, gstep, tick, n, lr, p1, p2, eps, grad_scale);
    }
    return after_launch("adaptive_kernel");
}

}  // extern "C"
