// BatchNormalization fused with the 2x2/stride-2 MaxPooling2D that follows it (the
// Conv -> BN -> Pool block of BASELINE configs 3 and 4).  The normalised map is never written to
// HBM: forward reads the conv output once for the statistics and once to normalise + pool;
// backward scatters the pooled gradient through the pool codes on the fly while it reduces and
// applies the BatchNorm gradient.  Versus the separate layers this removes the write + re-read of
// the BN output (forward) and the write + two re-reads of the pool's dX (backward): about 45 %
// of the block's element-wise HBM traffic.
//
// Semantics are exactly the two reference layers back to back (layers/Normalization.py:93-230,
// layers/Convolution.py:287-362): biased variance, eps in the sqrt, moving statistics, tie-split
// or first-argmax routing, and (optionally) the producing layer's ReLU mask applied to dX.
#include "common.cuh"
#include "bn_common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>
using namespace sn;

namespace {

constexpr int COLSUM_SLOTS = BN_SLOTS;     // slot rows of the dx column-sum reduction: scratch[20C, 28C)

__device__ __forceinline__ float comp(const float4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ void setc(float4 &v, int k, float x) { if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x; }

// window index space: idx = ((n*oh + i)*ow + j)*Q + cq, Q = C/4.  256 % Q == 0 and the grid stride
// is a multiple of 256, so a thread always works on the same channel quad.
// XT: element type of the full-resolution map x — float, or __nv_bfloat16 in the bf16 tier, where the
// producing convolution stores its output map in bf16 (half the bytes of the three passes over it: the
// conv's write, the forward read here, the backward read here).  All arithmetic stays fp32; a blocked
// ReLU element is -0.0 in either type.
template <typename XT> struct WinT { const XT *p00; long long rowstride; };
__device__ __forceinline__ float4 ldx4(const float *p) { return ldg4(p); }
__device__ __forceinline__ float4 ldx4(const __nv_bfloat16 *p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
}
// Index arithmetic of a window: the pooled maps of configs 3 and 4 (and Q, always) are powers of two, so the three
// div / mod pairs are shifts and masks; other maps take 32-bit divisions when the index space fits, 64-bit ones
// otherwise.  (All three used to be 64-bit divisions — ~300 of the ~370 instructions of an iteration of the
// backward apply kernel, which ran at half of its issue rate and 3.3 TB/s instead of at the HBM rate.)
struct WinGeo {
    int Q, ow, oh, W, C;
    int lq, low, loh;          // log2 of Q / ow / oh, -1 when not a power of two
};
__device__ __forceinline__ WinGeo win_geo(int Q, int ow, int oh, int W, int C) {
    WinGeo g;
    g.Q = Q; g.ow = ow; g.oh = oh; g.W = W; g.C = C;
    g.lq = __ffs(Q) - 1;
    g.low = (ow & (ow - 1)) ? -1 : __ffs(ow) - 1;
    g.loh = (oh & (oh - 1)) ? -1 : __ffs(oh) - 1;
    return g;
}
template <typename XT>
__device__ __forceinline__ WinT<XT> window(const XT *x, long long idx, const WinGeo &g) {
    int cq, j, i;
    long long n;
    if (g.low >= 0 && g.loh >= 0) {
        cq = (int)idx & (g.Q - 1);
        const long long pix = idx >> g.lq;
        j = (int)pix & (g.ow - 1);
        const long long t = pix >> g.low;
        i = (int)t & (g.oh - 1);
        n = t >> g.loh;
    } else if (idx < 0x7fffffffll) {
        const unsigned u = (unsigned)idx;
        cq = (int)(u & (unsigned)(g.Q - 1));
        const unsigned pix = u >> g.lq;
        const unsigned t = pix / (unsigned)g.ow;
        j = (int)(pix - t * (unsigned)g.ow);
        const unsigned nn = t / (unsigned)g.oh;
        i = (int)(t - nn * (unsigned)g.oh);
        n = nn;
    } else {
        cq = (int)(idx % g.Q);
        const long long pix = idx / g.Q;
        j = (int)(pix % g.ow);
        const long long t = pix / g.ow;
        i = (int)(t % g.oh);
        n = t / g.oh;
    }
    WinT<XT> w;
    w.rowstride = (long long)g.W * g.C;
    w.p00 = x + ((n * 2 * g.oh + 2 * i) * g.W + 2 * j) * g.C + cq * 4;      // H == 2*oh
    return w;
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16 *dst, const float4 &v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t *>(&lo); u.y = *reinterpret_cast<uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(dst) = u;
}

template <int MODE, typename XT>
__global__ void __launch_bounds__(256) bn_pool_fwd_kernel(const XT *__restrict__ x, float *__restrict__ y,
                                                          __nv_bfloat16 *__restrict__ y_bf16,
                                                          uint32_t *__restrict__ code, const double *__restrict__ scratch,
                                                          long long windows4, int C, int W, int oh, int ow) {
    pdl_wait();
    const int Q = C >> 2;
    const WinGeo geo = win_geo(Q, ow, oh, W, C);
    const int cq = threadIdx.x % Q;
    const float *prm = bn_params(scratch, C);
    const float4 sc4 = ldg4(prm + cq * 4), sh4 = ldg4(prm + C + cq * 4);
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
    const long long stride = (long long)gridDim.x * 256;
    for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < windows4; idx += stride) {
        const WinT<XT> w = window(x, idx, geo);
        float4 v[4];
        v[0] = ldx4(w.p00); v[1] = ldx4(w.p00 + C); v[2] = ldx4(w.p00 + w.rowstride); v[3] = ldx4(w.p00 + w.rowstride + C);
        float4 out;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a0 = fmaf(comp(v[0], k), sc[k], sh[k]), a1 = fmaf(comp(v[1], k), sc[k], sh[k]);
            const float a2 = fmaf(comp(v[2], k), sc[k], sh[k]), a3 = fmaf(comp(v[3], k), sc[k], sh[k]);
            float m = a0; uint32_t am = 0;
            if (a1 > m) { m = a1; am = 1; }
            if (a2 > m) { m = a2; am = 2; }
            if (a3 > m) { m = a3; am = 3; }
            uint32_t cd;
            if (MODE == SN_POOL_FIRST_ARGMAX) cd = am;
            else cd = (a0 == m ? 1u : 0u) | (a1 == m ? 2u : 0u) | (a2 == m ? 4u : 0u) | (a3 == m ? 8u : 0u);
            setc(out, k, m);
            packed |= cd << (8 * k);
        }
        *reinterpret_cast<float4 *>(y + idx * 4) = out;
        if (y_bf16) store_bf16x4(y_bf16 + idx * 4, out);          // the next convolution's bf16 operand copy
        code[idx] = packed;
    }
}

// gradient of the BN output at window element e of channel lane k, from the pooled gradient: every winner of the
// window receives share(g, code) — g itself, or g / (number of tied maxima) in tie-split mode (x0.5 / x0.25 are the
// exact quotients; the division by 3 stays an IEEE division) — computed ONCE per (window, channel), not per element
template <int MODE>
__device__ __forceinline__ float share(float g, uint32_t cd) {
    if (MODE == SN_POOL_FIRST_ARGMAX) return g;
    const int cnt = __popc(cd);
    if (cnt == 2) return g * 0.5f;
    if (cnt == 4) return g * 0.25f;
    if (cnt == 3) return __fdiv_rn(g, 3.f);
    return g;
}
template <int MODE>
__device__ __forceinline__ float routed(float sh, uint32_t cd, int e) {
    if (MODE == SN_POOL_FIRST_ARGMAX) return cd == (uint32_t)e ? sh : 0.f;
    return ((cd >> e) & 1u) ? sh : 0.f;
}

template <int MODE, typename XT>
__global__ void __launch_bounds__(256) bn_pool_bwd_reduce_kernel(const float *__restrict__ dyp, const uint32_t *__restrict__ code,
                                                                 const XT *__restrict__ x, const float *__restrict__ mean,
                                                                 const float *__restrict__ invstd, double *__restrict__ scratch,
                                                                 long long windows4, int C, int W, int oh, int ow, BnFinal fin) {
    const int Q = C >> 2;
    const WinGeo geo = win_geo(Q, ow, oh, W, C);
    const int cq = threadIdx.x % Q;
    float mu[4], is[4];
    {
        const float4 a = ldg4(mean + cq * 4), b = ldg4(invstd + cq * 4);
        mu[0] = a.x; mu[1] = a.y; mu[2] = a.z; mu[3] = a.w; is[0] = b.x; is[1] = b.y; is[2] = b.z; is[3] = b.w;
    }
    double da[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * 256;
    long long idx = blockIdx.x * 256ll + threadIdx.x;
    while (idx < windows4) {
        float fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0};
#pragma unroll 1
        for (int it = 0; it < 16 && idx < windows4; ++it, idx += stride) {
            const WinT<XT> w = window(x, idx, geo);
            const float4 g = ldg4(dyp + idx * 4);
            const uint32_t packed = __ldg(code + idx);
            float4 v[4];
            v[0] = ldx4(w.p00); v[1] = ldx4(w.p00 + C); v[2] = ldx4(w.p00 + w.rowstride); v[3] = ldx4(w.p00 + w.rowstride + C);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t cd = (packed >> (8 * k)) & 0xffu;
                const float gk = comp(g, k);
                const float sh = share<MODE>(gk, cd);
                fa[k] += gk;                                        // the routed shares of a window sum to gk
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float r = routed<MODE>(sh, cd, e);
                    fb[k] = fmaf(r, (comp(v[e], k) - mu[k]) * is[k], fb[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { da[k] += (double)fa[k]; db[k] += (double)fb[k]; }
    }
    __shared__ double sa[256][4], sb[256][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { sa[threadIdx.x][k] = da[k]; sb[threadIdx.x][k] = db[k]; }
    __syncthreads();
    if (threadIdx.x < Q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double ta = 0, tb = 0;
            for (int t = threadIdx.x; t < 256; t += Q) { ta += sa[t][k]; tb += sb[t][k]; }
            double *slot = scratch + (size_t)(blockIdx.x % BN_SLOTS) * 2 * C;
            atomicAdd(slot + cq * 4 + k, ta);
            atomicAdd(slot + C + cq * 4 + k, tb);
        }
    }
    bn_finalize_last_block(scratch, fin);
}

// The same two sums from the POOLED tensors alone.  Every winner of a window carries the pooled
// value P = gamma*xhat + beta, and the routed shares of a window add up to its pooled gradient g, so
//   sum(dy) = sum g          sum(dy * xhat) = sum g * (P - beta) / gamma
// with no look at the full-resolution map (4x the bytes of P) or at the codes.  Channels whose gamma
// is too small for that division to be accurate (|gamma| < max(1,|beta|)/16, including gamma == 0)
// take the window route above, per thread.
template <int MODE, typename XT>
__global__ void __launch_bounds__(256, 3) bn_pool_bwd_reduce_pooled_kernel(const float *__restrict__ dyp, const float *__restrict__ yp,
                                                                        const uint32_t *__restrict__ code, const XT *__restrict__ x,
                                                                        const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                        const float *__restrict__ mean, const float *__restrict__ invstd,
                                                                        double *__restrict__ scratch, long long windows4, int C, int W,
                                                                        int oh, int ow, BnFinal fin) {
    pdl_wait();
    const int Q = C >> 2;
    const WinGeo geo = win_geo(Q, ow, oh, W, C);
    const int cq = threadIdx.x % Q;
    float bt[4], rg[4];
    bool fast = true;
    {
        const float4 g = ldg4(gamma + cq * 4), be = ldg4(beta + cq * 4);
        bt[0] = be.x; bt[1] = be.y; bt[2] = be.z; bt[3] = be.w;
        const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            fast = fast && fabsf(gg[k]) * 16.f >= fmaxf(1.f, fabsf(bt[k]));
            rg[k] = 1.f / gg[k];
        }
    }
    // float64 running sums live in shared memory (they are folded into rarely): 16 registers less
    __shared__ double sa[256][4], sb[256][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { sa[threadIdx.x][k] = 0.0; sb[threadIdx.x][k] = 0.0; }
    const long long stride = (long long)gridDim.x * 256;
    long long idx = blockIdx.x * 256ll + threadIdx.x;
    while (idx < windows4) {
        float fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0};
        if (fast) {
            // 8 independent 16-byte loads in flight per thread, 8 such batches between float64 folds
            for (int it = 0; it < 8 && idx < windows4; ++it) {
                if (idx + 3 * stride < windows4) {
                    float4 g[4], pv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { g[u] = ldg4(dyp + (idx + u * stride) * 4); pv[u] = ldg4(yp + (idx + u * stride) * 4); }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float gk = comp(g[u], k);
                            fa[k] += gk;
                            fb[k] = fmaf(gk, (comp(pv[u], k) - bt[k]) * rg[k], fb[k]);
                        }
                    idx += 4 * stride;
                } else {
                    const float4 g = ldg4(dyp + idx * 4), pv = ldg4(yp + idx * 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float gk = comp(g, k);
                        fa[k] += gk;
                        fb[k] = fmaf(gk, (comp(pv, k) - bt[k]) * rg[k], fb[k]);
                    }
                    idx += stride;
                }
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < 16 && idx < windows4; ++it, idx += stride) {
                const WinT<XT> w = window(x, idx, geo);
                const float4 g = ldg4(dyp + idx * 4), mu4 = ldg4(mean + cq * 4), is4 = ldg4(invstd + cq * 4);
                const uint32_t packed = __ldg(code + idx);
                float4 v[4];
                v[0] = ldx4(w.p00); v[1] = ldx4(w.p00 + C); v[2] = ldx4(w.p00 + w.rowstride); v[3] = ldx4(w.p00 + w.rowstride + C);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t cd = (packed >> (8 * k)) & 0xffu;
                    const float gk = comp(g, k);
                    const float sh = share<MODE>(gk, cd);
                    fa[k] += gk;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float r = routed<MODE>(sh, cd, e);
                        fb[k] = fmaf(r, (comp(v[e], k) - comp(mu4, k)) * comp(is4, k), fb[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { sa[threadIdx.x][k] += (double)fa[k]; sb[threadIdx.x][k] += (double)fb[k]; }
    }
    __syncthreads();
    if (threadIdx.x < Q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double ta = 0, tb = 0;
            for (int t = threadIdx.x; t < 256; t += Q) { ta += sa[t][k]; tb += sb[t][k]; }
            double *slot = scratch + (size_t)(blockIdx.x % BN_SLOTS) * 2 * C;
            atomicAdd(slot + cq * 4 + k, ta);
            atomicAdd(slot + C + cq * 4 + k, tb);
        }
    }
    bn_finalize_last_block(scratch, fin);
}

template <int MODE, bool MASK, typename XT>
__global__ void __launch_bounds__(256, 4) bn_pool_bwd_apply_kernel(const float *__restrict__ dyp, const uint32_t *__restrict__ code,
                                                                const XT *__restrict__ x, float *__restrict__ dx,
                                                                __nv_bfloat16 *__restrict__ dx_bf16, float *__restrict__ colsum,
                                                                double *__restrict__ scratch, long long windows4, int C,
                                                                int W, int oh, int ow) {
    pdl_wait();
    const int Q = C >> 2;
    const WinGeo geo = win_geo(Q, ow, oh, W, C);
    const int cq = threadIdx.x % Q;
    const float *prm = bn_params(scratch, C);
    const float4 a4 = ldg4(prm + cq * 4), b4 = ldg4(prm + C + cq * 4), c4 = ldg4(prm + 2 * C + cq * 4);
    const float pa[4] = {a4.x, a4.y, a4.z, a4.w}, pb[4] = {b4.x, b4.y, b4.z, b4.w}, pc[4] = {c4.x, c4.y, c4.z, c4.w};
    float cs[4] = {0.f, 0.f, 0.f, 0.f};                  // this thread's share of sum_rows dx[., c] (the producer's bias gradient)
    const long long stride = (long long)gridDim.x * 256;
    for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < windows4; idx += stride) {
        const WinT<XT> w = window(x, idx, geo);
        const float4 g = ldg4(dyp + idx * 4);
        const uint32_t packed = __ldg(code + idx);
        float4 v[4], o[4];
        v[0] = ldx4(w.p00); v[1] = ldx4(w.p00 + C); v[2] = ldx4(w.p00 + w.rowstride); v[3] = ldx4(w.p00 + w.rowstride + C);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t cd = (packed >> (8 * k)) & 0xffu;
            const float gk = comp(g, k);
            const float sh = share<MODE>(gk, cd);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xv = comp(v[e], k);
                float r = fmaf(pa[k], routed<MODE>(sh, cd, e), fmaf(pb[k], xv, pc[k]));
                if (MASK && relu_blocked(xv)) r = 0.f;
                setc(o[e], k, r);
                cs[k] += r;
            }
        }
        const long long off = w.p00 - x;
        if (dx) {
            float *d00 = dx + off;
            *reinterpret_cast<float4 *>(d00) = o[0];
            *reinterpret_cast<float4 *>(d00 + C) = o[1];
            *reinterpret_cast<float4 *>(d00 + w.rowstride) = o[2];
            *reinterpret_cast<float4 *>(d00 + w.rowstride + C) = o[3];
        }
        if (dx_bf16) {
            __nv_bfloat16 *b00 = dx_bf16 + off;
            store_bf16x4(b00, o[0]);
            store_bf16x4(b00 + C, o[1]);
            store_bf16x4(b00 + w.rowstride, o[2]);
            store_bf16x4(b00 + w.rowstride + C, o[3]);
        }
    }
    if (colsum) {
        // block sums -> BN_SLOTS slot rows in the scratch (same-address atomics serialise at ~10 ns
        // each: one row would cost a chain of gridDim.x of them); colsum_fold_kernel adds the rows up.
        // (A last-block fold in here — fence, ticket, reload — cost 5.5 us of serial tail per launch.)
        __shared__ float scs[256][4];
        double *slots = scratch + 20 * (size_t)C;        // [COLSUM_SLOTS][C], float64: the result does not depend on the arrival order
#pragma unroll
        for (int k = 0; k < 4; ++k) scs[threadIdx.x][k] = cs[k];
        __syncthreads();
        if (threadIdx.x < Q) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = 0.f;
                for (int j = threadIdx.x; j < 256; j += Q) t += scs[j][k];
                atomicAdd(slots + (size_t)(blockIdx.x % COLSUM_SLOTS) * C + cq * 4 + k, (double)t);
            }
        }
    }
}

// adds the slot rows up and leaves them zero again for the next launch (they start zero: the scratch is
// allocated zeroed, and only this pair of kernels touches the region)
__global__ void colsum_fold_kernel(double *__restrict__ scratch, float *__restrict__ colsum, int C) {
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double *slots = scratch + 20 * (size_t)C;
    // loads first, in one batch (a load behind a store to the same array cannot be hoisted: 8 dependent L2 round trips)
    double v[COLSUM_SLOTS];
#pragma unroll
    for (int sl = 0; sl < COLSUM_SLOTS; ++sl) v[sl] = __ldcg(slots + (size_t)sl * C + c);
    const float old = colsum[c];
    double t = 0.0;
#pragma unroll
    for (int sl = 0; sl < COLSUM_SLOTS; ++sl) { t += v[sl]; slots[(size_t)sl * C + c] = 0.0; }
    colsum[c] = old + (float)t;
}

int grid_for_windows(long long windows4, int ctas_per_sm = 8) {
    long long need = (windows4 + 255) / 256;
    long long cap = (long long)num_sms() * ctas_per_sm;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

bool shape_ok(int N, int H, int W, int C) {
    if (N <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || (C & 3)) return false;
    const int q = C >> 2;
    return q >= 1 && q <= 256 && (q & (q - 1)) == 0;
}

}  // namespace

extern "C" {

int sn_bn_pool_supported(int N, int H, int W, int C) { return shape_ok(N, H, W, C) ? 1 : 0; }

int sn_bn_pool_fwd_train(const float *x, const float *gamma, const float *beta, float *y_pool, void *y_pool_bf16, uint8_t *code,
                         float *save_mean, float *save_invstd, float *moving_mean, float *moving_var, double *scratch,
                         int N, int H, int W, int C, float eps, float momentum, int pool_mode, void *stream) {
    SN_REQUIRE(x && gamma && beta && y_pool && code && save_mean && save_invstd && moving_mean && moving_var && scratch, "null tensor");
    SN_REQUIRE(shape_ok(N, H, W, C), "unsupported shape (needs even H, W and C = 4*2^k <= 1024)");
    SN_REQUIRE(pool_mode == SN_POOL_TIE_SPLIT || pool_mode == SN_POOL_FIRST_ARGMAX, "unknown pool mode");
    SN_REQUIRE(((((uintptr_t)x) | ((uintptr_t)y_pool)) & 15) == 0 && (((uintptr_t)code) & 3) == 0, "misaligned tensor");
    const long long rows = (long long)N * H * W;
    // pass 1: per-channel statistics of x (the plain BatchNorm reduction, scratch zeroed inside)
    int e = sn_batchnorm_stats(x, gamma, beta, save_mean, save_invstd, moving_mean, moving_var, scratch, rows, C, eps, momentum, stream);
    if (e) return e;
    return sn_bn_pool_fwd_apply(x, y_pool, y_pool_bf16, code, scratch, N, H, W, C, pool_mode, stream);
}

static int bn_pool_fwd_apply_impl(const void *x, bool xbf16, float *y_pool, void *y_pool_bf16, uint8_t *code, const double *scratch, int N,
                                  int H, int W, int C, int pool_mode, void *stream) {
    SN_REQUIRE(x && y_pool && code && scratch, "null tensor");
    SN_REQUIRE(shape_ok(N, H, W, C), "unsupported shape (needs even H, W and C = 4*2^k <= 1024)");
    SN_REQUIRE(pool_mode == SN_POOL_TIE_SPLIT || pool_mode == SN_POOL_FIRST_ARGMAX, "unknown pool mode");
    SN_REQUIRE(((((uintptr_t)x) | ((uintptr_t)y_pool)) & 15) == 0 && (((uintptr_t)code) & 3) == 0 && (((uintptr_t)y_pool_bf16) & 7) == 0,
               "misaligned tensor");
    cudaStream_t st = as_stream(stream);
    const int oh = H / 2, ow = W / 2;
    const long long windows4 = (long long)N * oh * ow * (C / 4);
    const int grid = grid_for_windows(windows4);
#define SN_FWD(M, XT) SN_CUDA(launch_dependent(bn_pool_fwd_kernel<M, XT>, dim3(grid), dim3(256), 0, st, (const XT *)x, y_pool, (__nv_bfloat16 *)y_pool_bf16, (uint32_t *)code, scratch, windows4, C, W, oh, ow))
    if (pool_mode == SN_POOL_TIE_SPLIT) { if (xbf16) SN_FWD(SN_POOL_TIE_SPLIT, __nv_bfloat16); else SN_FWD(SN_POOL_TIE_SPLIT, float); }
    else { if (xbf16) SN_FWD(SN_POOL_FIRST_ARGMAX, __nv_bfloat16); else SN_FWD(SN_POOL_FIRST_ARGMAX, float); }
#undef SN_FWD
    return after_launch("bn_pool_fwd_kernel");
}

int sn_bn_pool_fwd_apply(const float *x, float *y_pool, void *y_pool_bf16, uint8_t *code, const double *scratch, int N, int H,
                         int W, int C, int pool_mode, void *stream) {
    return bn_pool_fwd_apply_impl(x, false, y_pool, y_pool_bf16, code, scratch, N, H, W, C, pool_mode, stream);
}
int sn_bn_pool_fwd_apply_xbf16(const void *x_bf16, float *y_pool, void *y_pool_bf16, uint8_t *code, const double *scratch, int N, int H,
                               int W, int C, int pool_mode, void *stream) {
    return bn_pool_fwd_apply_impl(x_bf16, true, y_pool, y_pool_bf16, code, scratch, N, H, W, C, pool_mode, stream);
}

// phases: 1 = memset + reductions (with `defer`: local sums stay in the scratch), 2 = apply
static int bn_pool_bwd_impl(int phases, int defer, bool xbf16, const float *dy_pool, const uint8_t *code, const void *x, const float *y_pool,
                            const float *gamma, const float *beta, const float *save_mean, const float *save_invstd, float *dx,
                            void *dx_bf16, float *dx_colsum, float *dgamma, float *dbeta, double *scratch, int N, int H, int W,
                            int C, int pool_mode, int mask_relu, void *stream) {
    SN_REQUIRE(dy_pool && code && x && scratch, "null tensor");
    SN_REQUIRE(shape_ok(N, H, W, C), "unsupported shape (needs even H, W and C = 4*2^k <= 1024)");
    SN_REQUIRE(pool_mode == SN_POOL_TIE_SPLIT || pool_mode == SN_POOL_FIRST_ARGMAX, "unknown pool mode");
    SN_REQUIRE(((((uintptr_t)x) | ((uintptr_t)dx) | ((uintptr_t)dy_pool)) & 15) == 0 && (((uintptr_t)dx_bf16) & 7) == 0, "misaligned tensor");
    cudaStream_t st = as_stream(stream);
    const long long rows = (long long)N * H * W;
    const int oh = H / 2, ow = W / 2;
    const long long windows4 = (long long)N * oh * ow * (C / 4);
    const int grid = grid_for_windows(windows4);
    const uint32_t *cd = (const uint32_t *)code;
    if (phases & 1) {
        SN_REQUIRE(gamma && save_mean && save_invstd, "null tensor");
        // fewer, longer-lived reduce blocks on small maps (see bn.cu: bn_reduce_ctas)
        int rgrid = grid_for_windows(windows4, rows * C * 4 >= (100ll << 20) ? 8 : 3);
        SN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * BN_SCRATCH_DOUBLES_PER_C * C, st));
        BnFinal fin{};
        fin.bwd = 1; fin.defer = defer; fin.rows = rows; fin.C = C; fin.gamma = gamma; fin.mean = save_mean; fin.invstd = save_invstd;
        fin.dgamma = dgamma; fin.dbeta = dbeta;
        if (y_pool && beta) {
            SN_REQUIRE((((uintptr_t)y_pool) & 15) == 0, "misaligned tensor");
            // one resident wave (3 CTAs/SM by registers); on small maps at least 8 windows per thread, so
            // that the per-block tail (shared-memory fold, 2*C float64 atomics, ticket) stays amortised
            rgrid = grid_for_windows(windows4, 3);
            const long long by_work = windows4 / (256 * 8);
            if (by_work < rgrid) rgrid = by_work < 1 ? 1 : (int)by_work;
#define SN_RP(M, XT) bn_pool_bwd_reduce_pooled_kernel<M, XT><<<rgrid, 256, 0, st>>>(dy_pool, y_pool, cd, (const XT *)x, gamma, beta, save_mean, save_invstd, scratch, windows4, C, W, oh, ow, fin)
            if (pool_mode == SN_POOL_TIE_SPLIT) { if (xbf16) SN_RP(SN_POOL_TIE_SPLIT, __nv_bfloat16); else SN_RP(SN_POOL_TIE_SPLIT, float); }
            else { if (xbf16) SN_RP(SN_POOL_FIRST_ARGMAX, __nv_bfloat16); else SN_RP(SN_POOL_FIRST_ARGMAX, float); }
#undef SN_RP
        } else {
#define SN_RW(M, XT) bn_pool_bwd_reduce_kernel<M, XT><<<rgrid, 256, 0, st>>>(dy_pool, cd, (const XT *)x, save_mean, save_invstd, scratch, windows4, C, W, oh, ow, fin)
            if (pool_mode == SN_POOL_TIE_SPLIT) { if (xbf16) SN_RW(SN_POOL_TIE_SPLIT, __nv_bfloat16); else SN_RW(SN_POOL_TIE_SPLIT, float); }
            else { if (xbf16) SN_RW(SN_POOL_FIRST_ARGMAX, __nv_bfloat16); else SN_RW(SN_POOL_FIRST_ARGMAX, float); }
#undef SN_RW
        }
        int e = after_launch("bn_pool_bwd_reduce_kernel");
        if (e) return e;
    }
    if (phases & 2) {
        SN_REQUIRE(dx || dx_bf16, "null tensor");
#define SN_LAUNCH2(M, K, XT) SN_CUDA(launch_dependent(bn_pool_bwd_apply_kernel<M, K, XT>, dim3(grid), dim3(256), 0, st, dy_pool, cd, (const XT *)x, dx, (__nv_bfloat16 *)dx_bf16, dx_colsum, scratch, windows4, C, W, oh, ow))
#define SN_LAUNCH(M, K) do { if (xbf16) SN_LAUNCH2(M, K, __nv_bfloat16); else SN_LAUNCH2(M, K, float); } while (0)
        if (pool_mode == SN_POOL_TIE_SPLIT) { if (mask_relu) SN_LAUNCH(SN_POOL_TIE_SPLIT, true); else SN_LAUNCH(SN_POOL_TIE_SPLIT, false); }
        else { if (mask_relu) SN_LAUNCH(SN_POOL_FIRST_ARGMAX, true); else SN_LAUNCH(SN_POOL_FIRST_ARGMAX, false); }
#undef SN_LAUNCH2
#undef SN_LAUNCH
        int e = after_launch("bn_pool_bwd_apply_kernel");
        if (e || !dx_colsum) return e;
        SN_CUDA(launch_dependent(colsum_fold_kernel, dim3((C + 255) / 256), dim3(256), 0, st, scratch, dx_colsum, C));
        return after_launch("colsum_fold_kernel");
    }
    return SN_OK;
}

int sn_bn_pool_bwd(const float *dy_pool, const uint8_t *code, const float *x, const float *y_pool, const float *gamma,
                   const float *beta, const float *save_mean, const float *save_invstd, float *dx, void *dx_bf16,
                   float *dx_colsum, float *dgamma, float *dbeta, double *scratch, int N, int H, int W, int C, int pool_mode,
                   int mask_relu, void *stream) {
    return bn_pool_bwd_impl(3, 0, false, dy_pool, code, x, y_pool, gamma, beta, save_mean, save_invstd, dx, dx_bf16, dx_colsum, dgamma, dbeta,
                            scratch, N, H, W, C, pool_mode, mask_relu, stream);
}

int sn_bn_pool_bwd_local(const float *dy_pool, const uint8_t *code, const float *x, const float *y_pool, const float *gamma,
                         const float *beta, const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta,
                         double *scratch, int N, int H, int W, int C, int pool_mode, void *stream) {
    return bn_pool_bwd_impl(1, 1, false, dy_pool, code, x, y_pool, gamma, beta, save_mean, save_invstd, nullptr, nullptr, nullptr, dgamma, dbeta,
                            scratch, N, H, W, C, pool_mode, 0, stream);
}

int sn_bn_pool_bwd_apply(const float *dy_pool, const uint8_t *code, const float *x, float *dx, void *dx_bf16, float *dx_colsum,
                         double *scratch, int N, int H, int W, int C, int pool_mode, int mask_relu, void *stream) {
    return bn_pool_bwd_impl(2, 0, false, dy_pool, code, x, nullptr, nullptr, nullptr, nullptr, nullptr, dx, dx_bf16, dx_colsum, nullptr, nullptr,
                            scratch, N, H, W, C, pool_mode, mask_relu, stream);
}

// the same three with the full-resolution map x stored in bf16 (the bf16 tier's convolution outputs)
int sn_bn_pool_bwd_xbf16(const float *dy_pool, const uint8_t *code, const void *x_bf16, const float *y_pool, const float *gamma,
                         const float *beta, const float *save_mean, const float *save_invstd, float *dx, void *dx_bf16,
                         float *dx_colsum, float *dgamma, float *dbeta, double *scratch, int N, int H, int W, int C, int pool_mode,
                         int mask_relu, void *stream) {
    return bn_pool_bwd_impl(3, 0, true, dy_pool, code, x_bf16, y_pool, gamma, beta, save_mean, save_invstd, dx, dx_bf16, dx_colsum, dgamma, dbeta,
                            scratch, N, H, W, C, pool_mode, mask_relu, stream);
}

int sn_bn_pool_bwd_local_xbf16(const float *dy_pool, const uint8_t *code, const void *x_bf16, const float *y_pool, const float *gamma,
                               const float *beta, const float *save_mean, const float *save_invstd, float *dgamma, float *dbeta,
                               double *scratch, int N, int H, int W, int C, int pool_mode, void *stream) {
    return bn_pool_bwd_impl(1, 1, true, dy_pool, code, x_bf16, y_pool, gamma, beta, save_mean, save_invstd, nullptr, nullptr, nullptr, dgamma, dbeta,
                            scratch, N, H, W, C, pool_mode, 0, stream);
}

int sn_bn_pool_bwd_apply_xbf16(const float *dy_pool, const uint8_t *code, const void *x_bf16, float *dx, void *dx_bf16, float *dx_colsum,
                               double *scratch, int N, int H, int W, int C, int pool_mode, int mask_relu, void *stream) {
    return bn_pool_bwd_impl(2, 0, true, dy_pool, code, x_bf16, nullptr, nullptr, nullptr, nullptr, nullptr, dx, dx_bf16, dx_colsum, nullptr, nullptr,
                            scratch, N, H, W, C, pool_mode, mask_relu, stream);
}

}  // extern "C"
