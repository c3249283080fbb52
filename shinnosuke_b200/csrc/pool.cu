// MaxPooling2D forward / backward on NHWC maps (layers/Convolution.py:254-362).
// HBM-bound: 128-bit loads/stores over the channel dimension, one byte of "code" per output
// element instead of the reference's cached float64 window tensor / int64 argmax vector.
#include "common.cuh"
using namespace sn;

template <int V> struct Vec;
template <> struct Vec<4> { typedef float4 T; };
template <> struct Vec<1> { typedef float T; };

__device__ __forceinline__ float lane(const float4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ float lane(const float &v, int) { return v; }
__device__ __forceinline__ void setlane(float4 &v, int k, float x) { if (k == 0) v.x = x; else if (k == 1) v.y = x; else if (k == 2) v.z = x; else v.w = x; }
__device__ __forceinline__ void setlane(float &v, int, float x) { v = x; }

// One thread per (output pixel, V channels).  CODE_T is uint8_t (pool*pool <= 8 in tie mode, or
// any first-argmax pool <= 16) or uint16_t (tie mode with 9..16 window elements).
template <int V, typename CODE_T, int MODE>
__global__ void maxpool_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, CODE_T *__restrict__ code,
                                   int N, int H, int W, int C, int oh, int ow, int pool, int stride) {
    typedef typename Vec<V>::T VT;
    const int CV = C / V;
    const long long total = (long long)N * oh * ow * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int j = (int)(pix % ow);
        int i = (int)((pix / ow) % oh);
        int n = (int)(pix / ((long long)ow * oh));
        const float *base = x + (((long long)n * H + (long long)i * stride) * W + (long long)j * stride) * C + (long long)cv * V;
        VT best;
        unsigned codes[V];
#pragma unroll
        for (int k = 0; k < V; ++k) { setlane(best, k, -INFINITY); codes[k] = 0; }
        // pass 1: maximum and first argmax
        for (int u = 0; u < pool; ++u)
            for (int v = 0; v < pool; ++v) {
                VT val = *reinterpret_cast<const VT *>(base + ((long long)u * W + v) * C);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    float a = lane(val, k);
                    if (a > lane(best, k) || (u == 0 && v == 0)) { setlane(best, k, a); if (MODE == SN_POOL_FIRST_ARGMAX) codes[k] = u * pool + v; }
                }
            }
        if (MODE == SN_POOL_TIE_SPLIT && code) {
            // pass 2 (L1-resident re-read): bit mask of every element equal to the maximum
            for (int u = 0; u < pool; ++u)
                for (int v = 0; v < pool; ++v) {
                    VT val = *reinterpret_cast<const VT *>(base + ((long long)u * W + v) * C);
#pragma unroll
                    for (int k = 0; k < V; ++k)
                        if (lane(val, k) == lane(best, k)) codes[k] |= 1u << (u * pool + v);
                }
        }
        *reinterpret_cast<VT *>(y + idx * V) = best;
        if (code) {
#pragma unroll
            for (int k = 0; k < V; ++k) code[idx * V + k] = (CODE_T)codes[k];
        }
    }
}

// Gather form: one thread per (INPUT pixel, V channels); handles overlapping windows and the
// rows/columns beyond the floor (they get 0), writes are fully coalesced.
template <int V, typename CODE_T, int MODE, bool ACC>
__global__ void maxpool_bwd_kernel(const float *__restrict__ dy, const CODE_T *__restrict__ code, float *__restrict__ dx,
                                   int N, int H, int W, int C, int oh, int ow, int pool, int stride) {
    typedef typename Vec<V>::T VT;
    const int CV = C / V;
    const long long total = (long long)N * H * W * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int w = (int)(pix % W);
        int h = (int)((pix / W) % H);
        int n = (int)(pix / ((long long)W * H));
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        int i_lo = h - pool + 1; i_lo = i_lo <= 0 ? 0 : (i_lo + stride - 1) / stride;
        int j_lo = w - pool + 1; j_lo = j_lo <= 0 ? 0 : (j_lo + stride - 1) / stride;
        int i_hi = min(h / stride, oh - 1), j_hi = min(w / stride, ow - 1);
        for (int i = i_lo; i <= i_hi; ++i)
            for (int j = j_lo; j <= j_hi; ++j) {
                unsigned bit = (unsigned)((h - i * stride) * pool + (w - j * stride));
                long long o = ((((long long)n * oh + i) * ow + j) * C) + (long long)cv * V;
                VT g = *reinterpret_cast<const VT *>(dy + o);
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    unsigned cd = code[o + k];
                    if (MODE == SN_POOL_FIRST_ARGMAX) { if (cd == bit) acc[k] += lane(g, k); }
                    else if ((cd >> bit) & 1u) acc[k] += lane(g, k) / (float)__popc(cd);
                }
            }
        VT out;
        if (ACC) out = *reinterpret_cast<VT *>(dx + idx * V);
#pragma unroll
        for (int k = 0; k < V; ++k) setlane(out, k, ACC ? lane(out, k) + acc[k] : acc[k]);
        *reinterpret_cast<VT *>(dx + idx * V) = out;
    }
}


// ------------------------------------------------------------------------------------------
// Fast path: 2x2 windows, stride 2, C % 4 == 0 (every BASELINE config).  One thread per
// (output pixel, 4 channels): 128-bit loads of the four window rows, one 128-bit store of y and
// ONE 32-bit store of the four code bytes.  The backward is the mirror image: one 128-bit load of
// dY, one 32-bit load of the codes, four 128-bit stores of dX; odd trailing rows/columns (floor
// semantics) are zeroed by the threads of the last window row/column.
template <int MODE>
__global__ void __launch_bounds__(256) maxpool2x2_fwd_kernel(const float *__restrict__ x, float *__restrict__ y,
                                                             uint32_t *__restrict__ code, int N, int H, int W, int C,
                                                             int oh, int ow) {
    const int CV = C >> 2;
    const long long total = (long long)N * oh * ow * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int j = (int)(pix % ow);
        long long t = pix / ow;
        int i = (int)(t % oh);
        long long n = t / oh;
        const float *base = x + ((n * H + 2 * i) * W + 2 * j) * C + cv * 4;
        float4 v[4];
        v[0] = ldg4(base); v[1] = ldg4(base + C); v[2] = ldg4(base + (long long)W * C); v[3] = ldg4(base + (long long)W * C + C);
        float4 out;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a0 = lane(v[0], k), a1 = lane(v[1], k), a2 = lane(v[2], k), a3 = lane(v[3], k);
            float m = a0; uint32_t am = 0;
            if (a1 > m) { m = a1; am = 1; }
            if (a2 > m) { m = a2; am = 2; }
            if (a3 > m) { m = a3; am = 3; }
            uint32_t cd;
            if (MODE == SN_POOL_FIRST_ARGMAX) cd = am;
            else cd = (a0 == m ? 1u : 0u) | (a1 == m ? 2u : 0u) | (a2 == m ? 4u : 0u) | (a3 == m ? 8u : 0u);
            setlane(out, k, m);
            packed |= cd << (8 * k);
        }
        *reinterpret_cast<float4 *>(y + idx * 4) = out;
        if (code) code[idx] = packed;
    }
}

template <int MODE, bool ACC>
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const float *__restrict__ dy, const uint32_t *__restrict__ code,
                                                             float *__restrict__ dx, int N, int H, int W, int C, int oh,
                                                             int ow) {
    const int CV = C >> 2;
    const long long total = (long long)N * oh * ow * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv = (int)(idx % CV);
        long long pix = idx / CV;
        int j = (int)(pix % ow);
        long long t = pix / ow;
        int i = (int)(t % oh);
        long long n = t / oh;
        const float4 g = ldg4(dy + idx * 4);
        const uint32_t packed = __ldg(code + idx);
        float4 o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t cd = (packed >> (8 * k)) & 0xffu;
            const float gv = lane(g, k);
            float share = gv;
            if (MODE == SN_POOL_TIE_SPLIT) { int cnt = __popc(cd); if (cnt > 1) share = gv / (float)cnt; }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                bool hit = (MODE == SN_POOL_FIRST_ARGMAX) ? (cd == (uint32_t)e) : ((cd >> e) & 1u);
                setlane(o[e], k, hit ? share : 0.f);
            }
        }
        float *base = dx + ((n * H + 2 * i) * W + 2 * j) * C + cv * 4;
        float *ptrs[4] = {base, base + C, base + (long long)W * C, base + (long long)W * C + C};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float4 *d = reinterpret_cast<float4 *>(ptrs[e]);
            if (ACC) { float4 old = *d; o[e].x += old.x; o[e].y += old.y; o[e].z += old.z; o[e].w += old.w; }
            *d = o[e];
        }
        if (!ACC) {
            // floor semantics: inputs beyond the last full window get no gradient
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((W & 1) && j == ow - 1) {
                *reinterpret_cast<float4 *>(base + 2 * C) = z;
                *reinterpret_cast<float4 *>(base + (long long)W * C + 2 * C) = z;
            }
            if ((H & 1) && i == oh - 1) {
                *reinterpret_cast<float4 *>(base + 2ll * W * C) = z;
                *reinterpret_cast<float4 *>(base + 2ll * W * C + C) = z;
                if ((W & 1) && j == ow - 1) *reinterpret_cast<float4 *>(base + 2ll * W * C + 2 * C) = z;
            }
        }
    }
}

template <int V, typename CODE_T>
static int launch_fwd(const float *x, float *y, void *code, int N, int H, int W, int C, int oh, int ow, int pool,
                      int stride, int mode, cudaStream_t st) {
    long long total = (long long)N * oh * ow * (C / V);
    int grid = grid_for(total, 256);
    if (mode == SN_POOL_TIE_SPLIT)
        maxpool_fwd_kernel<V, CODE_T, SN_POOL_TIE_SPLIT><<<grid, 256, 0, st>>>(x, y, (CODE_T *)code, N, H, W, C, oh, ow, pool, stride);
    else
        maxpool_fwd_kernel<V, CODE_T, SN_POOL_FIRST_ARGMAX><<<grid, 256, 0, st>>>(x, y, (CODE_T *)code, N, H, W, C, oh, ow, pool, stride);
    return after_launch("maxpool_fwd_kernel");
}
template <int V, typename CODE_T>
static int launch_bwd(const float *dy, const void *code, float *dx, int N, int H, int W, int C, int oh, int ow,
                      int pool, int stride, int mode, int acc, cudaStream_t st) {
    long long total = (long long)N * H * W * (C / V);
    int grid = grid_for(total, 256);
    const CODE_T *cd = (const CODE_T *)code;
#define SN_LAUNCH(M, A) maxpool_bwd_kernel<V, CODE_T, M, A><<<grid, 256, 0, st>>>(dy, cd, dx, N, H, W, C, oh, ow, pool, stride)
    if (mode == SN_POOL_TIE_SPLIT) { if (acc) SN_LAUNCH(SN_POOL_TIE_SPLIT, true); else SN_LAUNCH(SN_POOL_TIE_SPLIT, false); }
    else { if (acc) SN_LAUNCH(SN_POOL_FIRST_ARGMAX, true); else SN_LAUNCH(SN_POOL_FIRST_ARGMAX, false); }
#undef SN_LAUNCH
    return after_launch("maxpool_bwd_kernel");
}

static int pool_args_ok(int N, int H, int W, int C, int pool, int stride, int mode) {
    if (N < 0 || H <= 0 || W <= 0 || C <= 0 || pool <= 0 || stride <= 0 || pool > H || pool > W) {
        set_error("maxpool: bad shape N=%d H=%d W=%d C=%d pool=%d stride=%d", N, H, W, C, pool, stride);
        return SN_ERR_INVALID_ARG;
    }
    if (mode != SN_POOL_TIE_SPLIT && mode != SN_POOL_FIRST_ARGMAX) { set_error("maxpool: unknown mode %d", mode); return SN_ERR_INVALID_ARG; }
    if (pool > 16 || (mode == SN_POOL_TIE_SPLIT && pool * pool > 16)) {
        set_error("maxpool: pool=%d not supported in mode %d (tie masks hold up to 16 window elements)", pool, mode);
        return SN_ERR_UNSUPPORTED;
    }
    return SN_OK;
}

extern "C" {

int sn_maxpool2d_fwd(const float *x, float *y, uint8_t *code, int N, int H, int W, int C, int pool, int stride,
                     int mode, void *stream) {
    SN_REQUIRE(x && y, "null tensor");
    int e = pool_args_ok(N, H, W, C, pool, stride, mode);
    if (e) return e;
    if (N == 0) return SN_OK;
    int oh = (H - pool) / stride + 1, ow = (W - pool) / stride + 1;
    bool wide = (mode == SN_POOL_TIE_SPLIT && pool * pool > 8);
    bool vec = (C % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
    cudaStream_t st = as_stream(stream);
    if (vec && pool == 2 && stride == 2 && (((uintptr_t)code) & 3) == 0) {
        int grid = grid_for((long long)N * oh * ow * (C / 4), 256);
        if (mode == SN_POOL_TIE_SPLIT) maxpool2x2_fwd_kernel<SN_POOL_TIE_SPLIT><<<grid, 256, 0, st>>>(x, y, (uint32_t *)code, N, H, W, C, oh, ow);
        else maxpool2x2_fwd_kernel<SN_POOL_FIRST_ARGMAX><<<grid, 256, 0, st>>>(x, y, (uint32_t *)code, N, H, W, C, oh, ow);
        return after_launch("maxpool2x2_fwd_kernel");
    }
    if (vec) return wide ? launch_fwd<4, uint16_t>(x, y, code, N, H, W, C, oh, ow, pool, stride, mode, st)
                         : launch_fwd<4, uint8_t>(x, y, code, N, H, W, C, oh, ow, pool, stride, mode, st);
    return wide ? launch_fwd<1, uint16_t>(x, y, code, N, H, W, C, oh, ow, pool, stride, mode, st)
                : launch_fwd<1, uint8_t>(x, y, code, N, H, W, C, oh, ow, pool, stride, mode, st);
}

int sn_maxpool2d_bwd(const float *dy, const uint8_t *code, float *dx, int N, int H, int W, int C, int pool,
                     int stride, int mode, int accumulate, void *stream) {
    SN_REQUIRE(dy && code && dx, "null tensor");
    int e = pool_args_ok(N, H, W, C, pool, stride, mode);
    if (e) return e;
    if (N == 0) return SN_OK;
    int oh = (H - pool) / stride + 1, ow = (W - pool) / stride + 1;
    bool wide = (mode == SN_POOL_TIE_SPLIT && pool * pool > 8);
    bool vec = (C % 4 == 0) && (((uintptr_t)dx | (uintptr_t)dy) & 15) == 0;
    cudaStream_t st = as_stream(stream);
    if (vec && pool == 2 && stride == 2 && (((uintptr_t)code) & 3) == 0) {
        int grid = grid_for((long long)N * oh * ow * (C / 4), 256);
        const uint32_t *cd = (const uint32_t *)code;
#define SN_LAUNCH2(M, A) maxpool2x2_bwd_kernel<M, A><<<grid, 256, 0, st>>>(dy, cd, dx, N, H, W, C, oh, ow)
        if (mode == SN_POOL_TIE_SPLIT) { if (accumulate) SN_LAUNCH2(SN_POOL_TIE_SPLIT, true); else SN_LAUNCH2(SN_POOL_TIE_SPLIT, false); }
        else { if (accumulate) SN_LAUNCH2(SN_POOL_FIRST_ARGMAX, true); else SN_LAUNCH2(SN_POOL_FIRST_ARGMAX, false); }
#undef SN_LAUNCH2
        return after_launch("maxpool2x2_bwd_kernel");
    }
    if (vec) return wide ? launch_bwd<4, uint16_t>(dy, code, dx, N, H, W, C, oh, ow, pool, stride, mode, accumulate, st)
                         : launch_bwd<4, uint8_t>(dy, code, dx, N, H, W, C, oh, ow, pool, stride, mode, accumulate, st);
    return wide ? launch_bwd<1, uint16_t>(dy, code, dx, N, H, W, C, oh, ow, pool, stride, mode, accumulate, st)
                : launch_bwd<1, uint8_t>(dy, code, dx, N, H, W, C, oh, ow, pool, stride, mode, accumulate, st);
}

}  // extern "C"
