// Direct convolution for tiny channel counts (C*kh*kw <= ~64: config 1's 1-channel 2x2 filter,
// config 3's RGB stem).  The contraction is far too short for the tensor pipe (K = 4 or 27) and
// the arithmetic intensity is a few flops per byte, so these kernels are judged against HBM
// bandwidth: the input is read once through L1/L2, the weights live in shared memory, the output
// (or dY) is streamed with 128-bit fully coalesced accesses, and the FMAs are register-blocked so
// the CUDA cores keep up with the memory stream.  fp32 throughout (used by every precision tier).
#include "common.cuh"
using namespace sn;

namespace {

struct DirectP {
    const float *x, *w, *bias, *dy;
    float *y, *dw, *db;
    int N, C, H, W, F, kh, kw, stride, ph, pw, oh, ow;
    int K;            // C*kh*kw
    long long P;      // N*oh*ow
    int act;
};

// ---------------------------------------------------------------------------------- fprop
// lane -> (pixel slot, filter group of FG filters); every thread produces PPT pixels x FG filters.
// A warp writes 32/TPP consecutive pixels x F floats = one contiguous run of the NHWC output.
template <int FG, int PPT>
__global__ void __launch_bounds__(256) direct_fprop_kernel(DirectP p) {
    extern __shared__ float ws[];                     // ws[k][filter group][FG + 4]: the pad keeps the
    const int tpp = p.F / FG;                          // groups of one LDS.128 phase on distinct banks
    const int rs = tpp * (FG + 4);
    for (int i = threadIdx.x; i < p.K * p.F; i += blockDim.x) {
        int k = i / p.F, f = i % p.F;
        ws[k * rs + (f / FG) * (FG + 4) + (f % FG)] = __ldg(p.w + (long long)f * p.K + k);
    }
    __syncthreads();
    const int ppw = 32 / tpp;                          // pixels per warp per pass
    const int lane = threadIdx.x & 31;
    const int fg = lane % tpp, slot = lane / tpp;
    // 32-bit index arithmetic throughout (the host guarantees P < 2^31): 64-bit divisions would cost
    // more instructions than the FMAs of a pixel
    const unsigned warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const unsigned pix_per_iter = ppw * PPT;
    const unsigned P = (unsigned)p.P, OW = p.ow, OH = p.oh;
    float bias[FG];
#pragma unroll
    for (int j = 0; j < FG; ++j) bias[j] = p.bias ? __ldg(p.bias + fg * FG + j) : 0.f;

    for (unsigned base = warp_global * pix_per_iter; base < P; base += nwarps * pix_per_iter) {
        float acc[PPT][FG];
        const float *xb[PPT];
        int hb[PPT], wb[PPT];
        bool ok[PPT];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            unsigned pix = base + q * ppw + slot;
            ok[q] = pix < P;
            unsigned pp = ok[q] ? pix : 0u;
            unsigned t = pp / OW, j = pp - t * OW;
            unsigned n = t / OH, i = t - n * OH;
            hb[q] = (int)i * p.stride - p.ph;
            wb[q] = (int)j * p.stride - p.pw;
            xb[q] = p.x + (long long)n * p.H * p.W * p.C;
#pragma unroll
            for (int f = 0; f < FG; ++f) acc[q][f] = bias[f];
        }
        int k = 0;
        for (int u = 0; u < p.kh; ++u)
            for (int v = 0; v < p.kw; ++v)
                for (int c = 0; c < p.C; ++c, ++k) {
                    float xv[PPT];
#pragma unroll
                    for (int q = 0; q < PPT; ++q) {
                        int h = hb[q] + u, w = wb[q] + v;
                        xv[q] = ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W)
                                    ? __ldg(xb[q] + (h * p.W + w) * p.C + c) : 0.f;
                    }
                    const float4 *wr = reinterpret_cast<const float4 *>(ws + k * rs + fg * (FG + 4));
#pragma unroll
                    for (int f4 = 0; f4 < FG / 4; ++f4) {
                        float4 wv = wr[f4];
#pragma unroll
                        for (int q = 0; q < PPT; ++q) {
                            acc[q][f4 * 4 + 0] = fmaf(xv[q], wv.x, acc[q][f4 * 4 + 0]);
                            acc[q][f4 * 4 + 1] = fmaf(xv[q], wv.y, acc[q][f4 * 4 + 1]);
                            acc[q][f4 * 4 + 2] = fmaf(xv[q], wv.z, acc[q][f4 * 4 + 2]);
                            acc[q][f4 * 4 + 3] = fmaf(xv[q], wv.w, acc[q][f4 * 4 + 3]);
                        }
                    }
                }
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
            if (!ok[q]) continue;
            unsigned pix = base + q * ppw + slot;
            float4 *dst = reinterpret_cast<float4 *>(p.y + (long long)pix * p.F + fg * FG);
#pragma unroll
            for (int f4 = 0; f4 < FG / 4; ++f4) {
                float4 o = make_float4(acc[q][f4 * 4], acc[q][f4 * 4 + 1], acc[q][f4 * 4 + 2], acc[q][f4 * 4 + 3]);
                if (p.act == SN_ACT_RELU) { o.x = relu_keep_sign(o.x); o.y = relu_keep_sign(o.y); o.z = relu_keep_sign(o.z); o.w = relu_keep_sign(o.w); }
                dst[f4] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------- wgrad
// dW[f][k] += sum_p dY[p][f] * xcol[p][k] over a pixel range per block, K padded to KP = 32 with
// a constant-one column at index K so that the bias gradient sum_p dY[p][f] falls out as dW[f][K].
// 128 threads: thread -> (4 filters, 4 k's) = 16 accumulators; per pixel 2 LDS.128 feed 16 FMAs.
constexpr int WG_PT = 64;     // pixels per shared-memory tile
constexpr int WG_KP = 32;
constexpr int WG_FT = 64;     // filters per block

__global__ void __launch_bounds__(128) direct_wgrad_kernel(DirectP p, int pix_per_block) {
    __shared__ __align__(16) float dys[WG_PT][WG_FT];
    __shared__ __align__(16) float xcs[WG_PT][WG_KP];
    __shared__ int row_h[WG_PT], row_w[WG_PT], row_off[WG_PT];   // per tile row: i*s-ph, j*s-pw, n*H*W*C (or -1)
    const int t = threadIdx.x;
    const int f4 = (t & 15) * 4, k4 = (t >> 4) * 4;
    const int f0 = blockIdx.y * WG_FT;
    const int P = (int)p.P;
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(P, p_begin + pix_per_block);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // this thread always gathers column k = t % 32 of the column tile: decode (u, v, c) once
    const int gk = t & 31;
    const int kwC = p.kw * p.C;
    const int gu = gk / kwC, gv = (gk % kwC) / p.C, gc = gk % p.C;
    const int HWC = p.H * p.W * p.C;

    for (int pt = p_begin; pt < p_end; pt += WG_PT) {
        if (t < WG_PT) {
            int pix = pt + t;
            if (pix < p_end) {
                int tt = pix / p.ow, j = pix - tt * p.ow;
                int n = tt / p.oh, i = tt - n * p.oh;
                row_h[t] = i * p.stride - p.ph; row_w[t] = j * p.stride - p.pw; row_off[t] = n * HWC;
            } else {
                row_off[t] = -1;
            }
        }
        // dY tile: WG_PT x 64 floats, 128-bit coalesced
#pragma unroll
        for (int r = 0; r < (WG_PT * WG_FT / 4) / 128; ++r) {
            int idx = r * 128 + t;
            int row = idx / (WG_FT / 4), c4 = (idx % (WG_FT / 4)) * 4;
            int pix = pt + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pix < p_end) v = ldg4(p.dy + (long long)pix * p.F + f0 + c4);
            *reinterpret_cast<float4 *>(&dys[row][c4]) = v;
        }
        __syncthreads();
        // column tile: xcs[row][k] = x[n, i*s+u-ph, j*s+v-pw, c], k = (u*kw+v)*C + c; xcs[row][K] = 1
#pragma unroll
        for (int r = 0; r < (WG_PT * WG_KP) / 128; ++r) {
            int row = r * 4 + (t >> 5);
            float v = 0.f;
            int off = row_off[row];
            if (off >= 0) {
                if (gk < p.K) {
                    int h = row_h[row] + gu, w = row_w[row] + gv;
                    if ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W)
                        v = __ldg(p.x + (long long)off + (h * p.W + w) * p.C + gc);
                } else if (gk == p.K) {
                    v = 1.f;
                }
            }
            xcs[row][gk] = v;
        }
        __syncthreads();
#pragma unroll 8
        for (int row = 0; row < WG_PT; ++row) {
            float4 d = *reinterpret_cast<const float4 *>(&dys[row][f4]);
            float4 xk = *reinterpret_cast<const float4 *>(&xcs[row][k4]);
            float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {xk.x, xk.y, xk.z, xk.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int f = f0 + f4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k4 + j;
            if (k < p.K) atomicAdd(p.dw + (long long)f * p.K + k, acc[i][j]);
            else if (k == p.K && p.db) atomicAdd(p.db + f, acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------------------------- tiny wgrad
// F <= FT filters and K <= KT taps (config 1: F = 8, K = 4): dW and db together are a few dozen
// numbers, so every thread keeps ALL of them in registers and streams pixels (two 128-bit loads of
// dY, K scalar gathers of x per pixel); warps fold with shuffles, the block through shared memory,
// one atomic per output per block.  One wave of blocks: each output sees <= #SM same-address atomics.
template <int FT, int KT>
__global__ void __launch_bounds__(256) tiny_wgrad_kernel(DirectP p) {
    float acc[FT][KT + 1];
#pragma unroll
    for (int f = 0; f < FT; ++f)
#pragma unroll
        for (int k = 0; k <= KT; ++k) acc[f][k] = 0.f;
    const int P = (int)p.P, kwC = p.kw * p.C;
    const bool vec = (p.F & 3) == 0 && (((uintptr_t)p.dy) & 15) == 0;
    for (int pix = blockIdx.x * 256 + threadIdx.x; pix < P; pix += gridDim.x * 256) {
        const int t = pix / p.ow, j = pix - t * p.ow, n = t / p.oh, i = t - n * p.oh;
        float d[FT];
        const float *dyp = p.dy + (long long)pix * p.F;
        if (vec) {
#pragma unroll
            for (int f = 0; f < FT; f += 4) {
                float4 v = f < p.F ? ldg4(dyp + f) : make_float4(0.f, 0.f, 0.f, 0.f);
                d[f] = v.x; d[f + 1] = v.y; d[f + 2] = v.z; d[f + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int f = 0; f < FT; ++f) d[f] = f < p.F ? __ldg(dyp + f) : 0.f;
        }
        const int hb = i * p.stride - p.ph, wb = j * p.stride - p.pw;
        const float *xb = p.x + (long long)n * p.H * p.W * p.C;
        int u = 0, r = 0;                 // k = u*kw*C + r, r = v*C + c
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            float xv = 0.f;
            if (k < p.K) {
                const int v = r / p.C, c = r - v * p.C, h = hb + u, w = wb + v;
                if ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W) xv = __ldg(xb + (h * p.W + w) * p.C + c);
            }
#pragma unroll
            for (int f = 0; f < FT; ++f) acc[f][k] = fmaf(d[f], xv, acc[f][k]);
            if (++r == kwC) { r = 0; ++u; }
        }
#pragma unroll
        for (int f = 0; f < FT; ++f) acc[f][KT] += d[f];
    }
    __shared__ float red[8][FT * (KT + 1)];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int f = 0; f < FT; ++f)
#pragma unroll
        for (int k = 0; k <= KT; ++k) {
            float v = acc[f][k];
#pragma unroll
            for (int s = 16; s; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if (lane == 0) red[wid][f * (KT + 1) + k] = v;
        }
    __syncthreads();
    for (int i = threadIdx.x; i < FT * (KT + 1); i += 256) {
        const int f = i / (KT + 1), k = i - f * (KT + 1);
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[q][i];
        if (f < p.F) {
            if (k < p.K) atomicAdd(p.dw + (long long)f * p.K + k, v);
            else if (k == KT && p.db) atomicAdd(p.db + f, v);
        }
    }
}

}  // namespace

namespace sn {

bool direct_fprop_eligible(int C, int kh, int kw, int F) {
    int K = C * kh * kw;
    if (K > 64 || (long long)K * F * 5 > 48 * 1024) return false;
    if (F % 16 == 0) { int tpp = F / 16; return tpp <= 32 && (32 % tpp) == 0; }
    return F == 8;
}

int conv_fprop_direct(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                      int kh, int kw, int stride, int ph, int pw, int act, cudaStream_t st) {
    DirectP p{};
    p.x = x; p.w = w; p.bias = bias; p.y = y; p.N = N; p.C = C; p.H = H; p.W = W; p.F = F; p.kh = kh; p.kw = kw;
    p.stride = stride; p.ph = ph; p.pw = pw; p.act = act;
    p.oh = (H + 2 * ph - kh) / stride + 1; p.ow = (W + 2 * pw - kw) / stride + 1;
    p.K = C * kh * kw; p.P = (long long)N * p.oh * p.ow;
    if (p.P == 0) return SN_OK;
    const int FG = (F % 16 == 0) ? 16 : 8;
    size_t smem = sizeof(float) * p.K * (F / FG) * (FG + 4);
    const int ppw = 32 / (F / FG);
    long long warps_needed = (p.P + ppw * 2 - 1) / (ppw * 2);
    long long blocks = (warps_needed + 7) / 8;
    long long cap = (long long)num_sms() * 8;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (FG == 16) direct_fprop_kernel<16, 2><<<grid, 256, smem, st>>>(p);
    else          direct_fprop_kernel<8, 2><<<grid, 256, smem, st>>>(p);
    return after_launch("direct_fprop_kernel");
}

bool direct_wgrad_eligible(int C, int kh, int kw, int F) { return C * kh * kw < WG_KP && (F % WG_FT) == 0; }
// (callers have already checked that every pixel / element count fits in 31 bits: conv_geom)

// dw/db are accumulated into (+=).
int conv_wgrad_direct(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                      int kw, int stride, int ph, int pw, cudaStream_t st) {
    DirectP p{};
    p.x = x; p.dy = dy; p.dw = dw; p.db = db; p.N = N; p.C = C; p.H = H; p.W = W; p.F = F; p.kh = kh; p.kw = kw;
    p.stride = stride; p.ph = ph; p.pw = pw;
    p.oh = (H + 2 * ph - kh) / stride + 1; p.ow = (W + 2 * pw - kw) / stride + 1;
    p.K = C * kh * kw; p.P = (long long)N * p.oh * p.ow;
    if (p.P == 0) return SN_OK;
    int fy = F / WG_FT;
    long long tiles = (p.P + WG_PT - 1) / WG_PT;
    long long want = (long long)num_sms() * 8 / fy;          // 8 CTAs of 128 threads per SM
    if (want < 1) want = 1;
    long long bx = tiles < want ? tiles : want;
    long long ppb = ((tiles + bx - 1) / bx) * WG_PT;
    bx = (p.P + ppb - 1) / ppb;
    dim3 grid((unsigned)bx, fy);
    direct_wgrad_kernel<<<grid, 128, 0, st>>>(p, (int)ppb);
    return after_launch("direct_wgrad_kernel");
}

bool tiny_wgrad_eligible(int C, int kh, int kw, int F) {
    const int K = C * kh * kw;
    return (F <= 8 && K <= 9) || (F <= 16 && K <= 4);
}
// dw/db are accumulated into (+=).
int conv_wgrad_tiny(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                    int kw, int stride, int ph, int pw, cudaStream_t st) {
    DirectP p{};
    p.x = x; p.dy = dy; p.dw = dw; p.db = db; p.N = N; p.C = C; p.H = H; p.W = W; p.F = F; p.kh = kh; p.kw = kw;
    p.stride = stride; p.ph = ph; p.pw = pw;
    p.oh = (H + 2 * ph - kh) / stride + 1; p.ow = (W + 2 * pw - kw) / stride + 1;
    p.K = C * kh * kw; p.P = (long long)N * p.oh * p.ow;
    if (p.P == 0) return SN_OK;
    long long need = (p.P + 255) / 256;
    const int grid = (int)(need < num_sms() ? need : num_sms());
    if (F <= 8 && p.K <= 4) tiny_wgrad_kernel<8, 4><<<grid, 256, 0, st>>>(p);
    else if (F <= 8) tiny_wgrad_kernel<8, 9><<<grid, 256, 0, st>>>(p);
    else tiny_wgrad_kernel<16, 4><<<grid, 256, 0, st>>>(p);
    return after_launch("tiny_wgrad_kernel");
}

}  // namespace sn
