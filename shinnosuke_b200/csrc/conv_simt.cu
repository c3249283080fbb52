// fp32 CUDA-core implicit-GEMM convolution: the SN_PREC_FP32 tier (1e-5 against the float64
// reference) and the fallback for shapes the tcgen05 path does not take.  No column matrix is
// materialised (the reference's utils/ConvCol.py im2col/col2im define only the index order).
//
//   FPROP : D[m=(n,i,j)][f]      = sum_{k=(u,v,c)} x[n,i*s+u-ph,j*s+v-pw,c] * w[f][k]
//   DGRAD : D[m=(n,h,w)][c]      = sum_{k=(u,v,f)} dy[n,(h+ph-u)/s,(w+pw-v)/s,f] * w[f][u][v][c]
//   WGRAD : D[m=f][n=(u,v,c)]   += sum_{k=(n,i,j)} dy[k][f] * x[n,i*s+u-ph,j*s+v-pw,c]   (split-K, atomics)
#include "common.cuh"
using namespace sn;

namespace {

struct ConvP {
    const float *a;      // FPROP: x   DGRAD: dy   WGRAD: dy
    const float *b;      // FPROP: w   DGRAD: w    WGRAD: x
    const float *bias;   // FPROP only (may be null)
    float *d;
    int N, C, H, W, F, kh, kw, stride, ph, pw, oh, ow;
    int M, Nn, K;        // GEMM extents
    int act, accumulate;
    int k_per_split;     // WGRAD
};

enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };
constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <int MODE>
__device__ __forceinline__ float load_a(const ConvP &p, int m, int k) {
    if (m >= p.M || k >= p.K) return 0.f;
    if (MODE == FPROP) {
        int j = m % p.ow, t = m / p.ow, i = t % p.oh, n = t / p.oh;
        int c = k % p.C, t2 = k / p.C, v = t2 % p.kw, u = t2 / p.kw;
        int h = i * p.stride + u - p.ph, w = j * p.stride + v - p.pw;
        if ((unsigned)h >= (unsigned)p.H || (unsigned)w >= (unsigned)p.W) return 0.f;
        return __ldg(p.a + (((long long)n * p.H + h) * p.W + w) * p.C + c);
    } else if (MODE == DGRAD) {
        int w = m % p.W, t = m / p.W, h = t % p.H, n = t / p.H;
        int f = k % p.F, t2 = k / p.F, v = t2 % p.kw, u = t2 / p.kw;
        int ti = h + p.ph - u, tj = w + p.pw - v;
        if (ti < 0 || tj < 0) return 0.f;
        int i = ti / p.stride, j = tj / p.stride;
        if (i * p.stride != ti || j * p.stride != tj || i >= p.oh || j >= p.ow) return 0.f;
        return __ldg(p.a + (((long long)n * p.oh + i) * p.ow + j) * p.F + f);
    } else {
        return __ldg(p.a + (long long)k * p.F + m);
    }
}

template <int MODE>
__device__ __forceinline__ float load_b(const ConvP &p, int k, int n) {
    if (n >= p.Nn || k >= p.K) return 0.f;
    if (MODE == FPROP) {
        return __ldg(p.b + (long long)n * p.K + k);
    } else if (MODE == DGRAD) {
        int f = k % p.F, uv = k / p.F;      // w[f][u][v][c]
        return __ldg(p.b + ((long long)f * p.kh * p.kw + uv) * p.C + n);
    } else {
        int c = n % p.C, t2 = n / p.C, v = t2 % p.kw, u = t2 / p.kw;
        int j = k % p.ow, t = k / p.ow, i = t % p.oh, img = t / p.oh;
        int h = i * p.stride + u - p.ph, w = j * p.stride + v - p.pw;
        if ((unsigned)h >= (unsigned)p.H || (unsigned)w >= (unsigned)p.W) return 0.f;
        return __ldg(p.b + (((long long)img * p.H + h) * p.W + w) * p.C + c);
    }
}

// A_KFAST / B_KFAST: which tile coordinate walks consecutive threads, chosen so that
// consecutive threads touch consecutive addresses of the operand in HBM.
template <int MODE, bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256) igemm_simt_kernel(ConvP p) {
    __shared__ __align__(16) float As[BK][BM + PAD];
    __shared__ __align__(16) float Bs[BK][BN + PAD];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    int k_begin = 0, k_end = p.K;
    if (MODE == WGRAD) {
        k_begin = blockIdx.z * p.k_per_split;
        k_end = min(p.K, k_begin + p.k_per_split);
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        float ra[4], rb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int ml = A_KFAST ? (t >> 4) + 16 * r : (t & 63);
            int kl = A_KFAST ? (t & 15) : (t >> 6) + 4 * r;
            int kk = k0 + kl;
            ra[r] = (kk < k_end) ? load_a<MODE>(p, m0 + ml, kk) : 0.f;
            int nl = B_KFAST ? (t >> 4) + 16 * r : (t & 63);
            int kl2 = B_KFAST ? (t & 15) : (t >> 6) + 4 * r;
            int kk2 = k0 + kl2;
            rb[r] = (kk2 < k_end) ? load_b<MODE>(p, kk2, n0 + nl) : 0.f;
        }
        __syncthreads();     // previous tile fully consumed
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int ml = A_KFAST ? (t >> 4) + 16 * r : (t & 63);
            int kl = A_KFAST ? (t & 15) : (t >> 6) + 4 * r;
            As[kl][ml] = ra[r];
            int nl = B_KFAST ? (t >> 4) + 16 * r : (t & 63);
            int kl2 = B_KFAST ? (t & 15) : (t >> 6) + 4 * r;
            Bs[kl2][nl] = rb[r];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 av = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.Nn) continue;
            float v = acc[i][j];
            float *dst = p.d + (long long)m * p.Nn + n;
            if (MODE == FPROP) {
                if (p.bias) v += __ldg(p.bias + n);
                if (p.act == SN_ACT_RELU) v = relu_keep_sign(v);
                *dst = v;
            } else if (MODE == DGRAD) {
                *dst = p.accumulate ? *dst + v : v;
            } else {
                atomicAdd(dst, v);
            }
        }
    }
}

// db[f] += sum_p dy[p][f]  (Conv2D.backward :192, Dense.backward :146)
__global__ void colsum_kernel(const float *__restrict__ dy, float *__restrict__ db, long long rows, int F) {
    int f = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (f < F)
        for (long long r = blockIdx.y * (long long)blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.y * blockDim.y)
            s += __ldg(dy + r * F + f);
    __shared__ float sm[8][33];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && f < F) {
        float tsum = 0.f;
        for (int y = 0; y < 8; ++y) tsum += sm[y][threadIdx.x];
        atomicAdd(db + f, tsum);
    }
}

int conv_geom(ConvP &p, int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw) {
    if (N < 0 || C <= 0 || H <= 0 || W <= 0 || F <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || ph < 0 || pw < 0 ||
        H + 2 * ph < kh || W + 2 * pw < kw) {
        set_error("conv2d: bad geometry N=%d C=%d H=%d W=%d F=%d k=%dx%d stride=%d pad=%d,%d", N, C, H, W, F, kh, kw, stride, ph, pw);
        return SN_ERR_INVALID_ARG;
    }
    p.N = N; p.C = C; p.H = H; p.W = W; p.F = F; p.kh = kh; p.kw = kw; p.stride = stride; p.ph = ph; p.pw = pw;
    p.oh = (H + 2 * ph - kh) / stride + 1;
    p.ow = (W + 2 * pw - kw) / stride + 1;
    long long pix_out = (long long)N * p.oh * p.ow, pix_in = (long long)N * H * W, kk = (long long)kh * kw * (C > F ? C : F);
    if (pix_out >= (1ll << 31) || pix_in >= (1ll << 31) || kk >= (1ll << 31)) {
        set_error("conv2d: tensor extents exceed 2^31 rows");
        return SN_ERR_UNSUPPORTED;
    }
    return SN_OK;
}

}  // namespace

namespace sn {

int colsum_accumulate(const float *dy, float *db, long long rows, int F, cudaStream_t st);

int conv_fprop_simt(const float *x, const float *w, const float *bias, float *y, int N, int C, int H, int W, int F,
                    int kh, int kw, int stride, int ph, int pw, int act, cudaStream_t st) {
    ConvP p{};
    int e = conv_geom(p, N, C, H, W, F, kh, kw, stride, ph, pw);
    if (e) return e;
    if (N == 0) return SN_OK;
    p.a = x; p.b = w; p.bias = bias; p.d = y; p.act = act;
    p.M = N * p.oh * p.ow; p.Nn = F; p.K = kh * kw * C;
    dim3 grid((p.M + BM - 1) / BM, (p.Nn + BN - 1) / BN);
    igemm_simt_kernel<FPROP, true, true><<<grid, 256, 0, st>>>(p);
    return after_launch("igemm_simt_kernel<FPROP>");
}

int conv_dgrad_simt(const float *dy, const float *w, float *dx, int N, int C, int H, int W, int F, int kh, int kw,
                    int stride, int ph, int pw, int accumulate, cudaStream_t st) {
    ConvP p{};
    int e = conv_geom(p, N, C, H, W, F, kh, kw, stride, ph, pw);
    if (e) return e;
    if (N == 0) return SN_OK;
    p.a = dy; p.b = w; p.d = dx; p.accumulate = accumulate;
    p.M = N * H * W; p.Nn = C; p.K = kh * kw * F;
    dim3 grid((p.M + BM - 1) / BM, (p.Nn + BN - 1) / BN);
    igemm_simt_kernel<DGRAD, true, false><<<grid, 256, 0, st>>>(p);
    return after_launch("igemm_simt_kernel<DGRAD>");
}

int conv_wgrad_simt(const float *x, const float *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh,
                    int kw, int stride, int ph, int pw, int accumulate, cudaStream_t st) {
    ConvP p{};
    int e = conv_geom(p, N, C, H, W, F, kh, kw, stride, ph, pw);
    if (e) return e;
    long long nw = (long long)F * kh * kw * C;
    if (!accumulate) {
        SN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * nw, st));
        if (db) SN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * F, st));
    }
    if (N == 0) return SN_OK;
    p.a = dy; p.b = x; p.d = dw;
    p.M = F; p.Nn = kh * kw * C; p.K = N * p.oh * p.ow;
    int tiles = ((p.M + BM - 1) / BM) * ((p.Nn + BN - 1) / BN);
    int want = (2 * num_sms() + tiles - 1) / tiles;                 // ~2 CTAs per SM in total
    int max_split = (p.K + 4 * BK - 1) / (4 * BK);
    int split = want < 1 ? 1 : (want > max_split ? max_split : want);
    if (split > 65535) split = 65535;
    p.k_per_split = (((p.K + split - 1) / split) + BK - 1) / BK * BK;
    split = (p.K + p.k_per_split - 1) / p.k_per_split;
    dim3 grid((p.M + BM - 1) / BM, (p.Nn + BN - 1) / BN, split);
    igemm_simt_kernel<WGRAD, false, false><<<grid, 256, 0, st>>>(p);
    e = after_launch("igemm_simt_kernel<WGRAD>");
    if (e) return e;
    if (db) e = colsum_accumulate(dy, db, p.K, F, st);
    return e;
}

int colsum_accumulate(const float *dy, float *db, long long rows, int F, cudaStream_t st) {
    int gx = (F + 31) / 32;
    long long chunks = (rows + 7) / 8, cap = (long long)num_sms() * 4 / gx;
    if (cap < 1) cap = 1;
    dim3 g2(gx, (unsigned)(chunks < cap ? chunks : cap));
    colsum_kernel<<<g2, dim3(32, 8), 0, st>>>(dy, db, rows, F);
    return after_launch("colsum_kernel");
}

}  // namespace sn
