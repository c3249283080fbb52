// The classifier head: Dense(n_out <= 16) with linear / softmax activation (every BASELINE config
// ends in Dense(10, softmax)).  M = batch rows, K = features (1024 for config 3), n_out = 10: a
// GEMM far too thin for tiles — the tensor-core / SIMT tile kernels ran it as five launches
// (fprop, softmax, wgrad, bias column sums, dgrad) of 3-13 us each.  Here it is two launches
// that read x once each:
//
//   forward   (layers/FC.py:124-134 + utils/Activator.py:108-110)
//             block = 4 rows x all K: 256 threads stride over feature quads (coalesced 128-bit loads,
//             every W element fetched serves 4 rows), shuffle + shared-memory reduction, bias,
//             softmax by one warp per row.
//   backward  (layers/FC.py:140-153)
//             block = 32 features x a slice of the rows; dZ lives in shared memory; a thread owns one
//             feature column: dX[r][k] = sum_o dZ[r][o]*W[o][k] is written on the fly while
//             dW[o][k] += dZ[r][o]*x[r][k] accumulates in registers; db from the first feature block.
//
// Exact fp32 FMAs in every precision tier (the contraction is bandwidth-trivial).
#include "common.cuh"
using namespace sn;

namespace {

constexpr int MAXO = 16;

constexpr int HF_ROWS = 4;         // rows per block: each W element fetched by a thread serves 4 rows

// block = 4 rows x all K: thread t owns the feature quads {t, t+256, ...}; 4*O partial sums per thread,
// reduced by shuffles inside the warps and through shared memory across them
template <int O>
__global__ void __launch_bounds__(256) dense_head_fwd_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                             const float *__restrict__ bias, float *__restrict__ out, int M, int K,
                                                             int softmax) {
    __shared__ float sred[8][HF_ROWS][O];
    pdl_wait();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int r0 = blockIdx.x * HF_ROWS;
    float acc[HF_ROWS][O];
#pragma unroll
    for (int r = 0; r < HF_ROWS; ++r)
#pragma unroll
        for (int o = 0; o < O; ++o) acc[r][o] = 0.f;
    const float *xr[HF_ROWS];
#pragma unroll
    for (int r = 0; r < HF_ROWS; ++r) xr[r] = x + (long long)min(r0 + r, M - 1) * K;      // clamped: extra rows are discarded below
    if ((K & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)w)) & 15) == 0) {
        for (int k = threadIdx.x * 4; k < K; k += 1024) {
            float4 xv[HF_ROWS];
#pragma unroll
            for (int r = 0; r < HF_ROWS; ++r) xv[r] = __ldg(reinterpret_cast<const float4 *>(xr[r] + k));
#pragma unroll
            for (int o = 0; o < O; ++o) {
                const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + (long long)o * K + k));
#pragma unroll
                for (int r = 0; r < HF_ROWS; ++r)
                    acc[r][o] = fmaf(xv[r].x, wv.x, fmaf(xv[r].y, wv.y, fmaf(xv[r].z, wv.z, fmaf(xv[r].w, wv.w, acc[r][o]))));
            }
        }
    } else {
        for (int k = threadIdx.x; k < K; k += 256) {
            float xv[HF_ROWS];
#pragma unroll
            for (int r = 0; r < HF_ROWS; ++r) xv[r] = __ldg(xr[r] + k);
#pragma unroll
            for (int o = 0; o < O; ++o) {
                const float wv = __ldg(w + (long long)o * K + k);
#pragma unroll
                for (int r = 0; r < HF_ROWS; ++r) acc[r][o] = fmaf(xv[r], wv, acc[r][o]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < HF_ROWS; ++r)
#pragma unroll
        for (int o = 0; o < O; ++o) {
#pragma unroll
            for (int s = 16; s; s >>= 1) acc[r][o] += __shfl_xor_sync(0xffffffffu, acc[r][o], s);
            if (lane == 0) sred[wid][r][o] = acc[r][o];
        }
    __syncthreads();
    // warp r finishes row r: lane o sums the 8 warps' partials of logit o
    if (wid < HF_ROWS && r0 + wid < M) {
        float z = -INFINITY;
        if (lane < O) {
            z = bias ? __ldg(bias + lane) : 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) z += sred[q][wid][lane];
        }
        if (softmax) {
            float m = z;
#pragma unroll
            for (int s = 16; s; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
            float e = lane < O ? expf(z - m) : 0.f, t = e;
#pragma unroll
            for (int s = 16; s; s >>= 1) t += __shfl_xor_sync(0xffffffffu, t, s);
            z = e / t;
        }
        if (lane < O) out[(long long)(r0 + wid) * O + lane] = z;
    }
}

constexpr int HB_FEATS = 32;       // features per block (one warp-wide coalesced line of x / dX)
constexpr int HB_ROWGROUPS = 8;    // 256 threads = 8 row groups x 32 features

template <int O>
__global__ void __launch_bounds__(256) dense_head_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dz,
                                                             const float *__restrict__ w, float *__restrict__ dw,
                                                             float *__restrict__ db, float *__restrict__ dx, int M, int K,
                                                             int rows_per_block, int accumulate_dx) {
    extern __shared__ float sdz[];                       // [rows of this block][O]
    __shared__ float sred[HB_ROWGROUPS][O][HB_FEATS + 1];
    pdl_wait();
    const int f = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int k = blockIdx.x * HB_FEATS + f;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block), nr = r1 - r0;
    for (int i = threadIdx.x; i < nr * O; i += 256) sdz[i] = __ldg(dz + (long long)r0 * O + i);
    float wk[O], acc[O];
#pragma unroll
    for (int o = 0; o < O; ++o) { wk[o] = (k < K) ? __ldg(w + (long long)o * K + k) : 0.f; acc[o] = 0.f; }
    __syncthreads();
    if (k < K) {
        for (int r = g; r < nr; r += HB_ROWGROUPS) {
            const long long at = (long long)(r0 + r) * K + k;
            const float xv = x ? __ldg(x + at) : 0.f;
            float d = 0.f;
#pragma unroll
            for (int o = 0; o < O; ++o) {
                const float z = sdz[r * O + o];
                d = fmaf(z, wk[o], d);
                acc[o] = fmaf(z, xv, acc[o]);
            }
            if (dx) dx[at] = accumulate_dx ? dx[at] + d : d;
        }
    }
    if (dw) {
#pragma unroll
        for (int o = 0; o < O; ++o) sred[g][o][f] = acc[o];
        __syncthreads();
        // threads 0 .. O*32-1: (o, f) sums the 8 row groups
        for (int i = threadIdx.x; i < O * HB_FEATS; i += 256) {
            const int o = i >> 5, ff = i & 31, kk = blockIdx.x * HB_FEATS + ff;
            float s = 0.f;
#pragma unroll
            for (int gg = 0; gg < HB_ROWGROUPS; ++gg) s += sred[gg][o][ff];
            if (kk < K) atomicAdd(dw + (long long)o * K + kk, s);
        }
    }
    if (db && blockIdx.x == 0) {
        // column sums of this block's dZ rows
        for (int o = threadIdx.x; o < O; o += 256) {
            float s = 0.f;
            for (int r = 0; r < nr; ++r) s += sdz[r * O + o];
            atomicAdd(db + o, s);
        }
    }
}

template <int O>
int launch_fwd(const float *x, const float *w, const float *bias, float *out, int M, int K, int softmax, cudaStream_t st) {
    SN_CUDA(launch_dependent(dense_head_fwd_kernel<O>, dim3((M + HF_ROWS - 1) / HF_ROWS), dim3(256), 0, st, x, w, bias, out, M, K, softmax));
    return after_launch("dense_head_fwd_kernel");
}
template <int O>
int launch_bwd(const float *x, const float *dz, const float *w, float *dw, float *db, float *dx, int M, int K, int acc, cudaStream_t st) {
    const int fblocks = (K + HB_FEATS - 1) / HB_FEATS;
    // enough blocks for the machine, at most 256 rows of dZ per block (smem), at least 32
    int rb = (fblocks >= 2 * num_sms()) ? 256 : 64;
    if (rb > M) rb = M;
    const int yblocks = (M + rb - 1) / rb;
    SN_CUDA(launch_dependent(dense_head_bwd_kernel<O>, dim3(fblocks, yblocks), dim3(256), rb * O * sizeof(float), st, x, dz, w, dw, db, dx, M, K, rb, acc));
    return after_launch("dense_head_bwd_kernel");
}

}  // namespace

extern "C" {

int sn_dense_head_supported(int M, int K, int Nout) { return (M > 0 && K > 0 && Nout > 0 && Nout <= MAXO) ? 1 : 0; }

int sn_dense_head_fprop(const float *x, const float *w, const float *bias, float *out, int M, int K, int Nout, int softmax,
                        void *stream) {
    SN_REQUIRE(x && w && out, "null tensor");
    SN_REQUIRE(sn_dense_head_supported(M, K, Nout), "the classifier-head kernels take 1 <= n_out <= 16");
    cudaStream_t st = as_stream(stream);
    switch (Nout) {
#define SN_CASE(O) case O: return launch_fwd<O>(x, w, bias, out, M, K, softmax, st);
        SN_CASE(1) SN_CASE(2) SN_CASE(3) SN_CASE(4) SN_CASE(5) SN_CASE(6) SN_CASE(7) SN_CASE(8)
        SN_CASE(9) SN_CASE(10) SN_CASE(11) SN_CASE(12) SN_CASE(13) SN_CASE(14) SN_CASE(15) SN_CASE(16)
#undef SN_CASE
    }
    return SN_ERR_UNSUPPORTED;
}

int sn_dense_head_bwd(const float *x, const float *dz, const float *w, float *dw, float *db, float *dx, int M, int K, int Nout,
                      int accumulate_dx, void *stream) {
    SN_REQUIRE(dz && w && (dx || dw || db), "null tensor");
    SN_REQUIRE(x || !dw, "dW needs x");
    SN_REQUIRE(sn_dense_head_supported(M, K, Nout), "the classifier-head kernels take 1 <= n_out <= 16");
    cudaStream_t st = as_stream(stream);
    switch (Nout) {
#define SN_CASE(O) case O: return launch_bwd<O>(x, dz, w, dw, db, dx, M, K, accumulate_dx, st);
        SN_CASE(1) SN_CASE(2) SN_CASE(3) SN_CASE(4) SN_CASE(5) SN_CASE(6) SN_CASE(7) SN_CASE(8)
        SN_CASE(9) SN_CASE(10) SN_CASE(11) SN_CASE(12) SN_CASE(13) SN_CASE(14) SN_CASE(15) SN_CASE(16)
#undef SN_CASE
    }
    return SN_ERR_UNSUPPORTED;
}

}  // extern "C"
