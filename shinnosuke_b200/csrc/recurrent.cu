// Embedding and LSTM (BASELINE config 5: Embedding(1000 -> 64) -> LSTM(128) -> Dense, batch 256, T = 64).
//
// Embedding (layers/Embeddings.py:56-76): a row gather forward, np.add.at backward (atomic row
// scatter: duplicates of an index in the minibatch accumulate, as in the reference).
//
// LSTM (layers/Recurrent.py:216-507).  The reference runs, per timestep, ot = [h_{t-1}, x_t] . W + b
// with W (H + D, 4H) = rows [recurrent; input], gate columns (f, u, c~, o) (:296-299), then
//     c_t = f * c_{t-1} + u * c~,   h_t = o * tanh(c_t)                      (:300-302)
// On the device the input half of that product is ONE GEMM over all timesteps (zx = X . Wx + b, the
// Dense kernels), and the recurrence is one launch per step that fuses the recurrent product
// h_{t-1} . Wh with the gate non-linearities and the cell update (lstm_step_fwd_kernel): a block owns
// LS_ROWS batch rows x LS_UNITS hidden units, i.e. all four gate columns of those units, so the cell
// update needs no second pass.  Backward mirrors it: the step kernel turns dz_{t+1} . Wh^T (+ the
// upstream gradient of step t) into dh_t, then into the four gate pre-activation gradients dz_t
// (columns in the FORWARD order f, u, c~, o; the reference concatenates them as (o, c~, u, f), :358,
// which scrambles dW — DESIGN.md section 5) and dc_{t-1}; the weight / bias / input gradients are
// three GEMMs over all timesteps afterwards (Dense kernels).
// Everything here is exact fp32 FMA work in every precision tier: the per-step products are
// N x H x 4H = 16.8 MFLOP, latency-bound.
#include "common.cuh"
using namespace sn;

namespace {

__global__ void __launch_bounds__(256) embedding_fwd_kernel(const float *__restrict__ w, const float *__restrict__ idx,
                                                            float *__restrict__ out, long long n_idx, int vocab, int dim4) {
    // one thread per (index, float4 of the row)
    const long long total = n_idx * dim4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim4;
        const int q = (int)(i - r * dim4);
        int row = (int)__ldg(idx + r);
        row = row < 0 ? 0 : (row >= vocab ? vocab - 1 : row);        // NumPy would raise; clamp instead of reading out of bounds
        reinterpret_cast<float4 *>(out)[i] = __ldg(reinterpret_cast<const float4 *>(w) + (long long)row * dim4 + q);
    }
}

__global__ void __launch_bounds__(256) embedding_fwd_scalar_kernel(const float *__restrict__ w, const float *__restrict__ idx,
                                                                   float *__restrict__ out, long long n_idx, int vocab, int dim) {
    const long long total = n_idx * dim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim;
        int row = (int)__ldg(idx + r);
        row = row < 0 ? 0 : (row >= vocab ? vocab - 1 : row);
        out[i] = __ldg(w + (long long)row * dim + (i - r * dim));
    }
}

__global__ void __launch_bounds__(256) embedding_bwd_kernel(const float *__restrict__ idx, const float *__restrict__ dy,
                                                            float *__restrict__ dw, long long n_idx, int vocab, int dim) {
    const long long total = n_idx * dim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim;
        int row = (int)__ldg(idx + r);
        row = row < 0 ? 0 : (row >= vocab ? vocab - 1 : row);
        atomicAdd(dw + (long long)row * dim + (i - r * dim), __ldg(dy + i));
    }
}

// dst[b][a][:] (=|+=) src[a][b][:]   (batch-major <-> time-major sequences)
__global__ void __launch_bounds__(256) swap_axes_kernel(const float *__restrict__ src, float *__restrict__ dst, int A, int B, int D,
                                                        int accumulate) {
    const long long total = (long long)A * B * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const long long ab = i / D;
        const int b = (int)(ab % B), a = (int)(ab / B);
        const long long o = ((long long)b * A + a) * D + d;
        dst[o] = accumulate ? dst[o] + src[i] : src[i];
    }
}

constexpr int LS_ROWS = 16;        // batch rows per block
constexpr int LS_UNITS = 16;       // hidden units per block (x 4 gates = 64 weight rows)
constexpr int LS_KC = 64;          // contraction chunk staged in shared memory

__device__ __forceinline__ float sigmoidf_ref(float z) { return 1.f / (1.f + expf(-z)); }    // utils/Activator.py:58

// z[r][g*H + j] = zx[r][g*H + j] + sum_k h_prev[r][k] * wh[g*H + j][k];  gates;  c, h.
// thread (r, j) of the 16 x 16 block owns the four gates of unit j for row r.
__global__ void __launch_bounds__(LS_ROWS * LS_UNITS) lstm_step_fwd_kernel(const float *__restrict__ zx, long long zx_stride,
                                                                           const float *__restrict__ wh,
                                                                           const float *__restrict__ h_prev,
                                                                           const float *__restrict__ c_prev, float *__restrict__ acts,
                                                                           float *__restrict__ c_out, float *__restrict__ h_out,
                                                                           float *__restrict__ h_seq, long long hseq_stride, int N,
                                                                           int H) {
    __shared__ float hs[LS_ROWS][LS_KC + 1];
    __shared__ float ws[4][LS_UNITS][LS_KC + 1];
    const int tj = threadIdx.x % LS_UNITS, tr = threadIdx.x / LS_UNITS;
    const int r = blockIdx.y * LS_ROWS + tr, j = blockIdx.x * LS_UNITS + tj;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < H; k0 += LS_KC) {
        const int kc = min(LS_KC, H - k0);
        for (int i = threadIdx.x; i < LS_ROWS * LS_KC; i += LS_ROWS * LS_UNITS) {
            const int rr = i / LS_KC, kk = i % LS_KC, row = blockIdx.y * LS_ROWS + rr;
            hs[rr][kk] = (row < N && kk < kc && h_prev) ? __ldg(h_prev + (long long)row * H + k0 + kk) : 0.f;
        }
        for (int i = threadIdx.x; i < 4 * LS_UNITS * LS_KC; i += LS_ROWS * LS_UNITS) {
            const int kk = i % LS_KC, u = (i / LS_KC) % LS_UNITS, g = i / (LS_KC * LS_UNITS), unit = blockIdx.x * LS_UNITS + u;
            ws[g][u][kk] = (unit < H && kk < kc) ? __ldg(wh + ((long long)g * H + unit) * H + k0 + kk) : 0.f;
        }
        __syncthreads();
        if (h_prev) {
#pragma unroll 8
            for (int kk = 0; kk < LS_KC; ++kk) {
                const float hv = hs[tr][kk];
                acc[0] = fmaf(hv, ws[0][tj][kk], acc[0]);
                acc[1] = fmaf(hv, ws[1][tj][kk], acc[1]);
                acc[2] = fmaf(hv, ws[2][tj][kk], acc[2]);
                acc[3] = fmaf(hv, ws[3][tj][kk], acc[3]);
            }
        }
        __syncthreads();
    }
    if (r >= N || j >= H) return;
    const float *z = zx + (long long)r * zx_stride;
    const float f = sigmoidf_ref(acc[0] + __ldg(z + j));
    const float u = sigmoidf_ref(acc[1] + __ldg(z + H + j));
    const float ct = tanhf(acc[2] + __ldg(z + 2 * H + j));
    const float o = sigmoidf_ref(acc[3] + __ldg(z + 3 * H + j));
    const float cp = c_prev ? __ldg(c_prev + (long long)r * H + j) : 0.f;
    const float c = fmaf(f, cp, u * ct);
    const float h = o * tanhf(c);
    float *a = acts + (long long)r * 4 * H;
    a[j] = f; a[H + j] = u; a[2 * H + j] = ct; a[3 * H + j] = o;
    c_out[(long long)r * H + j] = c;
    h_out[(long long)r * H + j] = h;
    if (h_seq) h_seq[(long long)r * hseq_stride + j] = h;
}

// dh[r][j] = dh_ext[r][j] + sum_k dz_next[r][k] * wh[k][j]   (k over the 4H gate columns of step t+1)
// then the cell backward of layers/Recurrent.py:326-366 with the gate gradients in forward column order.
__global__ void __launch_bounds__(LS_ROWS * LS_UNITS) lstm_step_bwd_kernel(const float *__restrict__ dh_ext, long long dh_stride,
                                                                           const float *__restrict__ dz_next,
                                                                           const float *__restrict__ wh, const float *__restrict__ acts,
                                                                           const float *__restrict__ c, const float *__restrict__ c_prev,
                                                                           float *__restrict__ dc, float *__restrict__ dz, int N, int H,
                                                                           int first) {
    __shared__ float ds[LS_ROWS][LS_KC + 1];
    __shared__ float ws[LS_KC][LS_UNITS + 1];
    const int tj = threadIdx.x % LS_UNITS, tr = threadIdx.x / LS_UNITS;
    const int r = blockIdx.y * LS_ROWS + tr, j = blockIdx.x * LS_UNITS + tj;
    float dh = 0.f;
    if (dz_next) {
        const int K = 4 * H;
        for (int k0 = 0; k0 < K; k0 += LS_KC) {
            const int kc = min(LS_KC, K - k0);
            for (int i = threadIdx.x; i < LS_ROWS * LS_KC; i += LS_ROWS * LS_UNITS) {
                const int rr = i / LS_KC, kk = i % LS_KC, row = blockIdx.y * LS_ROWS + rr;
                ds[rr][kk] = (row < N && kk < kc) ? __ldg(dz_next + (long long)row * K + k0 + kk) : 0.f;
            }
            for (int i = threadIdx.x; i < LS_KC * LS_UNITS; i += LS_ROWS * LS_UNITS) {
                const int u = i % LS_UNITS, kk = i / LS_UNITS, unit = blockIdx.x * LS_UNITS + u;
                ws[kk][u] = (unit < H && kk < kc) ? __ldg(wh + (long long)(k0 + kk) * H + unit) : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < LS_KC; ++kk) dh = fmaf(ds[tr][kk], ws[kk][tj], dh);
            __syncthreads();
        }
    }
    if (r >= N || j >= H) return;
    if (dh_ext) dh += __ldg(dh_ext + (long long)r * dh_stride + j);
    const float *a = acts + (long long)r * 4 * H;
    const float f = a[j], u = a[H + j], ct = a[2 * H + j], o = a[3 * H + j];
    const long long at = (long long)r * H + j;
    const float tc = tanhf(__ldg(c + at));
    const float cp = c_prev ? __ldg(c_prev + at) : 0.f;
    const float dcv = (first ? 0.f : dc[at]) + dh * o * (1.f - tc * tc);          // :337-338
    float *d = dz + (long long)r * 4 * H;
    d[j] = dcv * cp * f * (1.f - f);                    // forget gate   (:351-352)
    d[H + j] = dcv * ct * u * (1.f - u);                // update gate   (:345-346)
    d[2 * H + j] = dcv * u * (1.f - ct * ct);           // candidate     (:340-341)
    d[3 * H + j] = dh * tc * o * (1.f - o);             // output gate   (:329-330)
    dc[at] = dcv * f;                                   // :366
}

}  // namespace

extern "C" {

int sn_embedding_fwd(const float *w, const float *idx, float *out, long long n_idx, int vocab, int dim, void *stream) {
    SN_REQUIRE(w && idx && out, "null tensor");
    SN_REQUIRE(n_idx >= 0 && vocab > 0 && dim > 0, "bad sizes");
    if (n_idx == 0) return SN_OK;
    cudaStream_t st = as_stream(stream);
    if ((dim & 3) == 0 && ((((uintptr_t)w) | ((uintptr_t)out)) & 15) == 0) {
        embedding_fwd_kernel<<<grid_for(n_idx * (dim / 4), 256), 256, 0, st>>>(w, idx, out, n_idx, vocab, dim / 4);
        return after_launch("embedding_fwd_kernel");
    }
    embedding_fwd_scalar_kernel<<<grid_for(n_idx * dim, 256), 256, 0, st>>>(w, idx, out, n_idx, vocab, dim);
    return after_launch("embedding_fwd_scalar_kernel");
}

int sn_embedding_bwd(const float *idx, const float *dy, float *dw, long long n_idx, int vocab, int dim, void *stream) {
    SN_REQUIRE(idx && dy && dw, "null tensor");
    SN_REQUIRE(n_idx >= 0 && vocab > 0 && dim > 0, "bad sizes");
    if (n_idx == 0) return SN_OK;
    embedding_bwd_kernel<<<grid_for(n_idx * dim, 256), 256, 0, as_stream(stream)>>>(idx, dy, dw, n_idx, vocab, dim);
    return after_launch("embedding_bwd_kernel");
}

int sn_swap_leading_axes(const float *src, float *dst, int A, int B, int D, int accumulate, void *stream) {
    SN_REQUIRE(src && dst, "null tensor");
    SN_REQUIRE(A > 0 && B > 0 && D > 0, "bad sizes");
    swap_axes_kernel<<<grid_for((long long)A * B * D, 256), 256, 0, as_stream(stream)>>>(src, dst, A, B, D, accumulate);
    return after_launch("swap_axes_kernel");
}

int sn_lstm_step_fwd(const float *zx, long long zx_stride, const float *wh, const float *h_prev, const float *c_prev, float *acts,
                     float *c, float *h, float *h_seq, long long hseq_stride, int N, int H, void *stream) {
    SN_REQUIRE(zx && wh && acts && c && h, "null tensor");
    SN_REQUIRE(N > 0 && H > 0 && zx_stride >= 4ll * H, "bad sizes");
    dim3 grid((H + LS_UNITS - 1) / LS_UNITS, (N + LS_ROWS - 1) / LS_ROWS);
    lstm_step_fwd_kernel<<<grid, LS_ROWS * LS_UNITS, 0, as_stream(stream)>>>(zx, zx_stride, wh, h_prev, c_prev, acts, c, h, h_seq,
                                                                             hseq_stride, N, H);
    return after_launch("lstm_step_fwd_kernel");
}

int sn_lstm_step_bwd(const float *dh_ext, long long dh_stride, const float *dz_next, const float *wh, const float *acts, const float *c,
                     const float *c_prev, float *dc, float *dz, int N, int H, int first, void *stream) {
    SN_REQUIRE(wh && acts && c && dc && dz, "null tensor");
    SN_REQUIRE(dh_ext || dz_next, "a step needs a gradient from above or from the next step");
    SN_REQUIRE(N > 0 && H > 0, "bad sizes");
    dim3 grid((H + LS_UNITS - 1) / LS_UNITS, (N + LS_ROWS - 1) / LS_ROWS);
    lstm_step_bwd_kernel<<<grid, LS_ROWS * LS_UNITS, 0, as_stream(stream)>>>(dh_ext, dh_stride, dz_next, wh, acts, c, c_prev, dc, dz, N, H,
                                                                             first);
    return after_launch("lstm_step_bwd_kernel");
}

}  // extern "C"
