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
// Everything here is exact fp32 FMA work: the per-step products are N x H x 4H = 16.8 MFLOP, latency-bound.
// (The bf16 tier runs the recurrence of H = 128 on the tensor pipe instead: recurrent_tc.cu.)
#include "common.cuh"
using namespace sn;

namespace {

// Row lookup with NumPy's index semantics (the reference's W[idx]): ids in [-vocab, 0) wrap to the end
// of the table; anything else outside [0, vocab) raises IndexError there.  A kernel cannot raise: such an
// id is counted in g_embedding_oob (read back by sn_embedding_oob_count, which fit() checks at the end
// of every epoch) and clamped so that the access stays in bounds.
__device__ unsigned long long g_embedding_oob = 0ull;
__device__ __forceinline__ int embedding_row(float id, int vocab) {
    int row = (int)id;
    if (row < 0) row += vocab;
    if (row < 0 || row >= vocab) {
        atomicAdd(&g_embedding_oob, 1ull);
        row = row < 0 ? 0 : vocab - 1;
    }
    return row;
}

__global__ void __launch_bounds__(256) embedding_fwd_kernel(const float *__restrict__ w, const float *__restrict__ idx,
                                                            float *__restrict__ out, long long n_idx, int vocab, int dim4) {
    // one thread per (index, float4 of the row)
    const long long total = n_idx * dim4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim4;
        const int q = (int)(i - r * dim4);
        const int row = embedding_row(__ldg(idx + r), vocab);
        reinterpret_cast<float4 *>(out)[i] = __ldg(reinterpret_cast<const float4 *>(w) + (long long)row * dim4 + q);
    }
}

__global__ void __launch_bounds__(256) embedding_fwd_scalar_kernel(const float *__restrict__ w, const float *__restrict__ idx,
                                                                   float *__restrict__ out, long long n_idx, int vocab, int dim) {
    const long long total = n_idx * dim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim;
        const int row = embedding_row(__ldg(idx + r), vocab);
        out[i] = __ldg(w + (long long)row * dim + (i - r * dim));
    }
}

__global__ void __launch_bounds__(256) embedding_bwd_kernel(const float *__restrict__ idx, const float *__restrict__ dy,
                                                            float *__restrict__ dw, long long n_idx, int vocab, int dim) {
    const long long total = n_idx * dim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dim;
        const int row = embedding_row(__ldg(idx + r), vocab);
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
#pragma unroll
        for (int i = threadIdx.x; i < LS_ROWS * LS_KC; i += LS_ROWS * LS_UNITS) {
            const int rr = i / LS_KC, kk = i % LS_KC, row = blockIdx.y * LS_ROWS + rr;
            hs[rr][kk] = (row < N && kk < kc && h_prev) ? __ldg(h_prev + (long long)row * H + k0 + kk) : 0.f;
        }
#pragma unroll          // all 16 loads of a thread in flight before the first store (a rolled loop pays the L2 latency 16 times)
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
#pragma unroll
            for (int i = threadIdx.x; i < LS_ROWS * LS_KC; i += LS_ROWS * LS_UNITS) {
                const int rr = i / LS_KC, kk = i % LS_KC, row = blockIdx.y * LS_ROWS + rr;
                ds[rr][kk] = (row < N && kk < kc) ? __ldg(dz_next + (long long)row * K + k0 + kk) : 0.f;
            }
#pragma unroll
            for (int i = threadIdx.x; i < LS_KC * LS_UNITS; i += LS_ROWS * LS_UNITS) {
                const int u = i % LS_UNITS, kk = i / LS_UNITS, unit = blockIdx.x * LS_UNITS + u;
                ws[kk][u] = (unit < H && kk < kc) ? __ldg(wh + (long long)(k0 + kk) * H + unit) : 0.f;
            }
            __syncthreads();
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;      // four chains: one would be 512 dependent FMAs
#pragma unroll 4
            for (int kk = 0; kk < LS_KC; kk += 4) {
                p0 = fmaf(ds[tr][kk], ws[kk][tj], p0);
                p1 = fmaf(ds[tr][kk + 1], ws[kk + 1][tj], p1);
                p2 = fmaf(ds[tr][kk + 2], ws[kk + 2][tj], p2);
                p3 = fmaf(ds[tr][kk + 3], ws[kk + 3][tj], p3);
            }
            dh += (p0 + p1) + (p2 + p3);
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

// ---------------------------------------------------------------------------------------------
// Whole-sequence kernels: ONE launch runs all T steps of the recurrence.
//
// A thread-block CLUSTER of CL = H/32 CTAs owns SQ_ROWS batch rows; CTA `rank` owns the hidden units
// [32*rank, 32*rank + 32) — all four gate columns of them.  Its slice of Wh never leaves the SM: every
// thread keeps its 4 x H/8 weights in REGISTERS for the whole sequence (thread = unit u x k-slice ks;
// forward: Wh[g*H + unit][ks*H/8 ..], backward: Wh[ks*4H/8 ..][unit]), so a step costs no weight
// traffic at all.  Per step: (1) partial products of the thread's k-slice for all rows of the cluster
// (h / dz rows are read from shared memory as broadcast 128-bit loads), (2) an 8-way reduction of the
// k-slices through shared memory, (3) thread (row, unit) finishes the cell: gates, c (kept in a
// register across steps) and h, written to HBM for the backward pass and — through distributed
// shared memory (st.shared::cluster) — into the next-step buffer of every CTA of the cluster,
// (4) one cluster barrier.  Nothing but that barrier separates two steps: no launch, no L2 round trip
// for the recurrent operand.  At batch 256, H = 128: 32 clusters x 4 CTAs = 128 SMs.
constexpr int SQ_ROWS = 8;

__device__ __forceinline__ uint32_t sq_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sq_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void sq_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void sq_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void sq_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void sq_store_remote(uint32_t local_addr, uint32_t rank, float v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

template <int H>
__global__ void __launch_bounds__(256, 1) lstm_seq_fwd_kernel(const float *__restrict__ zx, const float *__restrict__ wh,
                                                               const float *__restrict__ h0, const float *__restrict__ c0,
                                                               float *__restrict__ acts, float *__restrict__ c_all,
                                                               float *__restrict__ h_all, float *__restrict__ h_seq, int T, int N) {
    constexpr int CL = H / 32, KPT = H / 8;                    // CTAs per cluster, contraction elements per thread
    __shared__ __align__(16) float hbuf[2][SQ_ROWS][H];
    __shared__ float red[8][4][SQ_ROWS][32];
    const int rank = (int)sq_ctarank();
    const int row0 = (int)(blockIdx.x / CL) * SQ_ROWS;
    const int u = threadIdx.x & 31, ks = threadIdx.x >> 5;     // FMA phase: unit x k-slice
    const int fr = threadIdx.x >> 5;                           // finish phase: row (same lane = unit)
    const int unit = rank * 32 + u;
    float w[4][KPT];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int k = 0; k < KPT; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(wh + ((long long)g * H + unit) * H + ks * KPT + k));
            w[g][k] = v.x; w[g][k + 1] = v.y; w[g][k + 2] = v.z; w[g][k + 3] = v.w;
        }
    for (int i = threadIdx.x; i < SQ_ROWS * H; i += 256) {
        const int r = i / H, k = i % H, row = row0 + r;
        hbuf[0][r][k] = (h0 && row < N) ? __ldg(h0 + (long long)row * H + k) : 0.f;
    }
    const int frow = row0 + fr;
    const bool live = frow < N;
    float c = (c0 && live) ? __ldg(c0 + (long long)frow * H + unit) : 0.f;
    sq_cluster_sync();                                         // every CTA of the cluster is resident and initialised
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1, nxt = cur ^ 1;
        // this step's input projections for the finishing thread: issued now, consumed after the FMAs
        float zin[4] = {0.f, 0.f, 0.f, 0.f};
        if (live) {
            const float *z = zx + ((long long)t * N + frow) * 4 * H + unit;
#pragma unroll
            for (int g = 0; g < 4; ++g) zin[g] = __ldg(z + g * H);
        }
        float acc[4][SQ_ROWS];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int r = 0; r < SQ_ROWS; ++r) acc[g][r] = 0.f;
#pragma unroll
        for (int r = 0; r < SQ_ROWS; ++r) {
#pragma unroll
            for (int k = 0; k < KPT; k += 4) {
                const float4 hv = *reinterpret_cast<const float4 *>(&hbuf[cur][r][ks * KPT + k]);     // broadcast
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    acc[g][r] = fmaf(hv.w, w[g][k + 3], fmaf(hv.z, w[g][k + 2], fmaf(hv.y, w[g][k + 1], fmaf(hv.x, w[g][k], acc[g][r]))));
            }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int r = 0; r < SQ_ROWS; ++r) red[ks][g][r][u] = acc[g][r];
        __syncthreads();
        float z[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float s = zin[g];
#pragma unroll
            for (int q = 0; q < 8; ++q) s += red[q][g][fr][u];
            z[g] = s;
        }
        const float f = sigmoidf_ref(z[0]), ug = sigmoidf_ref(z[1]), ct = tanhf(z[2]), o = sigmoidf_ref(z[3]);
        c = fmaf(f, c, ug * ct);
        const float h = o * tanhf(c);
        const uint32_t slot = sq_smem_u32(&hbuf[nxt][fr][unit]);
#pragma unroll
        for (int pr = 0; pr < CL; ++pr) sq_store_remote(slot, (uint32_t)pr, h);
        sq_cluster_arrive();                                   // the HBM stores below overlap the barrier's round trip
        if (live) {
            float *a = acts + ((long long)t * N + frow) * 4 * H + unit;
            a[0] = f; a[H] = ug; a[2 * H] = ct; a[3 * H] = o;
            const long long at = ((long long)(t + 1) * N + frow) * H + unit;
            c_all[at] = c;
            h_all[at] = h;
            if (h_seq) h_seq[((long long)frow * T + t) * H + unit] = h;
        }
        sq_cluster_wait();                                     // h_t is in every CTA's next buffer; `red` is free again
    }
}

// backward: dh_t = upstream + dz_{t+1} . Wh^T, contraction over the 4H gate columns; the cluster exchanges dz
template <int H>
__global__ void __launch_bounds__(256, 1) lstm_seq_bwd_kernel(const float *__restrict__ dh_last, const float *__restrict__ dh_seq,
                                                               const float *__restrict__ wh, const float *__restrict__ acts,
                                                               const float *__restrict__ c_all, int c0_is_zero,
                                                               float *__restrict__ dz, int T, int N) {
    constexpr int CL = H / 32, K = 4 * H, KPT = K / 8;
    __shared__ __align__(16) float zbuf[2][SQ_ROWS][K];
    __shared__ float red[8][SQ_ROWS][32];
    const int rank = (int)sq_ctarank();
    const int row0 = (int)(blockIdx.x / CL) * SQ_ROWS;
    const int u = threadIdx.x & 31, ks = threadIdx.x >> 5, fr = threadIdx.x >> 5;
    const int unit = rank * 32 + u;
    float w[KPT];                                              // Wh^T column of this unit, k-slice ks
#pragma unroll
    for (int k = 0; k < KPT; ++k) w[k] = __ldg(wh + (long long)(ks * KPT + k) * H + unit);
    const int frow = row0 + fr;
    const bool live = frow < N;
    float dc = 0.f;
    sq_cluster_sync();
    for (int t = T - 1; t >= 0; --t) {
        const int cur = t & 1, nxt = cur ^ 1;
        // operands of the finishing thread, issued before the FMAs
        float f = 0.f, ug = 0.f, ct = 0.f, o = 0.f, cv = 0.f, cp = 0.f, dh = 0.f;
        if (live) {
            const float *a = acts + ((long long)t * N + frow) * 4 * H + unit;
            f = __ldg(a); ug = __ldg(a + H); ct = __ldg(a + 2 * H); o = __ldg(a + 3 * H);
            cv = __ldg(c_all + ((long long)(t + 1) * N + frow) * H + unit);
            cp = (t == 0 && c0_is_zero) ? 0.f : __ldg(c_all + ((long long)t * N + frow) * H + unit);
            if (dh_seq) dh = __ldg(dh_seq + ((long long)frow * T + t) * H + unit);
            else if (t == T - 1 && dh_last) dh = __ldg(dh_last + (long long)frow * H + unit);
        }
        if (t != T - 1) {
            float acc[SQ_ROWS];
#pragma unroll
            for (int r = 0; r < SQ_ROWS; ++r) acc[r] = 0.f;
#pragma unroll
            for (int r = 0; r < SQ_ROWS; ++r) {
#pragma unroll
                for (int k = 0; k < KPT; k += 4) {
                    const float4 dv = *reinterpret_cast<const float4 *>(&zbuf[cur][r][ks * KPT + k]);     // broadcast
                    acc[r] = fmaf(dv.w, w[k + 3], fmaf(dv.z, w[k + 2], fmaf(dv.y, w[k + 1], fmaf(dv.x, w[k], acc[r]))));
                }
            }
#pragma unroll
            for (int r = 0; r < SQ_ROWS; ++r) red[ks][r][u] = acc[r];
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) dh += red[q][fr][u];
        }
        const float tc = tanhf(cv);
        const float dcv = dc + dh * o * (1.f - tc * tc);                       // layers/Recurrent.py:337-338
        float d[4];
        d[0] = dcv * cp * f * (1.f - f);                                       // forget gate   (:351-352)
        d[1] = dcv * ct * ug * (1.f - ug);                                     // update gate   (:345-346)
        d[2] = dcv * ug * (1.f - ct * ct);                                     // candidate     (:340-341)
        d[3] = dh * tc * o * (1.f - o);                                        // output gate   (:329-330)
        dc = dcv * f;                                                          // :366
        if (t > 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t slot = sq_smem_u32(&zbuf[nxt][fr][g * H + unit]);
#pragma unroll
                for (int pr = 0; pr < CL; ++pr) sq_store_remote(slot, (uint32_t)pr, d[g]);
            }
        }
        sq_cluster_arrive();
        if (live) {
            float *out = dz + ((long long)t * N + frow) * 4 * H + unit;
#pragma unroll
            for (int g = 0; g < 4; ++g) out[g * H] = d[g];
        }
        sq_cluster_wait();
    }
}

template <int H>
int launch_seq_fwd(const float *zx, const float *wh, const float *h0, const float *c0, float *acts, float *c_all, float *h_all,
                   float *h_seq, int T, int N, cudaStream_t st) {
    constexpr int CL = H / 32;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(((N + SQ_ROWS - 1) / SQ_ROWS) * CL);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SN_CUDA(cudaLaunchKernelEx(&cfg, lstm_seq_fwd_kernel<H>, zx, wh, h0, c0, acts, c_all, h_all, h_seq, T, N));
    return after_launch("lstm_seq_fwd_kernel");
}
template <int H>
int launch_seq_bwd(const float *dh_last, const float *dh_seq, const float *wh, const float *acts, const float *c_all, int c0_is_zero,
                   float *dz, int T, int N, cudaStream_t st) {
    constexpr int CL = H / 32;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(((N + SQ_ROWS - 1) / SQ_ROWS) * CL);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SN_CUDA(cudaLaunchKernelEx(&cfg, lstm_seq_bwd_kernel<H>, dh_last, dh_seq, wh, acts, c_all, c0_is_zero, dz, T, N));
    return after_launch("lstm_seq_bwd_kernel");
}

}  // namespace

extern "C" {

int sn_lstm_seq_supported(int N, int H) { return (N > 0 && (H == 64 || H == 128)) ? 1 : 0; }

int sn_lstm_seq_fwd(const float *zx, const float *wh, const float *h0, const float *c0, float *acts, float *c_all, float *h_all,
                    float *h_seq, int T, int N, int H, void *stream) {
    SN_REQUIRE(zx && wh && acts && c_all && h_all, "null tensor");
    SN_REQUIRE(T > 0 && sn_lstm_seq_supported(N, H), "the whole-sequence LSTM kernels take H = 64 or 128");
    SN_REQUIRE((((uintptr_t)wh) & 15) == 0, "wh must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    return H == 128 ? launch_seq_fwd<128>(zx, wh, h0, c0, acts, c_all, h_all, h_seq, T, N, st)
                    : launch_seq_fwd<64>(zx, wh, h0, c0, acts, c_all, h_all, h_seq, T, N, st);
}

int sn_lstm_seq_bwd(const float *dh_last, const float *dh_seq, const float *wh, const float *acts, const float *c_all, int c0_is_zero,
                    float *dz, int T, int N, int H, void *stream) {
    SN_REQUIRE(wh && acts && c_all && dz, "null tensor");
    SN_REQUIRE(dh_last || dh_seq, "no upstream gradient");
    SN_REQUIRE(T > 0 && sn_lstm_seq_supported(N, H), "the whole-sequence LSTM kernels take H = 64 or 128");
    cudaStream_t st = as_stream(stream);
    return H == 128 ? launch_seq_bwd<128>(dh_last, dh_seq, wh, acts, c_all, c0_is_zero, dz, T, N, st)
                    : launch_seq_bwd<64>(dh_last, dh_seq, wh, acts, c_all, c0_is_zero, dz, T, N, st);
}

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

int sn_embedding_oob_count(long long *count, int reset) {
    SN_REQUIRE(count, "null out pointer");
    unsigned long long v = 0ull;
    SN_CUDA(cudaMemcpyFromSymbol(&v, g_embedding_oob, sizeof(v)));
    *count = (long long)v;
    if (reset && v) { v = 0ull; SN_CUDA(cudaMemcpyToSymbol(g_embedding_oob, &v, sizeof(v))); }
    return SN_OK;
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
