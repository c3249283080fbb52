// BatchNormalization over x[rows][C] with the channel as the fastest index
// (layers/Normalization.py:93-230).  HBM-bound: one reduction pass + one elementwise pass per
// direction; per-channel sums are accumulated in float64 (B200 has full-rate-enough FP64 for a
// memory-bound reduction) so the 1e-5 tier holds even when |mean| >> std.
#include "common.cuh"
#include "bn_common.cuh"
#include "comm_dev.cuh"
#include <stdlib.h>
using namespace sn;

// blockDim = (32, 8): threadIdx.x walks 4 channels each (float4), threadIdx.y walks rows.
// grid = (ceil(C/128), row_chunks).  Accumulates sum a and sum b per channel into scratch
// (doubles): STATS: a = x, b = x*x.   BWD: a = dy, b = dy * xhat.
template <bool BWD>
__global__ void bn_reduce_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                 const float *__restrict__ mean, const float *__restrict__ invstd,
                                 double *__restrict__ scratch, long long rows, int C) {
    const int c0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
    if (c0 < C) {
        float mu[4] = {0, 0, 0, 0}, is[4] = {1, 1, 1, 1};
        if (BWD) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (c0 + k < C) { mu[k] = mean[c0 + k]; is[k] = invstd[c0 + k]; }
        }
        const bool vec = (C % 4 == 0);
        for (long long r = blockIdx.y * (long long)blockDim.y + threadIdx.y; r < rows; r += (long long)gridDim.y * blockDim.y) {
            float xv[4], gv[4];
            if (vec) {
                float4 t = ldg4(x + r * C + c0);
                xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
                if (BWD) { float4 g = ldg4(dy + r * C + c0); gv[0] = g.x; gv[1] = g.y; gv[2] = g.z; gv[3] = g.w; }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    xv[k] = (c0 + k < C) ? x[r * C + c0 + k] : 0.f;
                    if (BWD) gv[k] = (c0 + k < C) ? dy[r * C + c0 + k] : 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (BWD) { a[k] += (double)gv[k]; b[k] += (double)gv[k] * (double)((xv[k] - mu[k]) * is[k]); }
                else     { a[k] += (double)xv[k]; b[k] += (double)xv[k] * (double)xv[k]; }
            }
        }
    }
    __shared__ double sa[8][32][4], sb[8][32][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { sa[threadIdx.y][threadIdx.x][k] = a[k]; sb[threadIdx.y][threadIdx.x][k] = b[k]; }
    __syncthreads();
    if (threadIdx.y == 0 && c0 < C) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (c0 + k >= C) break;
            double ta = 0, tb = 0;
            for (int yy = 0; yy < 8; ++yy) { ta += sa[yy][threadIdx.x][k]; tb += sb[yy][threadIdx.x][k]; }
            atomicAdd(scratch + c0 + k, ta);
            atomicAdd(scratch + C + c0 + k, tb);
        }
    }
}

// y = gamma*(x-mean)*invstd + beta; block 0 also publishes mean / invstd / moving statistics.
__global__ void bn_apply_train_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                                      const float *__restrict__ beta, float *__restrict__ y,
                                      float *__restrict__ save_mean, float *__restrict__ save_invstd,
                                      float *__restrict__ moving_mean, float *__restrict__ moving_var,
                                      const double *__restrict__ scratch, long long rows, int C, float eps, float momentum) {
    extern __shared__ float sm[];            // scale[C], shift[C]
    float *scale = sm, *shift = sm + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double m = scratch[c] / (double)rows;
        double var = scratch[C + c] / (double)rows - m * m;
        if (var < 0) var = 0;
        double is = 1.0 / sqrt(var + (double)eps);
        float g = gamma[c];
        scale[c] = (float)(g * is);
        shift[c] = (float)(beta[c] - g * is * m);
        if (blockIdx.x == 0) {
            save_mean[c] = (float)m;
            save_invstd[c] = (float)is;
            moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * (float)m;
            moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * (float)var;
        }
    }
    __syncthreads();
    const long long total = rows * C;
    if (C % 4 == 0) {
        const int C4 = C >> 2;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (total >> 2); i += (long long)gridDim.x * blockDim.x) {
            int c = (int)(i % C4) * 4;
            float4 v = ldg4(x + 4 * i);
            v.x = v.x * scale[c] + shift[c]; v.y = v.y * scale[c + 1] + shift[c + 1];
            v.z = v.z * scale[c + 2] + shift[c + 2]; v.w = v.w * scale[c + 3] + shift[c + 3];
            reinterpret_cast<float4 *>(y)[i] = v;
        }
    } else {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            int c = (int)(i % C);
            y[i] = x[i] * scale[c] + shift[c];
        }
    }
}

__global__ void bn_apply_infer_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                                      const float *__restrict__ beta, float *__restrict__ y,
                                      const float *__restrict__ moving_mean, const float *__restrict__ moving_var,
                                      long long rows, int C, float eps) {
    extern __shared__ float sm[];
    float *scale = sm, *shift = sm + C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = gamma[c] / sqrtf(moving_var[c] + eps);
        scale[c] = s;
        shift[c] = beta[c] - moving_mean[c] * s;
    }
    __syncthreads();
    const long long total = rows * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        y[i] = x[i] * scale[c] + shift[c];
    }
}

// dx = gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat)); block 0 adds dgamma / dbeta.
template <bool ACC, bool MASK>
__global__ void bn_bwd_apply_kernel(const float *__restrict__ dy, const float *__restrict__ x,
                                    const float *__restrict__ gamma, const float *__restrict__ mean,
                                    const float *__restrict__ invstd, float *__restrict__ dx,
                                    float *__restrict__ dgamma, float *__restrict__ dbeta,
                                    const double *__restrict__ scratch, long long rows, int C) {
    extern __shared__ float sm[];            // k1[C] = gamma*invstd, k2[C] = mean(dy), k3[C] = mean(dy*xhat), mu[C], is[C]
    float *k1 = sm, *k2 = sm + C, *k3 = sm + 2 * C, *mu = sm + 3 * C, *is = sm + 4 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double sdy = scratch[c], sdyx = scratch[C + c];
        k1[c] = gamma[c] * invstd[c];
        k2[c] = (float)(sdy / (double)rows);
        k3[c] = (float)(sdyx / (double)rows);
        mu[c] = mean[c];
        is[c] = invstd[c];
        if (blockIdx.x == 0) {
            if (dbeta) dbeta[c] += (float)sdy;
            if (dgamma) dgamma[c] += (float)sdyx;
        }
    }
    __syncthreads();
    const long long total = rows * C;
    if (C % 4 == 0) {
        const int C4 = C >> 2;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (total >> 2); i += (long long)gridDim.x * blockDim.x) {
            int c = (int)(i % C4) * 4;
            float4 g = ldg4(dy + 4 * i), xv = ldg4(x + 4 * i), o;
            float xs[4] = {xv.x, xv.y, xv.z, xv.w}, gs[4] = {g.x, g.y, g.z, g.w}, r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float xh = (xs[k] - mu[c + k]) * is[c + k];
                r[k] = k1[c + k] * (gs[k] - k2[c + k] - xh * k3[c + k]);
                if (MASK && relu_blocked(xs[k])) r[k] = 0.f;
            }
            if (ACC) { float4 p = reinterpret_cast<float4 *>(dx)[i]; r[0] += p.x; r[1] += p.y; r[2] += p.z; r[3] += p.w; }
            o.x = r[0]; o.y = r[1]; o.z = r[2]; o.w = r[3];
            reinterpret_cast<float4 *>(dx)[i] = o;
        }
    } else {
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            int c = (int)(i % C);
            float xh = (x[i] - mu[c]) * is[c];
            float r = k1[c] * (dy[i] - k2[c] - xh * k3[c]);
            if (MASK && relu_blocked(x[i])) r = 0.f;
            dx[i] = ACC ? dx[i] + r : r;
        }
    }
}


// ------------------------------------------------------------------------------------------
// Fast path for C = 4 * 2^k <= 1024 (every BASELINE config): a block of 256 threads walks the
// tensor as one flat stream of float4s, so thread t always sees the same channel quad
// (256 % (C/4) == 0) — no index arithmetic in the loop, per-channel constants live in registers,
// and a block reads/writes 4 KB contiguous per step.  Per-thread partial sums are fp32 over a
// bounded run (the forward shifts by a per-channel pivot x[0][c], so E[(x-k)^2] - E[x-k]^2 has no
// cancellation); across threads and blocks they are combined in float64.
constexpr int BN_UNROLL = 4;

template <bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_pow2_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                             const float *__restrict__ mean, const float *__restrict__ invstd,
                                                             double *__restrict__ scratch, long long total4, int C, BnFinal fin) {
    const int Q = C >> 2;
    const int cq = threadIdx.x % Q;
    float k0[4], k1[4];          // STATS: pivot, unused.   BWD: mean, invstd
    {
        // forward pivot: row 0 of x, or 0 when the sums must be combinable across replicas (fin.x == nullptr)
        float4 a = (BWD || fin.x) ? ldg4((BWD ? mean : x) + cq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        k0[0] = a.x; k0[1] = a.y; k0[2] = a.z; k0[3] = a.w;
        if (BWD) { float4 b = ldg4(invstd + cq * 4); k1[0] = b.x; k1[1] = b.y; k1[2] = b.z; k1[3] = b.w; }
    }
    double da[4] = {0, 0, 0, 0}, db[4] = {0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * 256;
    long long i = blockIdx.x * 256ll + threadIdx.x;
    while (i < total4) {
        float fa[4] = {0, 0, 0, 0}, fb[4] = {0, 0, 0, 0};
#pragma unroll 1
        for (int chunk = 0; chunk < 8 && i < total4; ++chunk) {
            float4 xv[BN_UNROLL], gv[BN_UNROLL];
#pragma unroll
            for (int u = 0; u < BN_UNROLL; ++u) {
                long long ii = i + u * stride;
                bool ok = ii < total4;
                xv[u] = ok ? ldg4(x + 4 * ii) : make_float4(k0[0], k0[1], k0[2], k0[3]);
                if (BWD) gv[u] = ok ? ldg4(dy + 4 * ii) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < BN_UNROLL; ++u) {
                float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
                float gs[4] = {0, 0, 0, 0};
                if (BWD) { gs[0] = gv[u].x; gs[1] = gv[u].y; gs[2] = gv[u].z; gs[3] = gv[u].w; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (BWD) { fa[k] += gs[k]; fb[k] = fmaf(gs[k], (xs[k] - k0[k]) * k1[k], fb[k]); }
                    else { float d = xs[k] - k0[k]; fa[k] += d; fb[k] = fmaf(d, d, fb[k]); }
                }
            }
            i += BN_UNROLL * stride;
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

// RELU: the Activation('relu') that follows the layer, applied on the way out (blocked elements are
// written as -0.0f, the convention every ReLU mask on the device reads back)
template <bool RELU>
__global__ void __launch_bounds__(256) bn_apply_train_pow2_kernel(const float *__restrict__ x, float *__restrict__ y,
                                                                  const double *__restrict__ scratch, long long total4, int C) {
    pdl_wait();
    const int Q = C >> 2;
    const int cq = threadIdx.x % Q;
    const float *prm = bn_params(scratch, C);
    const float4 sc4 = ldg4(prm + cq * 4), sh4 = ldg4(prm + C + cq * 4);
    const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total4; i += BN_UNROLL * stride) {
        float4 v[BN_UNROLL];
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) if (i + u * stride < total4) v[u] = ldg4(x + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < BN_UNROLL; ++u) {
            if (i + u * stride >= total4) break;
            float4 o;
            o.x = fmaf(v[u].x, sc[0], sh[0]); o.y = fmaf(v[u].y, sc[1], sh[1]);
            o.z = fmaf(v[u].z, sc[2], sh[2]); o.w = fmaf(v[u].w, sc[3], sh[3]);
            if (RELU) { o.x = relu_keep_sign(o.x); o.y = relu_keep_sign(o.y); o.z = relu_keep_sign(o.z); o.w = relu_keep_sign(o.w); }
            reinterpret_cast<float4 *>(y)[i + u * stride] = o;
        }
    }
}

template <bool ACC, bool MASK>
__global__ void __launch_bounds__(256) bn_bwd_apply_pow2_kernel(const float *__restrict__ dy, const float *__restrict__ x,
                                                                float *__restrict__ dx, const double *__restrict__ scratch,
                                                                long long total4, int C) {
    pdl_wait();
    const int Q = C >> 2;
    const int cq = threadIdx.x % Q;
    const float *prm = bn_params(scratch, C);
    const float4 a4 = ldg4(prm + cq * 4), b4 = ldg4(prm + C + cq * 4), c4 = ldg4(prm + 2 * C + cq * 4);
    const float pa[4] = {a4.x, a4.y, a4.z, a4.w}, pb[4] = {b4.x, b4.y, b4.z, b4.w}, pc[4] = {c4.x, c4.y, c4.z, c4.w};
    const long long stride = (long long)gridDim.x * 256;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total4; i += 2 * stride) {
        float4 g[2], xv[2], old[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (i + u * stride < total4) {
                g[u] = ldg4(dy + 4 * (i + u * stride));
                xv[u] = ldg4(x + 4 * (i + u * stride));
                if (ACC) old[u] = reinterpret_cast<const float4 *>(dx)[i + u * stride];
            }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (i + u * stride >= total4) break;
            float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, gs[4] = {g[u].x, g[u].y, g[u].z, g[u].w}, r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                r[k] = fmaf(pa[k], gs[k], fmaf(pb[k], xs[k], pc[k]));       // gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat))
                if (MASK && relu_blocked(xs[k])) r[k] = 0.f;
            }
            if (ACC) { r[0] += old[u].x; r[1] += old[u].y; r[2] += old[u].z; r[3] += old[u].w; }
            reinterpret_cast<float4 *>(dx)[i + u * stride] = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
}

__global__ void bn_finalize_kernel(double *scratch, BnFinal fin) {
    pdl_wait();
    bn_finalize_channels(scratch, fin, (int)(blockIdx.x * blockDim.x + threadIdx.x), (int)(gridDim.x * blockDim.x));
}

// Data-parallel finalize: the cross-replica sum of the statistics happens INSIDE this kernel, over
// peer-mapped memory: each thread folds the local slots of its channel and stores its {sum_a, sum_b}
// pair into this rank's slot in EVERY replica's window as four 8-byte "LL" words {32 data bits, sequence
// number} (csrc/comm_dev.cuh); it then polls the G slots of its own window until each word carries this
// exchange's sequence number, adds them in rank order (bit-identical on every replica) and finalises with
// the GLOBAL row count.  No flag, no fence, no barrier: the exchange costs one one-way NVLink latency.
// One launch instead of all-reduce + finalize, 32 bytes per channel and peer on the wire.
// Slots are double-buffered by the parity of the sequence number (a replica can be at most one exchange
// ahead of the slowest one, because finishing exchange k needs every replica's words of exchange k).
// PUSHED: the local sums were folded and sent by bn_dp_push_kernel earlier in the stream (other kernels ran in
// between, so the replicas' words are normally already here): only the poll + finalize half remains.
template <bool PUSHED>
__global__ void __launch_bounds__(32) bn_finalize_dp_kernel(comm::DevComm dc, double *scratch, BnFinal fin) {
    using namespace sn::comm;
    pdl_wait();
    const unsigned seq = *reinterpret_cast<volatile unsigned *>(&dc.state[0]) + 1u;     // device state: written only by exchange kernels of this window
    const int C = fin.C, c = blockIdx.x * 32 + threadIdx.x;
    const size_t slot0 = dc.ll_off + (size_t)(seq & 1u) * MAX_RANKS * LL_SLOT_BYTES;
    if (c < C) {
        const BnChannelIn in = bn_channel_inputs(fin, c);
        if (!PUSHED) {
            double a, b;
            bn_fold_slots(scratch, C, c, a, b);
#pragma unroll
            for (int p = 0; p < MAX_RANKS; ++p)
                if (p < dc.world) ll_store_f64x2(dc.t.win[p] + slot0 + (size_t)dc.rank * LL_SLOT_BYTES + (size_t)c * 32, a, b, seq);
        }
        const char *mine = dc.t.win[dc.rank] + slot0 + (size_t)c * 32;
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int p = 0; p < MAX_RANKS; ++p)
            if (p < dc.world) {
                double va, vb;
                ll_load_f64x2(mine + (size_t)p * LL_SLOT_BYTES, va, vb, seq, dc.state);
                if (p == 0) { sa = va; sb = vb; } else { sa += va; sb += vb; }
            }
        bn_finalize_one(bn_params(scratch, C), fin, in, c, sa, sb, bn_inv_rows(fin.rows));
    }
    finish_exchange(dc.state, seq);
}

// first half of a split exchange: fold the local slots and send them; the sequence number is NOT advanced (the
// pull half, bn_finalize_dp_kernel<true>, reads the same one and finishes the exchange)
__global__ void __launch_bounds__(32) bn_dp_push_kernel(comm::DevComm dc, double *scratch, int C) {
    using namespace sn::comm;
    pdl_wait();
    const unsigned seq = *reinterpret_cast<volatile unsigned *>(&dc.state[0]) + 1u;
    const int c = blockIdx.x * 32 + threadIdx.x;
    const size_t slot0 = dc.ll_off + (size_t)(seq & 1u) * MAX_RANKS * LL_SLOT_BYTES;
    if (c < C) {
        double a, b;
        bn_fold_slots(scratch, C, c, a, b);
#pragma unroll
        for (int p = 0; p < MAX_RANKS; ++p)
            if (p < dc.world) ll_store_f64x2(dc.t.win[p] + slot0 + (size_t)dc.rank * LL_SLOT_BYTES + (size_t)c * 32, a, b, seq);
    }
}

static bool bn_pow2(int C) {
    if (C & 3) return false;
    int q = C >> 2;
    return q >= 1 && q <= 256 && (q & (q - 1)) == 0;
}
static int bn_grid(long long total4, int ctas_per_sm = 8) {
    long long need = (total4 + 256 * BN_UNROLL - 1) / (256 * BN_UNROLL);
    long long cap = (long long)num_sms() * ctas_per_sm;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}
// reduce kernels: every block pays a shared-memory reduction, 2*C float64 atomics, a fence and a
// ticket, so they run fewer, longer-lived blocks than the purely streaming apply kernels
// (measured: 3 CTAs/SM is ~20 % faster up to ~64 MB maps, 8 CTAs/SM wins on the 134 MB stem map)
static int bn_reduce_ctas(long long total4) { return total4 * 16 >= (100ll << 20) ? 8 : 3; }

static dim3 reduce_grid(long long rows, int C) {
    int gx = (C + 127) / 128;
    long long chunks = (rows + 7) / 8;
    long long want = (long long)num_sms() * 8 / gx;       // ~8 CTAs per SM in total
    if (want < 1) want = 1;
    int gy = (int)(chunks < want ? chunks : want);
    if (gy < 1) gy = 1;
    return dim3(gx, gy);
}

extern "C" {

int sn_batchnorm_stats(const float *x, const float *gamma, const float *beta, float *save_mean, float *save_invstd,
                       float *moving_mean, float *moving_var, double *scratch, long long rows, int C, float eps,
                       float momentum, void *stream) {
    SN_REQUIRE(x && scratch && rows > 0, "bad arguments");
    SN_REQUIRE(bn_pow2(C) && (((uintptr_t)x) & 15) == 0, "needs C = 4*2^k <= 1024 and a 16-byte aligned tensor");
    cudaStream_t st = as_stream(stream);
    SN_REQUIRE(gamma && beta && save_mean && save_invstd && moving_mean && moving_var, "null tensor");
    SN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * BN_SCRATCH_DOUBLES_PER_C * C, st));
    const long long total4 = rows * C / 4;
    BnFinal fin{};
    fin.bwd = 0; fin.rows = rows; fin.C = C; fin.x = x; fin.gamma = gamma; fin.beta = beta; fin.save_mean = save_mean;
    fin.save_invstd = save_invstd; fin.moving_mean = moving_mean; fin.moving_var = moving_var; fin.eps = eps; fin.momentum = momentum;
    bn_reduce_pow2_kernel<false><<<bn_grid(total4, bn_reduce_ctas(total4)), 256, 0, st>>>(x, nullptr, nullptr, nullptr, scratch, total4, C, fin);
    return after_launch("bn_reduce_pow2_kernel");
}

int sn_batchnorm_finalize_stats(const float *gamma, const float *beta, float *save_mean, float *save_invstd,
                                float *moving_mean, float *moving_var, double *scratch, long long rows, int C, float eps,
                                float momentum, void *stream) {
    SN_REQUIRE(gamma && beta && save_mean && save_invstd && moving_mean && moving_var && scratch, "null tensor");
    SN_REQUIRE(rows > 0 && bn_pow2(C), "needs C = 4*2^k <= 1024");
    BnFinal fin{};
    fin.bwd = 0; fin.rows = rows; fin.C = C; fin.x = nullptr; fin.gamma = gamma; fin.beta = beta; fin.save_mean = save_mean;
    fin.save_invstd = save_invstd; fin.moving_mean = moving_mean; fin.moving_var = moving_var; fin.eps = eps; fin.momentum = momentum;
    SN_CUDA(launch_dependent(bn_finalize_kernel, dim3((C + 31) / 32), dim3(32), 0, as_stream(stream), scratch, fin));     // one warp per SM
    return after_launch("bn_finalize_kernel");
}

int sn_batchnorm_finalize_stats_dp(void *comm_handle, const float *gamma, const float *beta, float *save_mean, float *save_invstd,
                                   float *moving_mean, float *moving_var, double *scratch, long long rows_global, int C, float eps,
                                   float momentum, void *stream) {
    SN_REQUIRE(comm_handle && gamma && beta && save_mean && save_invstd && moving_mean && moving_var && scratch, "null argument");
    SN_REQUIRE(rows_global > 0 && bn_pow2(C), "needs C = 4*2^k <= 1024");
    const comm::Comm *cm = static_cast<const comm::Comm *>(comm_handle);
    SN_REQUIRE(cm->connected, "the communicator is not connected");
    BnFinal fin{};
    fin.bwd = 0; fin.rows = rows_global; fin.C = C; fin.x = nullptr; fin.gamma = gamma; fin.beta = beta; fin.save_mean = save_mean;
    fin.save_invstd = save_invstd; fin.moving_mean = moving_mean; fin.moving_var = moving_var; fin.eps = eps; fin.momentum = momentum;
    SN_CUDA(launch_dependent(bn_finalize_dp_kernel<false>, dim3((C + 31) / 32), dim3(32), 0, as_stream(stream), comm::dev_comm(cm), scratch, fin));
    return after_launch("bn_finalize_dp_kernel");
}

int sn_batchnorm_dp_push(void *comm_handle, double *scratch, int C, void *stream) {
    SN_REQUIRE(comm_handle && scratch && bn_pow2(C), "bad arguments");
    const comm::Comm *cm = static_cast<const comm::Comm *>(comm_handle);
    SN_REQUIRE(cm->connected, "the communicator is not connected");
    SN_CUDA(launch_dependent(bn_dp_push_kernel, dim3((C + 31) / 32), dim3(32), 0, as_stream(stream), comm::dev_comm(cm), scratch, C));
    return after_launch("bn_dp_push_kernel");
}

int sn_batchnorm_finalize_bwd_dp_pushed(void *comm_handle, const float *gamma, const float *save_mean, const float *save_invstd,
                                        double *scratch, long long rows_global, int C, void *stream) {
    SN_REQUIRE(comm_handle && gamma && save_mean && save_invstd && scratch && rows_global > 0 && bn_pow2(C), "bad arguments");
    const comm::Comm *cm = static_cast<const comm::Comm *>(comm_handle);
    SN_REQUIRE(cm->connected, "the communicator is not connected");
    BnFinal fin{};
    fin.bwd = 1; fin.rows = rows_global; fin.C = C; fin.gamma = gamma; fin.mean = save_mean; fin.invstd = save_invstd;
    SN_CUDA(launch_dependent(bn_finalize_dp_kernel<true>, dim3((C + 31) / 32), dim3(32), 0, as_stream(stream), comm::dev_comm(cm), scratch, fin));
    return after_launch("bn_finalize_dp_kernel (pushed)");
}

int sn_batchnorm_finalize_bwd_dp(void *comm_handle, const float *gamma, const float *save_mean, const float *save_invstd,
                                 double *scratch, long long rows_global, int C, void *stream) {
    SN_REQUIRE(comm_handle && gamma && save_mean && save_invstd && scratch && rows_global > 0 && bn_pow2(C), "bad arguments");
    const comm::Comm *cm = static_cast<const comm::Comm *>(comm_handle);
    SN_REQUIRE(cm->connected, "the communicator is not connected");
    BnFinal fin{};
    fin.bwd = 1; fin.rows = rows_global; fin.C = C; fin.gamma = gamma; fin.mean = save_mean; fin.invstd = save_invstd;
    SN_CUDA(launch_dependent(bn_finalize_dp_kernel<false>, dim3((C + 31) / 32), dim3(32), 0, as_stream(stream), comm::dev_comm(cm), scratch, fin));
    return after_launch("bn_finalize_dp_kernel");
}

int sn_batchnorm_fwd_apply(const float *x, float *y, const double *scratch, long long rows, int C, int relu, void *stream) {
    SN_REQUIRE(x && y && scratch && rows > 0, "bad arguments");
    SN_REQUIRE(bn_pow2(C) && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0, "needs C = 4*2^k <= 1024 and 16-byte aligned tensors");
    const long long total4 = rows * C / 4;
    if (relu) SN_CUDA(launch_dependent(bn_apply_train_pow2_kernel<true>, dim3(bn_grid(total4)), dim3(256), 0, as_stream(stream), x, y, scratch, total4, C));
    else SN_CUDA(launch_dependent(bn_apply_train_pow2_kernel<false>, dim3(bn_grid(total4)), dim3(256), 0, as_stream(stream), x, y, scratch, total4, C));
    return after_launch("bn_apply_train_pow2_kernel");
}

// ---- cross-replica (SyncBN) halves: local sums -> [caller all-reduces scratch[0, 16*C)] -> finalise with the global row count
int sn_batchnorm_local_stats(const float *x, double *scratch, long long rows, int C, void *stream) {
    SN_REQUIRE(x && scratch && rows > 0, "bad arguments");
    SN_REQUIRE(bn_pow2(C) && (((uintptr_t)x) & 15) == 0, "needs C = 4*2^k <= 1024 and a 16-byte aligned tensor");
    cudaStream_t st = as_stream(stream);
    SN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * BN_SCRATCH_DOUBLES_PER_C * C, st));
    const long long total4 = rows * C / 4;
    BnFinal fin{};
    fin.bwd = 0; fin.defer = 1; fin.rows = rows; fin.C = C; fin.x = nullptr;          // pivot 0: sums of x and x^2
    bn_reduce_pow2_kernel<false><<<bn_grid(total4, bn_reduce_ctas(total4)), 256, 0, st>>>(x, nullptr, nullptr, nullptr, scratch, total4, C, fin);
    return after_launch("bn_reduce_pow2_kernel");
}

int sn_batchnorm_bwd_local(const float *dy, const float *x, const float *save_mean, const float *save_invstd, float *dgamma,
                           float *dbeta, double *scratch, long long rows, int C, void *stream) {
    SN_REQUIRE(dy && x && save_mean && save_invstd && scratch && rows > 0, "bad arguments");
    SN_REQUIRE(bn_pow2(C) && ((((uintptr_t)x) | ((uintptr_t)dy)) & 15) == 0, "needs C = 4*2^k <= 1024 and 16-byte aligned tensors");
    cudaStream_t st = as_stream(stream);
    SN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * BN_SCRATCH_DOUBLES_PER_C * C, st));
    const long long total4 = rows * C / 4;
    BnFinal fin{};
    fin.bwd = 1; fin.defer = 1; fin.rows = rows; fin.C = C; fin.mean = save_mean; fin.invstd = save_invstd;
    fin.dgamma = dgamma; fin.dbeta = dbeta;
    bn_reduce_pow2_kernel<true><<<bn_grid(total4, bn_reduce_ctas(total4)), 256, 0, st>>>(x, dy, save_mean, save_invstd, scratch, total4, C, fin);
    return after_launch("bn_reduce_pow2_kernel");
}

int sn_batchnorm_finalize_bwd(const float *gamma, const float *save_mean, const float *save_invstd, double *scratch,
                              long long rows, int C, void *stream) {
    SN_REQUIRE(gamma && save_mean && save_invstd && scratch && rows > 0 && bn_pow2(C), "bad arguments");
    BnFinal fin{};
    fin.bwd = 1; fin.rows = rows; fin.C = C; fin.gamma = gamma; fin.mean = save_mean; fin.invstd = save_invstd;
    SN_CUDA(launch_dependent(bn_finalize_kernel, dim3((C + 31) / 32), dim3(32), 0, as_stream(stream), scratch, fin));     // one warp per SM
    return after_launch("bn_finalize_kernel");
}

int sn_batchnorm_bwd_apply(const float *dy, const float *x, float *dx, const double *scratch, long long rows, int C,
                           int accumulate_dx, int mask_relu, void *stream) {
    SN_REQUIRE(dy && x && dx && scratch && rows > 0, "bad arguments");
    SN_REQUIRE(bn_pow2(C) && ((((uintptr_t)x) | ((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0, "needs C = 4*2^k <= 1024 and 16-byte aligned tensors");
    cudaStream_t st = as_stream(stream);
    const long long total4 = rows * C / 4;
    const int g2 = bn_grid(total4);
#define SN_LAUNCHP(A, M) SN_CUDA(launch_dependent(bn_bwd_apply_pow2_kernel<A, M>, dim3(g2), dim3(256), 0, st, dy, x, dx, scratch, total4, C))
    if (accumulate_dx) { if (mask_relu) SN_LAUNCHP(true, true); else SN_LAUNCHP(true, false); }
    else { if (mask_relu) SN_LAUNCHP(false, true); else SN_LAUNCHP(false, false); }
#undef SN_LAUNCHP
    return after_launch("bn_bwd_apply_pow2_kernel");
}

int sn_batchnorm_fwd_train(const float *x, const float *gamma, const float *beta, float *y, float *save_mean,
                           float *save_invstd, float *moving_mean, float *moving_var, double *scratch,
                           long long rows, int C, float eps, float momentum, void *stream) {
    SN_REQUIRE(x && gamma && beta && y && save_mean && save_invstd && moving_mean && moving_var && scratch, "null tensor");
    SN_REQUIRE(rows > 0 && C > 0 && C <= 4096, "bad shape (C must be <= 4096)");
    cudaStream_t st = as_stream(stream);
    if (bn_pow2(C) && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0) {
        int e2 = sn_batchnorm_stats(x, gamma, beta, save_mean, save_invstd, moving_mean, moving_var, scratch, rows, C, eps,
                                    momentum, stream);
        if (e2) return e2;
        const long long total4 = rows * C / 4;
        bn_apply_train_pow2_kernel<false><<<bn_grid(total4), 256, 0, st>>>(x, y, scratch, total4, C);
        return after_launch("bn_apply_train_pow2_kernel");
    }
    SN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
    bn_reduce_kernel<false><<<reduce_grid(rows, C), dim3(32, 8), 0, st>>>(x, nullptr, nullptr, nullptr, scratch, rows, C);
    int e = after_launch("bn_reduce_kernel");
    if (e) return e;
    bn_apply_train_kernel<<<grid_for((rows * C + 3) / 4, 256), 256, 2 * C * sizeof(float), st>>>(
        x, gamma, beta, y, save_mean, save_invstd, moving_mean, moving_var, scratch, rows, C, eps, momentum);
    return after_launch("bn_apply_train_kernel");
}

int sn_batchnorm_fwd_infer(const float *x, const float *gamma, const float *beta, float *y, const float *moving_mean,
                           const float *moving_var, long long rows, int C, float eps, void *stream) {
    SN_REQUIRE(x && gamma && beta && y && moving_mean && moving_var, "null tensor");
    SN_REQUIRE(rows > 0 && C > 0 && C <= 4096, "bad shape (C must be <= 4096)");
    bn_apply_infer_kernel<<<grid_for(rows * C, 256), 256, 2 * C * sizeof(float), as_stream(stream)>>>(
        x, gamma, beta, y, moving_mean, moving_var, rows, C, eps);
    return after_launch("bn_apply_infer_kernel");
}

int sn_batchnorm_bwd(const float *dy, const float *x, const float *gamma, const float *save_mean,
                     const float *save_invstd, float *dx, float *dgamma, float *dbeta, double *scratch,
                     long long rows, int C, int accumulate_dx, int mask_relu, void *stream) {
    SN_REQUIRE(dy && x && gamma && save_mean && save_invstd && dx && scratch, "null tensor");
    SN_REQUIRE(rows > 0 && C > 0 && C <= 2048, "bad shape (C must be <= 2048)");
    cudaStream_t st = as_stream(stream);
    if (bn_pow2(C) && ((((uintptr_t)x) | ((uintptr_t)dy) | ((uintptr_t)dx)) & 15) == 0) {
        SN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * BN_SCRATCH_DOUBLES_PER_C * C, st));
        const long long total4 = rows * C / 4;
        BnFinal fin{};
        fin.bwd = 1; fin.rows = rows; fin.C = C; fin.gamma = gamma; fin.mean = save_mean; fin.invstd = save_invstd;
        fin.dgamma = dgamma; fin.dbeta = dbeta;
        bn_reduce_pow2_kernel<true><<<bn_grid(total4, bn_reduce_ctas(total4)), 256, 0, st>>>(x, dy, save_mean, save_invstd, scratch, total4, C, fin);
        int e2 = after_launch("bn_reduce_pow2_kernel");
        if (e2) return e2;
        const int g2 = bn_grid(total4);
#define SN_LAUNCHP(A, M) SN_CUDA(launch_dependent(bn_bwd_apply_pow2_kernel<A, M>, dim3(g2), dim3(256), 0, st, dy, x, dx, scratch, total4, C))
        if (accumulate_dx) { if (mask_relu) SN_LAUNCHP(true, true); else SN_LAUNCHP(true, false); }
        else { if (mask_relu) SN_LAUNCHP(false, true); else SN_LAUNCHP(false, false); }
#undef SN_LAUNCHP
        return after_launch("bn_bwd_apply_pow2_kernel");
    }
    SN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
    bn_reduce_kernel<true><<<reduce_grid(rows, C), dim3(32, 8), 0, st>>>(x, dy, save_mean, save_invstd, scratch, rows, C);
    int e = after_launch("bn_reduce_kernel");
    if (e) return e;
    int grid = grid_for((rows * C + 3) / 4, 256);
    size_t smem = 5 * C * sizeof(float);
#define SN_LAUNCH(A, M) bn_bwd_apply_kernel<A, M><<<grid, 256, smem, st>>>(dy, x, gamma, save_mean, save_invstd, dx, dgamma, dbeta, scratch, rows, C)
    if (accumulate_dx) { if (mask_relu) SN_LAUNCH(true, true); else SN_LAUNCH(true, false); }
    else { if (mask_relu) SN_LAUNCH(false, true); else SN_LAUNCH(false, false); }
#undef SN_LAUNCH
    return after_launch("bn_bwd_apply_kernel");
}

}  // extern "C"
