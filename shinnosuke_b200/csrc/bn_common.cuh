// Shared by bn.cu and bn_pool.cu: layout of the float64 scratch and the "last block finalises"
// step that turns the per-channel sums into the few fp32 constants the apply pass needs.
//
// scratch (>= 20*C doubles, zeroed by the host wrapper before the reduce kernel; the finalise step
// leaves the sums and the ticket zero again):
//   [0, 16C)        BN_SLOTS copies of {sum_a[C], sum_b[C]}  (slot = blockIdx.x % BN_SLOTS: short atomic chains)
//   [16C, 18C)      4*C floats: p0[C], p1[C], p2[C] (apply-pass constants), 1 spare row
//   [18C]           ticket counter of the reduce kernels (unsigned)
//   [20C, 28C)      sn_bn_pool_bwd with dx_colsum only: BN_SLOTS x C doubles of block column sums (zero at rest: colsum_fold_kernel re-zeroes them)
//
// Doing the float64 division / sqrt ONCE per channel here (instead of in the prologue of every
// apply block) matters: B200 issues 64 FP64 ops/clk/SM, and a 256-thread block recomputing
// 4 channels each spent ~2 us in its prologue.
#pragma once
#include "common.cuh"

namespace sn {

constexpr int BN_SLOTS = 8;
constexpr int BN_SCRATCH_DOUBLES_PER_C = 20;

struct BnFinal {
    int bwd;                       // 0: forward statistics, 1: backward sums
    int defer;                     // 1: cross-replica (SyncBN) phase 1 — keep the LOCAL sums in the scratch for an
                                   //    all-reduce; only dgamma / dbeta (local by definition) are taken now
    long long rows;
    int C;
    // forward
    const float *x;                // pivot row x[0][c]
    const float *gamma, *beta;
    float *save_mean, *save_invstd, *moving_mean, *moving_var;
    float eps, momentum;
    // backward
    const float *mean, *invstd;
    float *dgamma, *dbeta;
};

__device__ __forceinline__ float *bn_params(double *scratch, int C) { return reinterpret_cast<float *>(scratch + 16 * (size_t)C); }
__device__ __forceinline__ const float *bn_params(const double *scratch, int C) { return reinterpret_cast<const float *>(scratch + 16 * (size_t)C); }

// Turns the per-channel sums in `scratch` into the apply-pass constants, per channel:
//   forward : p0 = gamma*invstd, p1 = beta - gamma*invstd*mean            (y  = p0*x + p1)
//   backward: p0 = gamma*invstd, p1 = -p0*k3*invstd, p2 = -p0*k2 - p1*mean (dx = p0*dy + p1*x + p2)
//             with k2 = sum(dy)/rows, k3 = sum(dy*xhat)/rows; dgamma += sum(dy*xhat), dbeta += sum(dy)
// Forward sums are of (x - pivot), pivot = f.x[0][c], or of x itself when f.x == nullptr (sums that
// a convolution epilogue produced).  One block's threads stride over the channels.
// (first, stride): which channels this thread takes — one block's threads by default; the stand-alone
// finalize kernel spreads single warps over many SMs instead (FP64 issue is the cost: ~40 double
// operations per channel at this part's 1/64 FP64 rate kept one SM busy for ~5 us at C = 256).
__device__ __forceinline__ double bn_inv_rows(long long rows) {
    // No float64 divide / sqrt (each is a long dependent chain on this part's few FP64 units, and this
    // code is the serial tail of a kernel): 1/rows comes from a float seed refined by two Newton steps,
    // 1/sqrt from rsqrtf refined by two — both exact to float64 rounding.
    const double r = (double)rows;
    double y = (double)(1.0f / (float)rows);
    y = y * (2.0 - r * y);
    return y * (2.0 - r * y);
}

// everything a channel's finalize step reads besides the sums, loaded up front (see the note on batched loads below)
struct BnChannelIn {
    float gamma, beta, pivot, mm, mv, is, mu, db, dg;
};
__device__ __forceinline__ BnChannelIn bn_channel_inputs(const BnFinal &f, int c) {
    BnChannelIn in;
    in.gamma = f.gamma ? f.gamma[c] : 0.f;
    in.beta = (!f.bwd && f.beta) ? f.beta[c] : 0.f;
    in.pivot = (!f.bwd && f.x) ? f.x[c] : 0.f;
    in.mm = (!f.bwd && f.moving_mean) ? f.moving_mean[c] : 0.f;
    in.mv = (!f.bwd && f.moving_var) ? f.moving_var[c] : 0.f;
    in.is = (f.bwd && f.invstd) ? f.invstd[c] : 0.f;
    in.mu = (f.bwd && f.mean) ? f.mean[c] : 0.f;
    in.db = (f.bwd && f.dbeta) ? f.dbeta[c] : 0.f;
    in.dg = (f.bwd && f.dgamma) ? f.dgamma[c] : 0.f;
    return in;
}

// sums (a, b) of channel c over the whole (global) batch -> apply-pass constants and saved statistics
__device__ __forceinline__ void bn_finalize_one(float *prm, const BnFinal &f, const BnChannelIn &in, int c, double a, double b,
                                                double inv_rows) {
    const int C = f.C;
    if (!f.bwd) {
        const double pivot = (double)in.pivot;
        const double m1 = a * inv_rows;                             // E[x - pivot]
        double var = b * inv_rows - m1 * m1;
        if (var < 0) var = 0;
        const double m = pivot + m1;
        const double ve = var + (double)f.eps;
        double is = (double)rsqrtf((float)ve);
        is = is * (1.5 - 0.5 * ve * is * is);
        is = is * (1.5 - 0.5 * ve * is * is);
        const double g = (double)in.gamma;
        prm[c] = (float)(g * is);
        prm[C + c] = (float)((double)in.beta - g * is * m);
        f.save_mean[c] = (float)m;
        f.save_invstd[c] = (float)is;
        f.moving_mean[c] = f.momentum * in.mm + (1.f - f.momentum) * (float)m;
        f.moving_var[c] = f.momentum * in.mv + (1.f - f.momentum) * (float)var;
    } else {
        const double is = (double)in.is, mu = (double)in.mu;
        const double k2 = a * inv_rows, k3 = b * inv_rows;
        const double p0 = (double)in.gamma * is;
        const double p1 = -p0 * k3 * is;
        prm[c] = (float)p0;
        prm[C + c] = (float)p1;
        prm[2 * C + c] = (float)(-p0 * k2 - p1 * mu);
        if (f.dbeta) f.dbeta[c] = in.db + (float)a;
        if (f.dgamma) f.dgamma[c] = in.dg + (float)b;
    }
}

// the BN_SLOTS partial sums of channel c -> (a, b); leaves the slots zero again, so that a producer
// that only ADDS into the scratch (the convolution epilogues, sn_conv2d_fprop_stats) needs no memset
__device__ __forceinline__ void bn_fold_slots(double *scratch, int C, int c, double &a, double &b) {
    // ALL loads first, in one batch: this is the serial tail of a kernel (or a one-block kernel), and a
    // load placed after a store to `scratch` cannot be hoisted above it by the compiler — 16 dependent
    // L2 round trips cost 5-6 us per launch when the zeroing stores were interleaved with the loads
    double va[BN_SLOTS], vb[BN_SLOTS];
#pragma unroll
    for (int s = 0; s < BN_SLOTS; ++s) {
        va[s] = __ldcg(scratch + (size_t)s * 2 * C + c);
        vb[s] = __ldcg(scratch + (size_t)s * 2 * C + C + c);
    }
    a = 0; b = 0;
#pragma unroll
    for (int s = 0; s < BN_SLOTS; ++s) { a += va[s]; b += vb[s]; }
#pragma unroll
    for (int s = 0; s < BN_SLOTS; ++s) {
        scratch[(size_t)s * 2 * C + c] = 0.0;
        scratch[(size_t)s * 2 * C + C + c] = 0.0;
    }
}

__device__ __forceinline__ void bn_finalize_channels(double *scratch, const BnFinal &f, int first = -1, int stride = 0) {
    if (first < 0) { first = threadIdx.x; stride = blockDim.x; }
    const int C = f.C;
    float *prm = bn_params(scratch, C);
    const double inv_rows = bn_inv_rows(f.rows);
    for (int c = first; c < C; c += stride) {
        const BnChannelIn in = bn_channel_inputs(f, c);
        double a, b;
        bn_fold_slots(scratch, C, c, a, b);
        bn_finalize_one(prm, f, in, c, a, b, inv_rows);
    }
}

// Call at the end of a reduce kernel, after this block's atomics: the last block to arrive
// runs bn_finalize_channels.
__device__ __forceinline__ void bn_finalize_last_block(double *scratch, const BnFinal &f) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned *ticket = reinterpret_cast<unsigned *>(scratch + 18 * (size_t)f.C);
        is_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) *reinterpret_cast<unsigned *>(scratch + 18 * (size_t)f.C) = 0u;    // ticket back to rest
    if (f.defer) {
        if (f.bwd)
            for (int c = threadIdx.x; c < f.C; c += blockDim.x) {
                double a = 0, b = 0;
#pragma unroll
                for (int s = 0; s < BN_SLOTS; ++s) {
                    a += __ldcg(scratch + (size_t)s * 2 * f.C + c);
                    b += __ldcg(scratch + (size_t)s * 2 * f.C + f.C + c);
                }
                if (f.dbeta) f.dbeta[c] += (float)a;
                if (f.dgamma) f.dgamma[c] += (float)b;
            }
        return;
    }
    bn_finalize_channels(scratch, f);
}

}  // namespace sn
