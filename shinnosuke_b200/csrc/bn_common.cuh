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
__device__ __forceinline__ void bn_finalize_channels(double *scratch, const BnFinal &f) {
    const int C = f.C;
    float *prm = bn_params(scratch, C);
    // No float64 divide / sqrt below (each is a long dependent chain on this part's few FP64 units,
    // and this code is the serial tail of a kernel): 1/rows comes from a float seed refined by two
    // Newton steps, 1/sqrt from rsqrtf refined by two — both exact to float64 rounding.
    double inv_rows;
    {
        const double r = (double)f.rows;
        double y = (double)(1.0f / (float)f.rows);
        y = y * (2.0 - r * y);
        inv_rows = y * (2.0 - r * y);
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a = 0, b = 0;
#pragma unroll
        for (int s = 0; s < BN_SLOTS; ++s) {
            a += __ldcg(scratch + (size_t)s * 2 * C + c);
            b += __ldcg(scratch + (size_t)s * 2 * C + C + c);
            // consumed: leave the sums zero again, so that a producer that only ADDS into the scratch
            // (the convolution epilogues, sn_conv2d_fprop_stats) needs no memset before it
            scratch[(size_t)s * 2 * C + c] = 0.0;
            scratch[(size_t)s * 2 * C + C + c] = 0.0;
        }
        if (!f.bwd) {
            const double pivot = f.x ? (double)f.x[c] : 0.0;
            const double m1 = a * inv_rows;                             // E[x - pivot]
            double var = b * inv_rows - m1 * m1;
            if (var < 0) var = 0;
            const double m = pivot + m1;
            const double ve = var + (double)f.eps;
            double is = (double)rsqrtf((float)ve);
            is = is * (1.5 - 0.5 * ve * is * is);
            is = is * (1.5 - 0.5 * ve * is * is);
            const double g = (double)f.gamma[c];
            prm[c] = (float)(g * is);
            prm[C + c] = (float)((double)f.beta[c] - g * is * m);
            f.save_mean[c] = (float)m;
            f.save_invstd[c] = (float)is;
            f.moving_mean[c] = f.momentum * f.moving_mean[c] + (1.f - f.momentum) * (float)m;
            f.moving_var[c] = f.momentum * f.moving_var[c] + (1.f - f.momentum) * (float)var;
        } else {
            const double is = (double)f.invstd[c], mu = (double)f.mean[c];
            const double k2 = a * inv_rows, k3 = b * inv_rows;
            const double p0 = (double)f.gamma[c] * is;
            const double p1 = -p0 * k3 * is;
            prm[c] = (float)p0;
            prm[C + c] = (float)p1;
            prm[2 * C + c] = (float)(-p0 * k2 - p1 * mu);
            if (f.dbeta) f.dbeta[c] += (float)a;
            if (f.dgamma) f.dgamma[c] += (float)b;
        }
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
