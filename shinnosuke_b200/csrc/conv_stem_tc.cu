// Tiny-C convolution (C*kh*kw <= 32: the RGB stem of config 3) on the tensor pipe, tf32 tier.
//
// The contraction is only K = 27 long, so a TMA box per filter tap would move almost nothing per
// load.  Instead the input rows a tile touches come in ONCE as a patch of the zero-padded image (a
// ring of TMA patches, see "input patches" below), builder warps assemble the COLUMN TILE from it in
// shared memory — one 128-byte row of K (padded to 32) fp32 values per output pixel, written
// directly in the swizzled layout the UMMA descriptor expects — and tcgen05.mma contracts it.  The
// column matrix never exists in HBM, and both kernels are bounded by their HBM stream (the output
// map for fprop, dY for wgrad).
//
//   fprop : D[128 px][F]  = Xcol[128 px][32] (K-major, SW128)      * W[F][32]^T (K-major, SW128)
//           warp 0 MMA, warp 1 patch producer, 8 builder warps, two sets of 8 epilogue warps that
//           take alternate tiles (TMEM -> bias/ReLU -> swizzled staging tile -> TMA store, plus the
//           BatchNorm column sums of the tile when asked).
//   wgrad : D[128 f][32k] += dY[64 px][128 f]^T (MN-major via TMA) * Xcol[64 px][32] (MN-major),
//           both in the 128B-swizzle/32B-atom layout fp32 MN-major operands need; column K of Xcol
//           is the constant 1, so D[f][K] = sum_p dY[p][f] is the bias gradient for free.
#include "common.cuh"
#include "tc_common.cuh"
#include "bn_common.cuh"

using namespace sn;
using namespace sn::tc;

namespace {

struct StemP {
    const float *x, *w, *bias;
    float *y, *dw, *db;
    int N, C, H, W, F, kh, kw, stride, ph, pw, oh, ow;
    int K;        // C*kh*kw <= 32
    int P;        // N*oh*ow
    int act;
    int tiles;    // fprop: ceil(P/128); wgrad: ceil(P/64)
    int tiles_per_cta;   // wgrad
    double *stats;       // fprop: optional BatchNorm sums of the output, stats[slot][{sum, sumsq}][F] (see conv_tc.cu)
    int Hp;              // H + 2*ph: rows of one image in padded coordinates
    int BW;              // floats per patch row: (W + 2*pw)*C rounded up to 32 (128-byte aligned TMA destinations)
    int c0;              // first column of a patch row in elements of the real row: -(pw*C rounded up to 4) (16-byte aligned box start)
    int shift;           // c0 + pw*C + shift == 0: where padded column 0 sits inside a patch row
    float inv_ow, inv_oh;
    int R128, R64;       // padded rows of a 128- / 64-pixel run that starts at column 0 inside one image (0: never happens)
};

// ---- input patches -------------------------------------------------------------------------
// The pixels of one tile are consecutive in (n, i, j) order, so their windows read a run of whole
// input rows.  A producer warp brings that run into shared memory with TMA, one box per row of the
// ZERO-PADDED image: tensor map (W*C, H, N), box (BW, 1, 1) starting at column -pw*C and row
// hp - ph, so the left/right halo columns, the top/bottom padding rows and the rows between two
// images all arrive as zeros from the TMA's out-of-bounds fill.  Patch row G - G_lo holds padded
// row G = n*(H + 2ph) + hp.  The builders then assemble im2col rows with 8 unconditional LDS each.
// (Gathering the taps from global memory, or copying the span with LDG, left the builder warps
// waiting out one L2 round trip per tile: the ring of PATCH_STAGES TMA patches hides it.)
constexpr int PATCH_STAGES = 4;

__device__ __forceinline__ int fast_div(int x, int d, float inv, int &rem) {      // exact for 0 <= x < 2^24
    int q = __float2int_rz(__int2float_rz(x) * inv);
    rem = x - q * d;
    if (rem < 0) { rem += d; --q; } else if (rem >= d) { rem -= d; ++q; }
    return q;
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// producer warp: patch of pixels [pix0, pix0 + npix) -> stage buffer `dst`, G_lo -> *info
// tmXr: the same tensor with a box of `rfull` rows — when the run of pixels starts at column 0 and
// stays inside one image (the common case: image sizes are multiples of the tile), the whole patch
// is ONE box whose rows above / below the image are zero-filled like the halo columns.
__device__ __forceinline__ void patch_produce(const StemP &p, const CUtensorMap *tmX, const CUtensorMap *tmXr, int rfull, float *dst,
                                              int *info, uint64_t *bar, int pix0, int npix, int lane) {
    const int pix1 = min(pix0 + npix, p.P) - 1;
    int j0, j, i0, i1;
    const int t0 = fast_div(pix0, p.ow, p.inv_ow, j0), n0 = fast_div(t0, p.oh, p.inv_oh, i0);
    const int t1 = fast_div(pix1, p.ow, p.inv_ow, j), n1 = fast_div(t1, p.oh, p.inv_oh, i1);
    const int g_lo = n0 * p.Hp + i0 * p.stride, g_hi = n1 * p.Hp + i1 * p.stride + p.kh - 1;
    const int rows = g_hi - g_lo + 1;
    if (lane == 0) { *info = g_lo; mbar_expect_tx(bar, (uint32_t)(rows * p.BW * 4)); }
    __syncwarp();
    if (rfull > 0 && j0 == 0 && n0 == n1 && rows == rfull) {
        if (elect_one()) tma_load_3d(dst, tmXr, bar, p.c0, i0 * p.stride - p.ph, n0);
        __syncwarp();
        return;
    }
    // warp-uniform loop, the elected lane issues: UTMALDG takes its operands from uniform registers
    int n = n0, hp = i0 * p.stride;
    for (int row = 0; row < rows; ++row) {
        if (elect_one()) tma_load_3d(dst + row * p.BW, tmX, bar, p.c0, hp - p.ph, n);
        if (++hp == p.Hp) { hp = 0; ++n; }
    }
    __syncwarp();
}

// A builder thread owns 8 consecutive k's of one pixel row; their patch offsets are loop invariant.
struct KSlice {
    int off[8];          // (u*BW + v*C + c) for k = 8*kq + e, or -1 when k >= K
};
__device__ __forceinline__ void make_kslice(const StemP &p, int kq, KSlice &ks) {
    const int kwC = p.kw * p.C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = kq * 8 + e;
        if (k < p.K) {
            const int u = k / kwC, r = k - u * kwC;          // r = v*C + c
            ks.off[e] = u * p.BW + r;
        } else {
            ks.off[e] = -1;
        }
    }
}
// the 8 column values of pixel `pix` for this slice, read from the patch; `one_at` (>= 0) plants the
// constant-one column.  Rows past the end of the tensor are all zero.
__device__ __forceinline__ void build_slice(const StemP &p, const KSlice &ks, const float *patch, int g_lo, int pix, int kq,
                                            int one_at, float (&col)[8]) {
    const bool live = pix < p.P;
    int j, i;
    const int t = fast_div(live ? pix : 0, p.ow, p.inv_ow, j), n = fast_div(t, p.oh, p.inv_oh, i);
    const float *base = patch + (n * p.Hp + i * p.stride - g_lo) * p.BW + j * p.stride * p.C + p.shift;
#pragma unroll
    for (int e = 0; e < 8; ++e)
        col[e] = (live && ks.off[e] >= 0) ? base[ks.off[e]] : ((live && kq * 8 + e == one_at) ? 1.f : 0.f);
}

// ============================================================================== fprop
constexpr int SF_BUILDERS = 256;               // 2 threads per pixel row, 16 k's (two 32-byte slices) each
// Two SETS of eight epilogue warps (two per TMEM lane quarter) take alternate tiles = alternate TMEM
// accumulators: a warp's tile is a serial chain (tcgen05.ld, bias/ReLU, staging, TMA store, column
// sums for the BatchNorm statistics) of ~2500 cycles + ~800 for the statistics, and with one set that
// chain WAS the tile time; with two sets each has two tile times for it.
constexpr int SF_EPI_WARPS = 16;
constexpr int SF_THREADS = 64 + SF_BUILDERS + 32 * SF_EPI_WARPS;   // warp 0: MMA, warp 1: patch producer, warps 2-9: column builders, warps 10-25: epilogue
constexpr int SF_PATCH_FLOATS = 4096;          // per patch stage; stem_tc_eligible bounds the rows by this
constexpr int SF_ABUF = 3;
constexpr int SF_A_BYTES = 128 * 128;

// OBF16: the output map is stored in bf16 (see conv_tc.cu): 64-byte staging rows, statistics of the rounded values
template <int BN, bool OBF16 = false>
__global__ void __launch_bounds__(SF_THREADS, 1) stem_fprop_kernel(const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX,
                                                                   const __grid_constant__ CUtensorMap tmXr, const StemP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem;                                   // [BN][128 B], SW128 K-major
    uint8_t *sA = smem + BN * 128;                        // SF_ABUF x [128][128 B]
    uint8_t *epi_stage = sA + SF_ABUF * SF_A_BYTES;       // per epilogue warp: one 32 rows x 128 B staging tile for its TMA stores
    float *sstat = reinterpret_cast<float *>(epi_stage + SF_EPI_WARPS * 4096);            // [set][quarter][{sum, sumsq}][BN] column sums
    float *patch = sstat + 8 * 2 * BN;                                                        // PATCH_STAGES x SF_PATCH_FLOATS
    uint64_t *a_full = reinterpret_cast<uint64_t *>(patch + PATCH_STAGES * SF_PATCH_FLOATS);
    uint64_t *a_empty = a_full + SF_ABUF;
    uint64_t *t_full = a_empty + SF_ABUF;
    uint64_t *t_empty = t_full + 2;
    uint64_t *p_full = t_empty + 2;
    uint64_t *p_empty = p_full + PATCH_STAGES;
    int *pinfo = reinterpret_cast<int *>(p_empty + PATCH_STAGES);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(pinfo + PATCH_STAGES);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    pdl_wait();        // (the weights below are global memory the previous kernels may have written)
    // weights -> smem in the swizzled K-major layout: row f, 16-byte chunk j at (j ^ (f & 7)) * 16
    for (int i = threadIdx.x; i < BN * 32; i += SF_THREADS) {
        const int f = i >> 5, k = i & 31;
        float v = (f < p.F && k < p.K) ? __ldg(p.w + (long long)f * p.K + k) : 0.f;
        const int chunk = (k >> 2) ^ (f & 7);
        *reinterpret_cast<float *>(sB + f * 128 + chunk * 16 + (k & 3) * 4) = v;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < SF_ABUF; ++i) { mbar_init(&a_full[i], SF_BUILDERS / 32); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], SF_EPI_WARPS / 2); }
        for (int i = 0; i < PATCH_STAGES; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], SF_BUILDERS / 32); }
        tma_prefetch_desc(&tmX);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(tmem_ptr, 2 * BN < 32 ? 32 : 2 * BN); tmem_relinquish(); }
    for (int i = threadIdx.x; i < 8 * 2 * BN; i += SF_THREADS) sstat[i] = 0.f;
    fence_proxy_async();            // the weight tile was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);     // shuffle from lane 0: provably warp-uniform

    if (warp == 0) {
        {   // warp-uniform control flow; the elected lane issues (operands stay in uniform registers)
            constexpr uint32_t idesc = make_idesc(2, 128, BN, 0, 0);
            const uint64_t bdesc = make_smem_desc(smem_u32(sB), 0, 1024);
            int buf = 0; uint32_t phase = 0; int iter = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
                const int acc = iter & 1;
                mbar_wait(&t_empty[acc], ((iter >> 1) & 1) ^ 1);
                mbar_wait(&a_full[buf], phase);
                tc_fence_after();
                const uint64_t adesc = make_smem_desc(smem_u32(sA + buf * SF_A_BYTES), 0, 1024);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma_tf32(tmem_base + acc * BN, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
                    mma_commit(&a_empty[buf]);
                    mma_commit(&t_full[acc]);
                }
                __syncwarp();
                if (++buf == SF_ABUF) { buf = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- patch producer: runs PATCH_STAGES tiles ahead of the builders ----------------
        int ps = 0; uint32_t pphase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            mbar_wait(&p_empty[ps], pphase ^ 1);
            patch_produce(p, &tmX, &tmXr, p.R128, patch + ps * SF_PATCH_FLOATS, pinfo + ps, &p_full[ps], tile * 128, 128, lane);
            if (++ps == PATCH_STAGES) { ps = 0; pphase ^= 1; }
        }
    } else if (warp < 2 + SF_BUILDERS / 32) {
        // ---------------- column builders: thread (r, h) builds k = 16h..16h+15 of pixel row r ----------------
        const int bt = threadIdx.x - 64;
        const int r = bt >> 1, kq0 = (bt & 1) * 2;
        KSlice ks0, ks1;
        make_kslice(p, kq0, ks0);
        make_kslice(p, kq0 + 1, ks1);
        int buf = 0; uint32_t phase = 0;
        int ps = 0; uint32_t pphase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            mbar_wait(&p_full[ps], pphase);
            float col[8], col2[8];
            build_slice(p, ks0, patch + ps * SF_PATCH_FLOATS, pinfo[ps], tile * 128 + r, kq0, -1, col);
            build_slice(p, ks1, patch + ps * SF_PATCH_FLOATS, pinfo[ps], tile * 128 + r, kq0 + 1, -1, col2);
            mbar_wait(&a_empty[buf], phase ^ 1);
            uint8_t *row = sA + buf * SF_A_BYTES + r * 128;
            // SW128 K-major: 16-byte chunk j of row r lives at (j ^ (r & 7)) * 16
            *reinterpret_cast<float4 *>(row + (((2 * kq0) ^ (r & 7)) * 16)) = make_float4(col[0], col[1], col[2], col[3]);
            *reinterpret_cast<float4 *>(row + (((2 * kq0 + 1) ^ (r & 7)) * 16)) = make_float4(col[4], col[5], col[6], col[7]);
            *reinterpret_cast<float4 *>(row + (((2 * kq0 + 2) ^ (r & 7)) * 16)) = make_float4(col2[0], col2[1], col2[2], col2[3]);
            *reinterpret_cast<float4 *>(row + (((2 * kq0 + 3) ^ (r & 7)) * 16)) = make_float4(col2[4], col2[5], col2[6], col2[7]);
            fence_proxy_async();               // make the generic-proxy writes visible to the tensor core
            __syncwarp();                      // every lane's stores are fenced, its patch reads complete
            if (lane == 0) { mbar_arrive(&a_full[buf]); mbar_arrive(&p_empty[ps]); }      // ONE arrival per warp (see stem_wgrad_kernel)
            if (++buf == SF_ABUF) { buf = 0; phase ^= 1; }
            if (++ps == PATCH_STAGES) { ps = 0; pphase ^= 1; }
        }
    } else {
        // ---------------- epilogue: TMEM -> bias/ReLU -> NHWC output ----------------
        const int q = warp & 3;
        const int ew = warp - 2 - SF_BUILDERS / 32;          // 0..15
        const int set = ew >> 3;                             // which tiles (= which TMEM accumulator) this warp takes
        const int half = (ew >> 2) & 1;                      // which 32-column chunks this warp takes
        float *mystat = sstat + (set * 4 + q) * 2 * BN;      // shared with the other half's warp only (disjoint columns)
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
            const int acc = iter & 1;
            if (acc != set) continue;
            mbar_wait(&t_full[acc], (iter >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = half * 32; c0 < BN; c0 += 64) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), rr);
                tmem_ld_wait();
                if (c0 >= p.F) continue;
                uint8_t *st = epi_stage + ew * 4096;
                if (lane == 0) tma_store_wait_read();               // this warp's previous store has drained the tile
                __syncwarp();
                if constexpr (OBF16) {
                    float v[32];
#pragma unroll
                    for (int jj = 0; jj < 32; jj += 4) {
                        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias && c0 + jj < p.F) b = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + jj));
                        v[jj] = __uint_as_float(rr[jj]) + b.x; v[jj + 1] = __uint_as_float(rr[jj + 1]) + b.y;
                        v[jj + 2] = __uint_as_float(rr[jj + 2]) + b.z; v[jj + 3] = __uint_as_float(rr[jj + 3]) + b.w;
                        if (p.act == SN_ACT_RELU) {
                            v[jj] = relu_keep_sign(v[jj]); v[jj + 1] = relu_keep_sign(v[jj + 1]);
                            v[jj + 2] = relu_keep_sign(v[jj + 2]); v[jj + 3] = relu_keep_sign(v[jj + 3]);
                        }
                    }
                    stage_row_bf16(st, lane, v);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&tmO, st, c0, tile * 128 + q * 32); tma_store_commit(); }
                    if (p.stats)
                        staged_tile_colsums_bf16(st, lane, p.P - (tile * 128 + q * 32), mystat + c0, mystat + BN + c0, p.F - c0);
                    continue;
                }
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4) {
                    float4 v = make_float4(__uint_as_float(rr[jj]), __uint_as_float(rr[jj + 1]), __uint_as_float(rr[jj + 2]),
                                           __uint_as_float(rr[jj + 3]));
                    if (p.bias && c0 + jj < p.F) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + jj));
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                    }
                    if (p.act == SN_ACT_RELU) { v.x = relu_keep_sign(v.x); v.y = relu_keep_sign(v.y); v.z = relu_keep_sign(v.z); v.w = relu_keep_sign(v.w); }
                    *reinterpret_cast<float4 *>(st + lane * 128 + (((jj >> 2) ^ (lane & 7)) * 16)) = v;
                }
                fence_proxy_async();
                __syncwarp();
                // one TMA store per 32 x 32 sub-tile: full 128-byte lines, clipped at the tensor bounds
                if (lane == 0) { tma_store_2d(&tmO, st, c0, tile * 128 + q * 32); tma_store_commit(); }
                if (p.stats)
                    staged_tile_colsums(st, lane, p.P - (tile * 128 + q * 32), mystat + c0, mystat + BN + c0, p.F - c0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN < 32 ? 32 : 2 * BN); }
    if (p.stats) {
        double *slot = p.stats + (size_t)(blockIdx.x % BN_SLOTS) * 2 * p.F;
        for (int f = threadIdx.x; f < p.F; f += SF_THREADS) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { s1 += (double)sstat[e * 2 * BN + f]; s2 += (double)sstat[(e * 2 + 1) * BN + f]; }
            atomicAdd(slot + f, s1);
            atomicAdd(slot + p.F + f, s2);
        }
    }
}

// ============================================================================== wgrad
constexpr int SW_BUILDERS = 256;               // 4 threads per pixel row build one stage ...
constexpr int SW_GROUPS = 3;                   // ... and three such groups take stages in turn: a builder warp's stage is a
                                               // ~200-instruction dependent chain (measured ~2700 cycles with mbarrier waits and
                                               // LDS latencies: two groups gave 1350 cycles per stage, the whole kernel's pace)
constexpr int SW_THREADS = 64 + SW_GROUPS * SW_BUILDERS + 128;   // warp 0: TMA (dY) + MMA, warp 1: patch producer, warps 2-17: column builders, warps 18-21: epilogue
constexpr int SW_PB = 64;                      // pixels per stage
// NA = 32-filter atoms of dY actually loaded per stage (2 when F <= 64, else 4).  The MMA is always
// M = 128: with NA = 2 its upper 64 rows read whatever follows in shared memory (the next stage /
// the patch ring) and produce accumulator rows the epilogue never looks at — rows of D depend only
// on the same rows of A — which halves the stage and deepens the ring from 4 to 6 stages.
// PATCHES: depth of the patch ring.  A patch slot is released when the builders of its stage are done and
// its refill is a TMA round trip (~2700 cycles under load); with the 4 slots the fprop kernel uses, a
// builder group found its next patch still in flight (ncu: 20 % of all warp samples on the p_full wait,
// 1350 cycles per stage against ~500 of work) — 8 slots keep the producer four group-stages ahead.
// (A 10-deep ring for the half-size bf16 stages — more dY bytes in flight per SM — changed nothing: 41.4 us either way.)
template <int NA> struct SwCfg {
    static constexpr int STAGES = NA <= 2 ? 6 : 4;
    static constexpr int PATCHES = NA <= 2 ? 8 : 6;
};
constexpr int SW_MAX_STAGES = 6;
constexpr int SW_ATOM = SW_PB * 128;           // one 32-wide atom column: 64 pixel rows x 128 B
constexpr int SW_PATCH_FLOATS = 2048;          // per patch stage (a 64-pixel stage spans about half the rows of a 128-pixel tile)

// BF16: dY arrives as the bf16 copy the fused BN+pool backward wrote (the fp32 map is then never written): atoms of
// 64 filters (NA = 1 when F <= 64), plain 128-byte swizzle, 16-pixel K-steps; the builders round the patch values to
// bf16 into the first 64 bytes of the same 128-byte rows (N = 64 for the MMA: columns 32..63 are whatever the rows
// hold and land in accumulator columns nobody reads).
template <int NA, bool BF16>
__global__ void __launch_bounds__(SW_THREADS, 1) stem_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                                                                   const __grid_constant__ CUtensorMap tmXr, const StemP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int SW_STAGES = SwCfg<NA>::STAGES;
    constexpr int SW_PATCHES = SwCfg<NA>::PATCHES;
    constexpr int SW_A_BYTES = NA * SW_ATOM;
    constexpr int SW_STAGE_BYTES = SW_A_BYTES + SW_ATOM;
    float *patch = reinterpret_cast<float *>(smem + SW_STAGES * SW_STAGE_BYTES);        // SW_PATCHES x SW_PATCH_FLOATS
    uint64_t *full = reinterpret_cast<uint64_t *>(patch + SW_PATCHES * SW_PATCH_FLOATS);
    uint64_t *empty = full + SW_MAX_STAGES;
    uint64_t *acc_full = empty + SW_MAX_STAGES;
    uint64_t *p_full = acc_full + 1;
    uint64_t *p_empty = p_full + SW_PATCHES;
    int *pinfo = reinterpret_cast<int *>(p_empty + SW_PATCHES);
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(pinfo + SW_PATCHES);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int f_tile = blockIdx.y;
    const int t0 = blockIdx.x * p.tiles_per_cta;
    const int t1 = min(p.tiles, t0 + p.tiles_per_cta);
    const int nst = t1 - t0;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmDY);
        // full: 1 arrive.expect_tx by the TMA thread + one arrival per builder warp
        for (int i = 0; i < SW_STAGES; ++i) { mbar_init(&full[i], 1 + SW_BUILDERS / 32); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        for (int i = 0; i < SW_PATCHES; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], SW_BUILDERS / 32); }
        tma_prefetch_desc(&tmX);
        fence_barrier_init();
    }
    constexpr uint32_t SW_TMEM_COLS = BF16 ? 64 : 32;
    if (warp == 0) { tmem_alloc(tmem_ptr, SW_TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);     // shuffle from lane 0: provably warp-uniform
    pdl_wait();

    if (nst > 0) {
        if (warp == 0) {
            {
                // one warp is both the TMA producer and the MMA issuer (the MMAs are negligible here): it runs the
                // loads SW_STAGES - 1 ahead of the MMAs.  Warp-uniform control flow, the elected lane issues.
                constexpr uint32_t idesc = BF16 ? make_idesc(1, 128, 64, 1, 1) : make_idesc(2, 128, 32, 1, 1);
                constexpr int A_ELEMS = BF16 ? 64 : 32;          // filters per dY atom
                int ld = 0, ld_stage = 0; uint32_t ld_phase = 0;
                int stage = 0; uint32_t phase = 0;
                for (int s = 0; s < nst; ++s) {
                    // SW_STAGES - 1 ahead, not SW_STAGES: the slot refilled now is the one stage s-2 used, whose MMAs
                    // completed long ago.  Refilling the slot of stage s-1 made this warp wait for the MMAs it had
                    // just issued before it could even look at stage s: one MMA latency + one TMA issue per stage,
                    // serially — the pace of the kernel (1380 cycles per stage whatever builders / rings did).
                    while (ld < nst && ld < s + SW_STAGES - 1) {
                        mbar_wait(&empty[ld_stage], ld_phase ^ 1);
                        uint8_t *sa = smem + ld_stage * SW_STAGE_BYTES;
                        if (elect_one()) {
                            mbar_expect_tx(&full[ld_stage], SW_A_BYTES);
#pragma unroll
                            for (int a = 0; a < NA; ++a)
                                tma_load_2d(sa + a * SW_ATOM, &tmDY, &full[ld_stage], f_tile * 128 + a * A_ELEMS, (t0 + ld) * SW_PB);
                        }
                        __syncwarp();
                        ++ld;
                        if (++ld_stage == SW_STAGES) { ld_stage = 0; ld_phase ^= 1; }
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    // descriptors = a precomputed template | (start address >> 4): one add per MMA operand
                    const uint32_t sa16 = smem_u32(smem + stage * SW_STAGE_BYTES) >> 4;
                    if constexpr (BF16) {
                        const uint64_t tmpl = make_smem_desc(0, SW_ATOM, 1024, LAYOUT_SW128);
#pragma unroll
                        for (int k = 0; k < SW_PB / 16; ++k) {       // K-step = 16 pixel rows = two 8-row K groups 1024 B apart
                            const uint64_t adesc = tmpl | (uint64_t)(sa16 + k * 128);
                            const uint64_t bdesc = tmpl | (uint64_t)(sa16 + (SW_A_BYTES >> 4) + k * 128);
                            if (elect_one()) mma_bf16(tmem_base, adesc, bdesc, idesc, (s | k) != 0 ? 1u : 0u);
                        }
                    } else {
                        const uint64_t tmpl = make_smem_desc(0, SW_ATOM, 512, LAYOUT_SW128_BASE32B);
#pragma unroll
                        for (int k = 0; k < SW_PB / 8; ++k) {
                            const uint64_t adesc = tmpl | (uint64_t)(sa16 + k * 64);
                            const uint64_t bdesc = tmpl | (uint64_t)(sa16 + (SW_A_BYTES >> 4) + k * 64);
                            if (elect_one()) mma_tf32(tmem_base, adesc, bdesc, idesc, (s | k) != 0 ? 1u : 0u);
                        }
                    }
                    if (elect_one()) mma_commit(&empty[stage]);
                    __syncwarp();
                    if (++stage == SW_STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) mma_commit(acc_full);
                __syncwarp();
            }
        } else if (warp == 1) {
            // patch producer: SW_PATCHES stages ahead of the builders.  The geometry of a stage (two divisions by ow, two
            // by oh, the row span) is a ~100-instruction dependent chain full of I2F / F2I / R2UR round trips: done per
            // stage by this one warp it was a serial ~1000 cycles.  Each LANE now works out one of the next 32 stages,
            // and the per-stage work is five shuffles, the expect_tx and the TMA.
            int ps = 0; uint32_t pphase = 0;
            for (int sb = 0; sb < nst; sb += 32) {
                int g_lo = 0, rows = 0, i0s = 0, n0 = 0, fast = 0;
                if (sb + lane < nst) {
                    const int pix0 = (t0 + sb + lane) * SW_PB, pix1 = min(pix0 + SW_PB, p.P) - 1;
                    int j0, j, i0, i1;
                    const int q0 = fast_div(pix0, p.ow, p.inv_ow, j0);
                    n0 = fast_div(q0, p.oh, p.inv_oh, i0);
                    const int q1 = fast_div(pix1, p.ow, p.inv_ow, j), n1 = fast_div(q1, p.oh, p.inv_oh, i1);
                    g_lo = n0 * p.Hp + i0 * p.stride;
                    rows = n1 * p.Hp + i1 * p.stride + p.kh - 1 - g_lo + 1;
                    i0s = i0 * p.stride;
                    fast = (p.R64 > 0 && j0 == 0 && n0 == n1 && rows == p.R64) ? 1 : 0;
                }
                const int cnt = min(32, nst - sb);
                for (int l = 0; l < cnt; ++l) {
                    const int g_lo_s = __shfl_sync(0xffffffffu, g_lo, l), rows_s = __shfl_sync(0xffffffffu, rows, l);
                    const int i0_s = __shfl_sync(0xffffffffu, i0s, l), n0_s = __shfl_sync(0xffffffffu, n0, l);
                    const int fast_s = __shfl_sync(0xffffffffu, fast, l);
                    mbar_wait(&p_empty[ps], pphase ^ 1);
                    float *dst = patch + ps * SW_PATCH_FLOATS;
                    if (lane == 0) { pinfo[ps] = g_lo_s; mbar_expect_tx(&p_full[ps], (uint32_t)(rows_s * p.BW * 4)); }
                    __syncwarp();
                    if (fast_s) {
                        if (elect_one()) tma_load_3d(dst, &tmXr, &p_full[ps], p.c0, i0_s - p.ph, n0_s);
                    } else {
                        int n = n0_s, hp = i0_s;          // one box per padded row (tiles that straddle images / start mid-row)
                        for (int row = 0; row < rows_s; ++row) {
                            if (elect_one()) tma_load_3d(dst + row * p.BW, &tmX, &p_full[ps], p.c0, hp - p.ph, n);
                            if (++hp == p.Hp) { hp = 0; ++n; }
                        }
                    }
                    __syncwarp();
                    if (++ps == SW_PATCHES) { ps = 0; pphase ^= 1; }
                }
            }
        } else if (warp < 2 + SW_GROUPS * SW_BUILDERS / 32) {
            // column builders: thread (r, kq) of group g builds k = 8kq..8kq+7 (one 32-byte chunk) of pixel
            // row r of the stages s = g, g + SW_GROUPS, ...
            const int grp = (threadIdx.x - 64) / SW_BUILDERS;
            const int bt = (threadIdx.x - 64) - grp * SW_BUILDERS;
            const int r = bt >> 2, kq = bt & 3;
            KSlice ks;
            make_kslice(p, kq, ks);
            const int one_at = p.db ? p.K : -1;
            for (int s = grp; s < nst; s += SW_GROUPS) {
                const int stage = s % SW_STAGES, ps = s % SW_PATCHES;
                const uint32_t phase = (s / SW_STAGES) & 1, pphase = (s / SW_PATCHES) & 1;
                mbar_wait(&p_full[ps], pphase);
                float col[8];
                build_slice(p, ks, patch + ps * SW_PATCH_FLOATS, pinfo[ps], (t0 + s) * SW_PB + r, kq, one_at, col);   // rows past the end stay 0 (dY is TMA zero-filled too)
                mbar_wait(&empty[stage], phase ^ 1);
                if constexpr (BF16) {
                    // plain 128B swizzle: 16-byte chunk j (8 bf16) of row r lives at (j ^ (r & 7)) * 16
                    uint8_t *dst = smem + stage * SW_STAGE_BYTES + SW_A_BYTES + r * 128 + ((kq ^ (r & 7)) * 16);
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(pack_bf16x2(col[0], col[1]), pack_bf16x2(col[2], col[3]),
                                                                 pack_bf16x2(col[4], col[5]), pack_bf16x2(col[6], col[7]));
                } else {
                    // 128B swizzle with 32-byte atoms: 32-byte chunk j of row r lives at (j ^ (r & 3)) * 32
                    uint8_t *dst = smem + stage * SW_STAGE_BYTES + SW_A_BYTES + r * 128 + ((kq ^ (r & 3)) * 32);
                    *reinterpret_cast<float4 *>(dst) = make_float4(col[0], col[1], col[2], col[3]);
                    *reinterpret_cast<float4 *>(dst + 16) = make_float4(col[4], col[5], col[6], col[7]);
                }
                // ONE arrival per warp: 256 per-thread arrivals on one mbarrier serialise in the shared-memory atomic
                // unit (~4 cycles each) and were the pace of this kernel — ~1380 cycles per stage whatever the
                // number of builder groups, the depth of the rings or the width of dY
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&full[stage]); mbar_arrive(&p_empty[ps]); }
            }
        } else {
            // epilogue (last four warps): dW[f][k] += D[f][k], db[f] += D[f][K]
            const int q = warp & 3;
            mbar_wait_sleep(acc_full, 0);        // whole-kernel wait: do not steal issue slots from the builders
            tc_fence_after();
            uint32_t rr[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), rr);
            tmem_ld_wait();
            const int f = f_tile * 128 + q * 32 + lane;
            if (f < p.F) {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    if (k < p.K) atomicAdd(p.dw + (long long)f * p.K + k, __uint_as_float(rr[k]));
                    else if (k == p.K && p.db) atomicAdd(p.db + f, __uint_as_float(rr[k]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, SW_TMEM_COLS); }
}

// one patch row = one TMA box: its shared-memory address must be 128-byte aligned, so the pitch is a multiple of 32 floats
int patch_row_floats(int W, int C, int pw) { return ((((pw * C + 3) & ~3) - pw * C) + (W + 2 * pw) * C + 31) & ~31; }

void fill_geom(StemP &p, int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw) {
    p.N = N; p.C = C; p.H = H; p.W = W; p.F = F; p.kh = kh; p.kw = kw; p.stride = stride; p.ph = ph; p.pw = pw;
    p.oh = (H + 2 * ph - kh) / stride + 1; p.ow = (W + 2 * pw - kw) / stride + 1;
    p.K = C * kh * kw; p.P = N * p.oh * p.ow;
    p.Hp = H + 2 * ph; p.BW = patch_row_floats(W, C, pw);
    p.c0 = -((pw * C + 3) & ~3); p.shift = -p.c0 - pw * C;
    // a run of npx pixels from column 0 inside one image spans (npx/ow - 1)*stride + kh padded rows
    auto full_rows = [&](int npx) {
        if (npx % p.ow || (p.oh * p.ow) % npx) return 0;
        const int r = (npx / p.ow - 1) * stride + kh;
        return r <= 256 ? r : 0;
    };
    p.R128 = full_rows(128); p.R64 = full_rows(64);
    p.inv_ow = 1.f / (float)p.ow; p.inv_oh = 1.f / (float)p.oh;
}

// x as (W*C, H, N) with one-row boxes of BW floats, no swizzle: see "input patches" above
int make_patch_tmap(CUtensorMap *tm, const StemP &p, int rows = 1) {
    long long dims[3] = {(long long)p.W * p.C, p.H, p.N};
    long long strs[3] = {1, (long long)p.W * p.C, (long long)p.H * p.W * p.C};
    int box[3] = {p.BW, rows > 0 ? rows : 1, 1};
    return make_tmap_linear_f32(tm, p.x, 3, dims, strs, box);
}

template <int BN, bool OBF16>
int launch_stem_fprop(const StemP &p, cudaStream_t st) {
    const int smem = BN * 128 + SF_ABUF * SF_A_BYTES + SF_EPI_WARPS * 4096 + 8 * 2 * BN * 4 + PATCH_STAGES * SF_PATCH_FLOATS * 4 + 1024 + 256;
    CUtensorMap tmO, tmX, tmXr;
    int ex = make_patch_tmap(&tmX, p);
    if (ex) return ex;
    ex = make_patch_tmap(&tmXr, p, p.R128);
    if (ex) return ex;
    long long odims[2] = {p.F, p.P};
    long long ostr[2] = {1, p.F};
    int obox[2] = {32, 32};
    int e = OBF16 ? make_tmap_bf16_sw64(&tmO, p.y, 2, odims, ostr, obox) : make_tmap_f32(&tmO, p.y, 2, odims, ostr, obox);
    if (e) return e;
    static bool configured_dev[16] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        SN_CUDA(cudaFuncSetAttribute(stem_fprop_kernel<BN, OBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    int grid = p.tiles < num_sms() ? p.tiles : num_sms();
    SN_CUDA(launch_dependent(stem_fprop_kernel<BN, OBF16>, dim3(grid), dim3(SF_THREADS), (size_t)smem, st, tmO, tmX, tmXr, p));
    return after_launch("stem_fprop_kernel");
}

}  // namespace

namespace sn {

bool stem_tc_eligible(int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw) {
    const int K = C * kh * kw;
    if (K > 32 || (F & 15) || F > 256 || F < 16) return false;
    const int OH = (H + 2 * ph - kh) / stride + 1, OW = (W + 2 * pw - kw) / stride + 1;
    const long long P = (long long)N * OH * OW;
    // TMA patch rows: 16-byte row pitch in global memory, box <= 256 elements, pixel indices exact in
    // fp32 (fast_div), and the padded rows that 128 (fprop) / 64 (wgrad) consecutive pixels touch —
    // image boundary included — must fit a patch stage
    const int bw = patch_row_floats(W, C, pw);
    if (((W * C) & 3) || bw > 256 || P >= (1 << 24)) return false;
    const long long extra = kh + 2 * ph + stride;
    const long long rows128 = (long long)(127 / OW + 2) * stride + extra, rows64 = (long long)(63 / OW + 2) * stride + extra;
    if (rows128 * bw > SF_PATCH_FLOATS || rows64 * bw > SW_PATCH_FLOATS) return false;
    return P >= 4096;          // small maps: the direct kernel is latency-equivalent
}

int conv_stem_fprop_tc(const float *x, const float *w, const float *bias, void *y, int N, int C, int H, int W, int F, int kh,
                       int kw, int stride, int ph, int pw, int act, double *stats, cudaStream_t st, bool out_bf16) {
    StemP p{};
    fill_geom(p, N, C, H, W, F, kh, kw, stride, ph, pw);
    p.x = x; p.w = w; p.bias = bias; p.y = static_cast<float *>(y); p.act = act;      // (a bf16 map when out_bf16: only the tensor map reads p.y)
    p.stats = stats;              // zero on entry (caller's contract): no memset node in the step graph
    p.tiles = (p.P + 127) / 128;
    if (out_bf16) {
        if (F & 7) { set_error("conv_stem_fprop_tc: a bf16 output map needs F %% 8 == 0"); return SN_ERR_UNSUPPORTED; }
        if (F <= 32) return launch_stem_fprop<32, true>(p, st);
        if (F <= 64) return launch_stem_fprop<64, true>(p, st);
        if (F <= 128) return launch_stem_fprop<128, true>(p, st);
        return launch_stem_fprop<256, true>(p, st);
    }
    if (F <= 32) return launch_stem_fprop<32, false>(p, st);
    if (F <= 64) return launch_stem_fprop<64, false>(p, st);
    if (F <= 128) return launch_stem_fprop<128, false>(p, st);
    return launch_stem_fprop<256, false>(p, st);
}

// dw/db accumulated into (+=).  Requires K < 32 when db != NULL (the ones column).
// dy_bf16: dy is the bf16 copy of the gradient map (bf16 tier, F % 8 == 0); x stays fp32 (the network input).
int conv_stem_wgrad_tc(const float *x, const void *dy, float *dw, float *db, int N, int C, int H, int W, int F, int kh, int kw,
                       int stride, int ph, int pw, cudaStream_t st, bool dy_bf16) {
    StemP p{};
    fill_geom(p, N, C, H, W, F, kh, kw, stride, ph, pw);
    p.x = x; p.dw = dw; p.db = db;
    p.tiles = (p.P + SW_PB - 1) / SW_PB;
    const int f_tiles = (F + 127) / 128;
    int ctas = num_sms() / f_tiles;
    if (ctas < 1) ctas = 1;
    if (ctas > p.tiles) ctas = p.tiles;
    p.tiles_per_cta = (p.tiles + ctas - 1) / ctas;
    ctas = (p.tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    CUtensorMap tmDY;
    long long ddims[2] = {F, p.P};
    long long dstr[2] = {1, F};
    int dbox[2] = {dy_bf16 ? 64 : 32, SW_PB};
    int e = dy_bf16 ? make_tmap(&tmDY, dy, 2, ddims, dstr, dbox, false, true) : make_tmap_f32(&tmDY, dy, 2, ddims, dstr, dbox, true);
    if (e) return e;
    const bool small = F <= 64;                          // half the dY atoms per stage, deeper rings
    const int na = dy_bf16 ? (small ? 1 : 2) : (small ? 2 : 4);
    const int smem = (na <= 2 ? SwCfg<2>::STAGES : SwCfg<4>::STAGES) * (na + 1) * SW_ATOM +
                     (na <= 2 ? SwCfg<2>::PATCHES : SwCfg<4>::PATCHES) * SW_PATCH_FLOATS * 4 + 1024 + 512;
    CUtensorMap tmX, tmXr;
    e = make_patch_tmap(&tmX, p);
    if (e) return e;
    e = make_patch_tmap(&tmXr, p, p.R64);
    if (e) return e;
    static bool configured_dev[16][4] = {};
    bool &configured = configured_dev[current_device_slot()][(dy_bf16 ? 2 : 0) + (small ? 1 : 0)];
#define SN_STEM_WGRAD_GO(NA, BF)                                                                                             \
    do {                                                                                                                    \
        if (!configured) {                                                                                                  \
            SN_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel<NA, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
            configured = true;                                                                                              \
        }                                                                                                                   \
        SN_CUDA(launch_dependent(stem_wgrad_kernel<NA, BF>, dim3(ctas, f_tiles), dim3(SW_THREADS), (size_t)smem, st, tmDY, tmX, tmXr, p)); \
        return after_launch("stem_wgrad_kernel");                                                                           \
    } while (0)
    if (dy_bf16) { if (small) SN_STEM_WGRAD_GO(1, true); else SN_STEM_WGRAD_GO(2, true); }
    if (small) SN_STEM_WGRAD_GO(2, false); else SN_STEM_WGRAD_GO(4, false);
#undef SN_STEM_WGRAD_GO
}

}  // namespace sn
