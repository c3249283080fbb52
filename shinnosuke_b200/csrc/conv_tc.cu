// Tensor-core implicit-GEMM convolution for sm_100a: tcgen05.mma (kind::tf32, fp32 operands read
// as TF32, fp32 accumulators in TMEM) fed by TMA, no column matrix.
//
// fprop / dgrad / Dense forward+dX  ("K-major" kernel, conv_igemm_kernel):
//     D[p][f] = sum_{tap=(u,v)} sum_{c} X[n, h+u-ph, w+v-pw, c] * Wt[f][tap][c]
//   * A tile  = 128 output pixels x 32 channels of ONE tap: a 4-D TMA box (32c, tw, th, tn) of the
//     NHWC map whose start coordinate is shifted by the tap; the halo is TMA's out-of-bounds
//     zero fill.  Rows land 128 B apart with the 128-byte swizzle = UMMA's K-major layout.
//   * B tile  = BN filters x 32 contraction elements of the [F][kh*kw*C] filter matrix (2-D box).
//   * one CTA per SM, persistent over output tiles; warp 0 = TMA producer, warp 1 = MMA issuer,
//     warps 2..9 = epilogue (TMEM -> registers -> bias/ReLU -> swizzled staging tile -> TMA store,
//     optionally the BatchNorm column sums of the tile).  smem ring of STAGES k-blocks (full/empty
//     mbarriers), two TMEM accumulator buffers so the epilogue of tile i overlaps the MMAs of
//     tile i+1.  The same kernel runs bf16 operand copies (kind::f16, 64 elements per 128-byte row).
//   dgrad (stride 1) is the same kernel on dY with the filters flipped and transposed
//   (flip_transpose_filters_kernel), padding k-1-p.
//
// wgrad (conv_wgrad_halo.cu is the kernel the layers use; the "MN-major" conv_wgrad_kernel below is
// the fallback for shapes it does not take):
//     dW[f][tap][c] += sum_{p} dY[p][f] * X[p shifted by tap][c]
//   the contraction runs over pixels, the slow index of both operands, so both are MN-major
//   UMMA operands: A = 4 TMA boxes (32 f x 32 px) of dY, B = C/32 boxes (32 c, tw, th, tn) of X,
//   in the 128B-swizzle-with-32B-atoms layout that fp32 MN-major operands require.
//   Split over pixel ranges across CTAs (the output is tiny), fp32 red.global.add epilogue.
#include "common.cuh"
#include "tc_common.cuh"
#include "bn_common.cuh"
#include <cuda_bf16.h>
#include <mutex>
#include <cstdlib>

using namespace sn;
using namespace sn::tc;

namespace sn {

long long g_igemm_pair_launches = 0;
int g_igemm_pair_mode = -1;          // -1: read SN_IGEMM_PAIR on first use; sn_igemm_set_pair_mode() overrides

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

int make_tmap_f32(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
                  const int *box, bool atom32b) {
    return make_tmap(out, base, rank, dims, strides_elems, box, atom32b, false);
}

// fp32, no swizzle: rows land in shared memory exactly as the box enumerates them
int make_tmap_linear_f32(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
                         const int *box) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return SN_ERR_UNSUPPORTED; }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = (cuuint64_t)dims[i]; bdim[i] = (cuuint32_t)box[i]; estr[i] = 1; }
    for (int i = 1; i < rank; ++i) gstr[i - 1] = (cuuint64_t)strides_elems[i] * 4;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (linear) failed (%d): rank %d dims %lld %lld %lld box %d %d %d", (int)r, rank, dims[0],
                  rank > 1 ? dims[1] : 0, rank > 2 ? dims[2] : 0, box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
        return SN_ERR_UNSUPPORTED;
    }
    return SN_OK;
}

int make_tmap_bf16_sw64(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
                        const int *box) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return SN_ERR_UNSUPPORTED; }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = (cuuint64_t)dims[i]; bdim[i] = (cuuint32_t)box[i]; estr[i] = 1; }
    for (int i = 1; i < rank; ++i) gstr[i - 1] = (cuuint64_t)strides_elems[i] * 2;
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (bf16, 64B swizzle) failed (%d): rank %d dims %lld %lld box %d %d", (int)r, rank, dims[0],
                  rank > 1 ? dims[1] : 0, box[0], rank > 1 ? box[1] : 0);
        return SN_ERR_UNSUPPORTED;
    }
    return SN_OK;
}

int make_tmap(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
              const int *box, bool atom32b, bool bf16) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return SN_ERR_UNSUPPORTED; }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = (cuuint64_t)dims[i]; bdim[i] = (cuuint32_t)box[i]; estr[i] = 1; }
    for (int i = 1; i < rank; ++i) gstr[i - 1] = (cuuint64_t)strides_elems[i] * (bf16 ? 2 : 4);
    CUresult r = enc(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %lld %lld %lld %lld box %d %d %d %d", (int)r, rank, dims[0],
                  rank > 1 ? dims[1] : 0, rank > 2 ? dims[2] : 0, rank > 3 ? dims[3] : 0, box[0], rank > 1 ? box[1] : 0,
                  rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return SN_ERR_UNSUPPORTED;
    }
    return SN_OK;
}

}  // namespace sn

namespace {

constexpr int BM = 128;          // output pixels per tile (UMMA M)
constexpr int BK = 32;           // fp32 contraction elements per k-block = one 128-byte swizzle row
constexpr int UK = 8;            // UMMA K for kind::tf32
constexpr int A_BYTES = BM * BK * 4;     // = 128 rows x 128 B for bf16 operands too (64 elements per row)
constexpr int NUM_THREADS = 192;

struct IgemmParams {
    float *out;
    const float *bias;
    int rows;              // total output pixels (N * OH * OW)
    int F;                 // output channels (GEMM N)
    int taps_w;            // kw
    int ntaps;             // kh * kw
    int C;                 // channels per tap (GEMM K per tap)
    int cblocks;           // ceil(C / 32)
    int tiles_m, tiles_n;
    int th, tn;            // box: th rows of tn images (tw = OW always)
    int bands_per_image;   // OH / th when an image is split into row bands, else 0
    int pad_h, pad_w;
    int act, accumulate;
    int tma_store;         // epilogue stages 32x32 sub-tiles in smem and writes them with TMA stores
    double *stats;         // optional: per-channel sums of the output and of its square (BatchNorm statistics), see below
};

// BatchNorm statistics in the epilogue: the BatchNormalization that consumes a convolution needs
// sum(y) and sum(y^2) per channel, which would otherwise cost a full re-read of y.  Each epilogue
// warp column-sums its 32x32 staging tile out of shared memory (lane = column; the swizzle makes
// the row-wise reads conflict-free), accumulates into its own fp32 table, and the CTA adds the four
// tables to stats[slot][{sum, sumsq}][F] (float64, slot = blockIdx.x % BN_SLOTS) once at the end.
constexpr int STAT_MAX_F = 512;
constexpr int STAT_BYTES = 4 * 2 * STAT_MAX_F * 4;          // [epilogue warp][{sum, sumsq}][STAT_MAX_F]
// the bias vector staged in shared memory once per CTA: fetched with __ldg per 32-column chunk of every tile it was eight
// L2 round trips (long-scoreboard stalls on every FADD in ncu) in a chain the kernel turned out to be bound by
constexpr int BIAS_MAX_F = 1024;
constexpr int BIAS_BYTES = BIAS_MAX_F * 4;

// Epilogue: EPI_WARPS warps, two per TMEM lane quarter (a warp may only read the quarter
// warp_id % 4), alternating over the 32-column chunks of the tile.  One epilogue warp per scheduler
// is a ~300-instruction dependent chain per chunk (~2300 cycles): with four warps the epilogue of a
// 128x128 tile (9150 cycles) took twice the tile's operand ingest, and the profile showed the
// epilogue warps never waiting.
constexpr int EPI_WARPS = 8;
constexpr int IG_THREADS = 64 + 32 * EPI_WARPS;         // warp 0: TMA producer, warp 1: MMA issuer, warps 2..: epilogue
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 4096;       // one 32-row x 128-byte staging tile per epilogue warp

// KS = k-blocks (128-byte K slabs) per pipeline stage.  A stage must carry enough MMA time
// (KS * 128*BN*32 / 2048 cycles) to amortise the per-stage barrier round trips of the single
// producer and issuer threads; narrow tiles therefore use KS = 2.
// PAIR: two CTAs of a cluster share every tile pair (cta_group::2, M = 256): each CTA keeps its own
// 128 pixels of A and only HALF of the filter tile, which cuts the operand bytes an SM has to take
// in per MMA cycle from (128 + BN) to (128 + BN/2) rows — the single-CTA tile is bound by that ingest.
template <int BN, int KS, bool PAIR = false> struct IgemmCfg {
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 4;
    static constexpr int SLAB_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGE_BYTES = KS * SLAB_BYTES;
    static constexpr int RING_BUDGET = 232448 - (EPI_STAGE_BYTES + STAT_BYTES + BIAS_BYTES + 1024 /*align slack*/ + 256 /*barriers*/);
    static constexpr int STAGES = RING_BUDGET / STAGE_BYTES > 8 ? 8 : RING_BUDGET / STAGE_BYTES;
    static_assert(STAGES >= 3, "pipeline too shallow: use KS = 1 for this tile width");
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;      // power of two >= 32
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + STAT_BYTES + BIAS_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// BF16: the operands are bf16 copies (64 elements per 128-byte row, UMMA K = 16, kind::f16); the
// accumulators, bias, activation and output stay fp32.  Byte geometry of the tiles is identical.
// OBF16: the OUTPUT map is stored in bf16 as well (the bf16 tier's convolutions in front of a fused
// BatchNormalization + pool: the map is written once and read twice, all three at half the bytes); the
// epilogue then stages 64-byte rows (tc_common.cuh) and the statistics are those of the rounded values.
template <int BN, int KS, bool BF16, bool PAIR, bool OBF16 = false>
__global__ void __launch_bounds__(IG_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmO, const IgemmParams p) {
    using Cfg = IgemmCfg<BN, KS, PAIR>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *epi_stage = smem + Cfg::STAGES * Cfg::STAGE_BYTES;                 // 1024-byte aligned (stage sizes are)
    float *sstat = reinterpret_cast<float *>(epi_stage + EPI_STAGE_BYTES);      // [4][2][STAT_MAX_F]
    float *sbias = reinterpret_cast<float *>(epi_stage + EPI_STAGE_BYTES + STAT_BYTES);              // [BIAS_MAX_F], zero when there is no bias
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(epi_stage + EPI_STAGE_BYTES + STAT_BYTES + BIAS_BYTES);
    uint64_t *empty_bar = full_bar + Cfg::STAGES;
    uint64_t *tmem_full = empty_bar + Cfg::STAGES;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
    const int lane = threadIdx.x & 31;
    constexpr int BKE = BF16 ? 64 : 32;               // contraction elements per 128-byte k-block
    // PAIR: the two CTAs of a cluster walk the same sequence of tile pairs; CTA `rank` owns pixel tile
    // 2*i + rank of pair i (tiles_m is even, host-checked), both share the filter tile
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const int tile_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int num_tiles = (PAIR ? (p.tiles_m >> 1) : p.tiles_m) * p.tiles_n;
    const int kblocks = p.ntaps * p.cblocks;          // a multiple of KS (host-checked)
    const int kstages = kblocks / KS;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        // PAIR: the leader's accumulator-free barrier collects the epilogue warps of both CTAs
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
        fence_barrier_init();
    }
    if constexpr (PAIR) cluster_sync_all();            // the peer's barriers exist before anything signals them
    if (warp == 1) {
        if constexpr (PAIR) { tmem_alloc_pair(tmem_ptr, Cfg::TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    if (p.stats)
        for (int i = threadIdx.x; i < 4 * 2 * STAT_MAX_F; i += IG_THREADS) sstat[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);     // shuffle from lane 0: provably warp-uniform
    pdl_wait();        // everything above (barriers, TMEM, descriptor prefetch) may overlap the previous kernel's tail
    const bool bias_staged = p.bias && p.F <= BIAS_MAX_F;
    if (warp >= 2 && bias_staged) {
        for (int i = threadIdx.x - 64; i < ((p.F + 31) & ~31); i += 32 * EPI_WARPS) sbias[i] = i < p.F ? __ldg(p.bias + i) : 0.f;
        asm volatile("bar.sync 2, %0;" ::"n"(32 * EPI_WARPS) : "memory");       // the epilogue warps only
    }

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the (uniform) control flow and address arithmetic; lane 0 issues.
        // UTMALDG / UTCHMMA take their operands from UNIFORM registers: under a divergent
        // `if (lane == 0)` every operand needed an R2UR move per instruction (~100 cycles per MMA).
        {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
                const int tm = PAIR ? 2 * (tile / p.tiles_n) + rank : tile / p.tiles_n, tn_idx = tile % p.tiles_n;
                int n0, h0;
                if (p.bands_per_image) { n0 = tm / p.bands_per_image; h0 = (tm % p.bands_per_image) * p.th; }
                else { n0 = tm * p.tn; h0 = 0; }
                int tap = 0, cb = 0, u = 0, v = 0;
                for (int ks = 0; ks < kstages; ++ks) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    // PAIR: both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
                    if (PAIR ? (rank == 0 && elect_one()) : elect_one())
                        mbar_expect_tx(&full_bar[stage], PAIR ? 2 * Cfg::STAGE_BYTES : Cfg::STAGE_BYTES);
                    const uint32_t full_leader = PAIR ? map_to_rank(smem_u32(&full_bar[stage]), 0) : 0u;
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        uint8_t *sa = smem + stage * Cfg::STAGE_BYTES + s * Cfg::SLAB_BYTES;
                        if (elect_one()) {
                            if constexpr (PAIR) {
                                tma_load_4d_pair(sa, &tmA, full_leader, cb * BKE, v - p.pad_w, h0 + u - p.pad_h, n0);
                                tma_load_2d_pair(sa + A_BYTES, &tmB, full_leader, tap * p.C + cb * BKE, tn_idx * BN + rank * (BN / 2));
                            } else {
                                tma_load_4d(sa, &tmA, &full_bar[stage], cb * BKE, v - p.pad_w, h0 + u - p.pad_h, n0);
                                tma_load_2d(sa + A_BYTES, &tmB, &full_bar[stage], tap * p.C + cb * BKE, tn_idx * BN);
                            }
                        }
                        if (++cb == p.cblocks) { cb = 0; ++tap; if (++v == p.taps_w) { v = 0; ++u; } }
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform control flow, lane 0 issues) =====================
        // PAIR: only the leader CTA issues; one instruction drives the tensor cores of both SMs
        if (!PAIR || rank == 0) {
            constexpr uint32_t idesc = make_idesc(BF16 ? 1 : 2, PAIR ? 2 * BM : BM, BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int iter = 0;
            for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++iter) {
                const int acc = iter & 1;
                if constexpr (PAIR) mbar_wait_cluster(&tmem_empty[acc], ((iter >> 1) & 1) ^ 1);
                else mbar_wait(&tmem_empty[acc], ((iter >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int ks = 0; ks < kstages; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
#pragma unroll
                    for (int s = 0; s < KS; ++s) {
                        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES + s * Cfg::SLAB_BYTES);
                        const uint64_t adesc = make_smem_desc(sa, 0, 1024);
                        const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 0, 1024);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / UK; ++k) {
                                // advancing K inside the 128-byte swizzle row: +32 bytes = +2 descriptor units
                                const uint32_t accum = (ks | s | k) != 0 ? 1u : 0u;
                                if constexpr (PAIR) {
                                    if constexpr (BF16) mma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                                    else mma_tf32_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                                } else {
                                    if constexpr (BF16) mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                                    else mma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                                }
                            }
                        }
                    }
                    // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                    if (elect_one()) { if constexpr (PAIR) mma_commit_pair(&empty_bar[stage]); else mma_commit(&empty_bar[stage]); }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue (of both CTAs)
                if (elect_one()) { if constexpr (PAIR) mma_commit_pair(&tmem_full[acc]); else mma_commit(&tmem_full[acc]); }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;        // which of the quarter's EPI_WARPS/4 warps: chunks half, half + EPI_WARPS/4, ...
        int iter = 0;
        for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++iter) {
            const int tm = PAIR ? 2 * (tile / p.tiles_n) + rank : tile / p.tiles_n, tn_idx = tile % p.tiles_n;
            const int acc = iter & 1;
            mbar_wait(&tmem_full[acc], (iter >> 1) & 1);
            tc_fence_after();
            const int row = tm * BM + q * 32 + lane;
            const bool row_ok = row < p.rows;
            float *orow = p.out + (long long)row * p.F;
            constexpr int CHUNK = (BN >= 32) ? 32 : 16;
#pragma unroll 1
            for (int c0 = half * CHUNK; c0 < BN; c0 += (EPI_WARPS / 4) * CHUNK) {
                uint32_t r[CHUNK];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
                if constexpr (CHUNK == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
                tmem_ld_wait();
                const int f0 = tn_idx * BN + c0;
                if constexpr (CHUNK == 32) {
                    if (p.tma_store) {
                        // bias / ReLU in registers, then this warp's 32 rows x 32 columns go to its staging
                        // tile in the 128-byte-swizzled layout of the output tensor map, and ONE TMA store
                        // writes them as full 128-byte lines (rows / columns past the tensor are clipped)
                        uint8_t *st = epi_stage + (warp - 2) * 4096;
                        if (lane == 0) tma_store_wait_read();           // previous store has drained the tile
                        __syncwarp();
                        if constexpr (OBF16) {
                            float v[32];
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (bias_staged) { if (f0 + j < p.F) b = *reinterpret_cast<const float4 *>(sbias + f0 + j); }
                                else if (p.bias && f0 + j < p.F) b = __ldg(reinterpret_cast<const float4 *>(p.bias + f0 + j));
                                v[j] = __uint_as_float(r[j]) + b.x; v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
                                v[j + 2] = __uint_as_float(r[j + 2]) + b.z; v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
                                if (p.act == SN_ACT_RELU) {
                                    v[j] = relu_keep_sign(v[j]); v[j + 1] = relu_keep_sign(v[j + 1]);
                                    v[j + 2] = relu_keep_sign(v[j + 2]); v[j + 3] = relu_keep_sign(v[j + 3]);
                                }
                            }
                            stage_row_bf16(st, lane, v);
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) { tma_store_2d(&tmO, st, f0, tm * BM + q * 32); tma_store_commit(); }
                            if (p.stats)
                                staged_tile_colsums_bf16(st, lane, p.rows - (tm * BM + q * 32), sstat + q * 2 * STAT_MAX_F + f0,
                                                         sstat + (q * 2 + 1) * STAT_MAX_F + f0, p.F - f0);
                            continue;
                        }
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                   __uint_as_float(r[j + 3]));
                            if (p.bias && f0 + j < p.F) {
                                const float4 b = bias_staged ? *reinterpret_cast<const float4 *>(sbias + f0 + j)
                                                             : __ldg(reinterpret_cast<const float4 *>(p.bias + f0 + j));
                                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                            }
                            if (p.act == SN_ACT_RELU) {
                                v.x = relu_keep_sign(v.x); v.y = relu_keep_sign(v.y); v.z = relu_keep_sign(v.z); v.w = relu_keep_sign(v.w);
                            }
                            *reinterpret_cast<float4 *>(st + lane * 128 + (((j >> 2) ^ (lane & 7)) * 16)) = v;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) { tma_store_2d(&tmO, st, f0, tm * BM + q * 32); tma_store_commit(); }
                        if (p.stats)
                            staged_tile_colsums(st, lane, p.rows - (tm * BM + q * 32), sstat + q * 2 * STAT_MAX_F + f0,
                                                sstat + (q * 2 + 1) * STAT_MAX_F + f0, p.F - f0);
                        continue;
                    }
                }
                if (row_ok) {
                    if (f0 + CHUNK <= p.F && (p.F & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < CHUNK; j += 4) {
                            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                   __uint_as_float(r[j + 3]));
                            float4 *dst = reinterpret_cast<float4 *>(orow + f0 + j);
                            // accumulate first: a 3xTF32 pass adds its partial product to the running sum and
                            // only the last pass carries the bias and the activation
                            if (p.accumulate) { float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                            if (p.bias) {
                                const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + f0 + j));
                                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                            }
                            if (p.act == SN_ACT_RELU) {
                                v.x = relu_keep_sign(v.x); v.y = relu_keep_sign(v.y); v.z = relu_keep_sign(v.z); v.w = relu_keep_sign(v.w);
                            }
                            *dst = v;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CHUNK; ++j) {
                            const int f = f0 + j;
                            if (f < p.F) {
                                float v = __uint_as_float(r[j]);
                                if (p.accumulate) v += orow[f];
                                if (p.bias) v += __ldg(p.bias + f);
                                if (p.act == SN_ACT_RELU) v = relu_keep_sign(v);
                                orow[f] = v;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(map_to_rank(smem_u32(&tmem_empty[acc]), 0));
                else mbar_arrive(&tmem_empty[acc]);
            }
        }
        if (p.tma_store && lane == 0) tma_store_wait_all();
        __syncwarp();
        if (p.stats) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");           // the epilogue warps only
            double *slot = p.stats + (size_t)(blockIdx.x % BN_SLOTS) * 2 * p.F;
            for (int f = threadIdx.x - 64; f < p.F; f += 32 * EPI_WARPS) {
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int wq = 0; wq < 4; ++wq) { s1 += (double)sstat[wq * 2 * STAT_MAX_F + f]; s2 += (double)sstat[(wq * 2 + 1) * STAT_MAX_F + f]; }
                atomicAdd(slot + f, s1);
                atomicAdd(slot + p.F + f, s2);
            }
        }
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();            // neither CTA leaves while the other may still signal it
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// Wt[c][kh-1-u][kw-1-v][f] = W[f][u][v][c]  (dgrad as a forward convolution of dY)
__global__ void flip_transpose_filters_kernel(const float *__restrict__ w, float *__restrict__ wt, int F, int C, int kh, int kw) {
    const long long total = (long long)F * kh * kw * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // i indexes Wt: ((c * kh + u2) * kw + v2) * F + f
        int f = (int)(i % F);
        long long t = i / F;
        int v2 = (int)(t % kw); t /= kw;
        int u2 = (int)(t % kh);
        int c = (int)(t / kh);
        wt[i] = w[(((long long)f * kh + (kh - 1 - u2)) * kw + (kw - 1 - v2)) * C + c];
    }
}

__global__ void flip_transpose_filters_bf16_kernel(const float *__restrict__ w, __nv_bfloat16 *__restrict__ wt, int F, int C, int kh, int kw) {
    const long long total = (long long)F * kh * kw * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int f = (int)(i % F);
        long long t = i / F;
        int v2 = (int)(t % kw); t /= kw;
        int u2 = (int)(t % kh);
        int c = (int)(t / kh);
        wt[i] = __float2bfloat16_rn(w[(((long long)f * kh + (kh - 1 - u2)) * kw + (kw - 1 - v2)) * C + c]);
    }
}

// One launch per training step for ALL convolutions of the bf16 tier: the bf16 shadow of the whole
// parameter arena (the fprop operands are slices of it) and every dgrad operand
// Wt[c][kh-1-u][kw-1-v][f] = W[f][u][v][c].  A block is one work unit: either 1024 consecutive arena
// elements to cast, or one 32 (f) x 32 (c) tile of one tap of one convolution, transposed through
// shared memory so that both the fp32 reads (c fastest) and the bf16 writes (f fastest) coalesce.
// tasks: 8 int64 per convolution {arena offset of W, address of Wt, F, C, kh, kw, first unit, units}.
__global__ void __launch_bounds__(256) prepare_weights_bf16_kernel(const float *__restrict__ arena, __nv_bfloat16 *__restrict__ shadow,
                                                                   long long n, int cast_units, const long long *__restrict__ tasks,
                                                                   int ntasks) {
    __shared__ float tile[32][33];
    pdl_wait();
    const int u = blockIdx.x;
    if (u < cast_units) {
        const long long i = (long long)u * 1024 + threadIdx.x * 4;
        if (i + 3 < n) {
            const float4 v = *reinterpret_cast<const float4 *>(arena + i);
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 o; o.x = *reinterpret_cast<uint32_t *>(&lo); o.y = *reinterpret_cast<uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(shadow + i) = o;
        } else {
            for (long long j = i; j < n && j < i + 4; ++j) shadow[j] = __float2bfloat16_rn(arena[j]);
        }
        return;
    }
    const int tu = u - cast_units;
    for (int t = 0; t < ntasks; ++t) {
        const long long *tk = tasks + 8 * t;
        const int first = (int)tk[6], units = (int)tk[7];
        if (tu < first || tu >= first + units) continue;
        const int F = (int)tk[2], C = (int)tk[3], kh = (int)tk[4], kw = (int)tk[5];
        const int ftiles = (F + 31) >> 5, ctiles = (C + 31) >> 5;
        int r = tu - first;
        const int ct = r % ctiles; r /= ctiles;
        const int ft = r % ftiles; r /= ftiles;
        const int tap = r;                                   // u*kw + v of W
        const int uu = tap / kw, vv = tap - uu * kw;
        const int tap2 = (kh - 1 - uu) * kw + (kw - 1 - vv); // position in Wt
        const float *w = arena + tk[0];
        __nv_bfloat16 *wt = reinterpret_cast<__nv_bfloat16 *>(tk[1]);
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        for (int rr = ty; rr < 32; rr += 8) {                // rows = f, columns = c (contiguous in W)
            const int f = ft * 32 + rr, c = ct * 32 + tx;
            tile[rr][tx] = (f < F && c < C) ? w[((long long)f * kh * kw + tap) * C + c] : 0.f;
        }
        __syncthreads();
        for (int rr = ty; rr < 32; rr += 8) {                // rows = c, columns = f (contiguous in Wt)
            const int c = ct * 32 + rr, f = ft * 32 + tx;
            if (c < C && f < F) wt[((long long)c * kh * kw + tap2) * F + f] = __float2bfloat16_rn(tile[tx][rr]);
        }
        return;
    }
}

// lo part of the 3xTF32 split: kind::tf32 truncates its operands to 10 mantissa bits, so
// hi = x & 0xFFFFE000 is what the tensor core sees of x itself, and x - hi is exact in fp32
__global__ void split_tf32_lo_kernel(const float *__restrict__ s, float *__restrict__ lo, long long n) {
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // lo = x - hi is exact; it is then ROUNDED to 10 mantissa bits here (nearest, ties away), so that the
    // tensor core's own truncation of it is exact and the split carries no systematic bias of its own
    auto part = [](float x) {
        const float lo = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        return __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
    };
    if (((((uintptr_t)s) | ((uintptr_t)lo)) & 15) == 0) {
        for (long long i = t; i < n4; i += stride) {
            const float4 v = ldg4(s + 4 * i);
            *reinterpret_cast<float4 *>(lo + 4 * i) = make_float4(part(v.x), part(v.y), part(v.z), part(v.w));
        }
        for (long long i = 4 * n4 + t; i < n; i += stride) lo[i] = part(s[i]);
    } else {
        for (long long i = t; i < n; i += stride) lo[i] = part(s[i]);
    }
}

__global__ void cast_f32_bf16_kernel(const float *__restrict__ s, __nv_bfloat16 *__restrict__ d, long long n) {
    pdl_wait();
    const long long n4 = n >> 2, stride = (long long)gridDim.x * blockDim.x, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if ((((uintptr_t)s) & 15) == 0 && (((uintptr_t)d) & 7) == 0) {
        for (long long i = t; i < n4; i += stride) {
            const float4 v = ldg4(s + 4 * i);
            __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
            uint2 o; o.x = *reinterpret_cast<uint32_t *>(&a); o.y = *reinterpret_cast<uint32_t *>(&b);
            reinterpret_cast<uint2 *>(d)[i] = o;
        }
        for (long long i = 4 * n4 + t; i < n; i += stride) d[i] = __float2bfloat16_rn(s[i]);
    } else {
        for (long long i = t; i < n; i += stride) d[i] = __float2bfloat16_rn(s[i]);
    }
}

// --------------------------------------------------------------------------------------------
// wgrad: dW[f][tap][c] += sum over a pixel range of dY[p][f] * X[p + tap][c]; both operands MN-major.
constexpr int PB = 32;                   // pixels per k-block (4 UMMA K-steps of 8)
constexpr int ATOM_BYTES = PB * 128;     // one MN-atom column: 32 pixel rows x 128 B
constexpr int WG_A_BYTES = 4 * ATOM_BYTES;   // 128 filters

struct WgradParams {
    float *dw;
    int F, C, ntaps, taps_w;
    int c_tiles, f_tiles;      // work items = f_tiles * c_tiles * ntaps
    int splits, pblocks_per_split, pblocks;
    int th, tn, S;             // pixel-block box: th rows of tn images; S = OH*OW
    int OW;
    int pad_h, pad_w;
    int tma_reduce;            // epilogue: TMA reduce-add boxes (1) or per-element atomics (0)
};

// dW as (C, taps, F), box (32 channels, 1 tap, 32 filters), 128-byte swizzle: channels / filters past the end are clipped
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// TAPS consecutive filter taps share one dY tile per stage (their X tiles differ only by the tap
// shift): 1/TAPS of the dY traffic from L2 and TAPS x the MMA time per barrier round trip.
constexpr int pow2_at_least(int v) { int r = 32; while (r < v) r *= 2; return r; }
template <int CN, int TAPS> struct WgradCfg {
    static constexpr int B_BYTES = (CN / 32) * ATOM_BYTES;           // per tap
    static constexpr int STAGE_BYTES = WG_A_BYTES + TAPS * B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES > 8 ? 8 : (192 * 1024) / STAGE_BYTES;
    static constexpr int TMEM_COLS = pow2_at_least(TAPS * CN);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static_assert(TAPS * CN <= 512, "accumulators exceed TMEM");
};

template <int CN, int TAPS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmDW, const WgradParams p) {
    using Cfg = WgradCfg<CN, TAPS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t *empty_bar = full_bar + Cfg::STAGES;
    uint64_t *acc_full = empty_bar + Cfg::STAGES;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_full + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform role index
    const int lane = threadIdx.x & 31;

    // work item of this CTA
    const int split = blockIdx.x % p.splits;
    int item = blockIdx.x / p.splits;
    const int tap_groups = p.ntaps / TAPS;
    const int tap0 = (item % tap_groups) * TAPS; item /= tap_groups;
    const int ct = item % p.c_tiles;
    const int ft = item / p.c_tiles;
    const int pb0 = split * p.pblocks_per_split;
    const int pb1 = min(p.pblocks, pb0 + p.pblocks_per_split);
    const int nkb = pb1 - pb0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int i = 0; i < Cfg::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);     // shuffle from lane 0: provably warp-uniform

    if (nkb > 0) {
        if (warp == 0) {
            {   // warp-uniform control flow; the elected lane issues (see conv_igemm_kernel)
                int stage = 0; uint32_t phase = 0;
                for (int pb = pb0; pb < pb1; ++pb) {
                    const int pix = pb * PB;
                    int n0, i0;
                    if (p.S >= PB) { n0 = pix / p.S; i0 = (pix % p.S) / p.OW; }
                    else { n0 = pix / p.S; i0 = 0; }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * Cfg::STAGE_BYTES;
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
                        for (int a = 0; a < 4; ++a)
                            tma_load_2d(sa + a * ATOM_BYTES, &tmDY, &full_bar[stage], ft * 128 + a * 32, pix);
#pragma unroll
                        for (int t = 0; t < TAPS; ++t) {
                            const int u = (tap0 + t) / p.taps_w, v = (tap0 + t) % p.taps_w;
#pragma unroll
                            for (int b = 0; b < CN / 32; ++b)
                                tma_load_4d(sa + WG_A_BYTES + t * Cfg::B_BYTES + b * ATOM_BYTES, &tmX, &full_bar[stage],
                                            ct * CN + b * 32, v - p.pad_w, i0 + u - p.pad_h, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            {
                constexpr uint32_t idesc = make_idesc(2, 128, CN, 1, 1);      // both operands MN-major
                int stage = 0; uint32_t phase = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < PB / UK; ++k) {
                        // UMMA K-step k = pixel rows 8k..8k+7 of every atom column (+1024 B) = two 4-row
                        // K groups 512 B apart (SBO); atom columns of 32 f / 32 c are ATOM_BYTES apart (LBO)
                        const uint64_t adesc = make_smem_desc(sa + k * 1024, ATOM_BYTES, 512, LAYOUT_SW128_BASE32B);
#pragma unroll
                        for (int t = 0; t < TAPS; ++t) {
                            const uint64_t bdesc = make_smem_desc(sa + WG_A_BYTES + t * Cfg::B_BYTES + k * 1024, ATOM_BYTES, 512,
                                                                  LAYOUT_SW128_BASE32B);
                            if (elect_one()) mma_tf32(tmem_base + t * CN, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    if (elect_one()) mma_commit(&empty_bar[stage]);
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) mma_commit(acc_full);
                __syncwarp();
            }
        } else {
            const int q = warp & 3;
            mbar_wait_sleep(acc_full, 0);
            tc_fence_after();
            const int f = ft * 128 + q * 32 + lane;
            constexpr int CHUNK = (CN >= 32) ? 32 : 16;
            if (CHUNK == 32 && p.tma_reduce) {
                // the operand ring is idle (every MMA has completed): two 32 x 128 B staging tiles per epilogue warp;
                // one reduce-add box per (tap, 32 channels) instead of 1024 atomics (see conv_wgrad_halo.cu)
                uint8_t *mine = smem + q * 2 * 4096;
                int nb = 0;
#pragma unroll 1
                for (int t = 0; t < TAPS; ++t) {
#pragma unroll 1
                    for (int c0 = 0; c0 < CN; c0 += 32, ++nb) {
                        if (ct * CN + c0 >= p.C) break;
                        uint8_t *st = mine + (nb & 1) * 4096;
                        if (nb >= 2) { if (lane == 0) tma_store_wait_read1(); __syncwarp(); }
                        uint32_t r[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * CN + c0), r);
                        tmem_ld_wait();
                        uint8_t *dst = st + lane * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4 *>(dst + ((j ^ (lane & 7)) * 16)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_reduce_add_3d(&tmDW, st, ct * CN + c0, tap0 + t, ft * 128 + q * 32);
                            tma_store_commit();
                        }
                    }
                }
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            } else
#pragma unroll 1
            for (int t = 0; t < TAPS; ++t) {
                float *row = p.dw + ((long long)f * p.ntaps + tap0 + t) * p.C + ct * CN;
#pragma unroll 1
                for (int c0 = 0; c0 < CN; c0 += CHUNK) {
                    uint32_t r[CHUNK];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * CN + c0);
                    if constexpr (CHUNK == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
                    tmem_ld_wait();
                    if (f < p.F) {
#pragma unroll
                        for (int j = 0; j < CHUNK; ++j)
                            if (ct * CN + c0 + j < p.C) atomicAdd(row + c0 + j, __uint_as_float(r[j]));
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

template <int CN, int TAPS>
int launch_wgrad(const CUtensorMap &tmDY, const CUtensorMap &tmX, const CUtensorMap &tmDW, const WgradParams &p, cudaStream_t st) {
    using Cfg = WgradCfg<CN, TAPS>;
    static bool configured_dev[16] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        SN_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<CN, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    int grid = p.f_tiles * p.c_tiles * (p.ntaps / TAPS) * p.splits;
    conv_wgrad_kernel<CN, TAPS><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, st>>>(tmDY, tmX, tmDW, p);
    return after_launch("conv_wgrad_kernel");
}

// --------------------------------------------------------------------------------------------
// Geometry of the 128-pixel tile: full-width boxes (tw = OW), th rows, tn images.
struct TileGeom { int th, tn, bands_per_image; };

bool tile_geometry(int N, int OH, int OW, TileGeom &g) {
    if (OW <= 0 || OW > BM || (BM % OW) != 0) return false;
    const int S = OH * OW;
    if (S >= BM) {
        g.th = BM / OW;
        if (OH % g.th) return false;
        g.tn = 1;
        g.bands_per_image = OH / g.th;
    } else {
        if (BM % S) return false;
        g.th = OH;
        g.tn = BM / S;
        g.bands_per_image = 0;
        if (g.tn > 256) return false;
    }
    return g.th <= 256;
}

template <int BN, int KS, bool BF16, bool OBF16 = false>
int launch_igemm_ks(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmO, const IgemmParams &p, cudaStream_t st) {
    using Cfg = IgemmCfg<BN, KS>;
    static bool configured_dev[16] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        SN_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, KS, BF16, false, OBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    int tiles = p.tiles_m * p.tiles_n;
    int grid = tiles < num_sms() ? tiles : num_sms();
    SN_CUDA(launch_dependent(conv_igemm_kernel<BN, KS, BF16, false, OBF16>, dim3(grid), dim3(IG_THREADS), Cfg::SMEM_BYTES, st, tmA, tmB, tmO, p));
    return after_launch("conv_igemm_kernel");
}
// CTA pairs: clusters of two, one pair per TPC, persistent over tile pairs
template <int BN, bool BF16, bool OBF16 = false>
int launch_igemm_pair(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmO, const IgemmParams &p, cudaStream_t st) {
    using Cfg = IgemmCfg<BN, 1, true>;
    static bool configured_dev[16] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        SN_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, 1, BF16, true, OBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int pairs = (p.tiles_m / 2) * p.tiles_n, max_pairs = num_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (pairs < max_pairs ? pairs : max_pairs));
    cfg.blockDim = dim3(IG_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SN_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BN, 1, BF16, true, OBF16>, tmA, tmB, tmO, p));
    ++sn::g_igemm_pair_launches;
    return after_launch("conv_igemm_kernel (pair)");
}
template <int BN>
int launch_igemm(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmO, const IgemmParams &p, bool bf16, bool pair,
                 bool out_bf16, cudaStream_t st) {
    if constexpr (BN >= 64) {
        if (out_bf16) {          // bf16 operands AND a bf16 output map (host-checked)
            if constexpr (BN >= 128) {
                if (pair) return launch_igemm_pair<BN, true, true>(tmA, tmB, tmO, p, st);
            }
            if constexpr (BN == 64) {
                if (((p.ntaps * p.cblocks) % 2) == 0) return launch_igemm_ks<BN, 2, true, true>(tmA, tmB, tmO, p, st);
            }
            return launch_igemm_ks<BN, 1, true, true>(tmA, tmB, tmO, p, st);
        }
    }
    if constexpr (BN >= 128) {
        if (pair) return bf16 ? launch_igemm_pair<BN, true>(tmA, tmB, tmO, p, st) : launch_igemm_pair<BN, false>(tmA, tmB, tmO, p, st);
    }
    const int kblocks = p.ntaps * p.cblocks;
    if constexpr (BN <= 64) {             // (a 2-slab stage of a 128-wide tile would leave only 2 ring stages)
        if ((kblocks % 2) == 0)
            return bf16 ? launch_igemm_ks<BN, 2, true>(tmA, tmB, tmO, p, st) : launch_igemm_ks<BN, 2, false>(tmA, tmB, tmO, p, st);
    }
    return bf16 ? launch_igemm_ks<BN, 1, true>(tmA, tmB, tmO, p, st) : launch_igemm_ks<BN, 1, false>(tmA, tmB, tmO, p, st);
}

}  // namespace

namespace sn {

bool conv_rows_eligible(int N, int C, int IH, int IW, int F, int kh, int kw, int OH, int OW, int pad_h, int pad_w, bool out_bf16);
int conv_rows_igemm_tc(const void *x, const void *w, const float *bias, void *out, int N, int C, int IH, int IW, int F, int kh, int kw,
                       int OH, int OW, int pad_h, int pad_w, int act, double *stats, bool out_bf16, cudaStream_t st);

// When do CTA pairs pay?  Measured (profiles/r01_igemm_pair_vs_single.txt): 256-wide tiles with at
// least four waves of tile pairs gain 7-8 % (tf32 reaches 725 TF/s = 89 % of the TF32 roof at
// batch 2048); with fewer tiles the longer barrier round trips across the pair cost more than the
// halved filter traffic saves, and 128-wide tiles never gained.
// SN_IGEMM_PAIR=0 never pairs, =2 pairs every eligible shape (tests, A/B measurements).
static bool igemm_use_pairs(int BN, int tiles_m, int tiles_n) {
    int &mode = g_igemm_pair_mode;
    if (mode < 0) { const char *e = getenv("SN_IGEMM_PAIR"); mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }
    if (mode == 0 || BN < 128 || (tiles_m % 2) != 0) return false;
    if (mode == 2) return true;
    return BN == 256 && (tiles_m / 2) * tiles_n >= 4 * (num_sms() / 2);
}

// Can the tensor-core path take this forward-style convolution (stride 1)?
bool conv_tc_eligible(int N, int C, int OH, int OW, int F, int stride, bool bf16) {
    TileGeom g;
    if (stride != 1 || (C & (bf16 ? 7 : 3)) || C < 8) return false;
    if (!tile_geometry(N, OH, OW, g)) return false;
    long long flops = 2ll * N * OH * OW * F * C;
    return flops >= (1ll << 22);            // tiny problems stay on the CUDA-core kernel (launch-latency bound anyway)
}

// x: NHWC (N, IH, IW, C); w: [F][kh*kw][C]; out: (N*OH*OW, F).
// out_bf16: `out` is a bf16 map (bf16 operands only, F % 8 == 0, F > 32, no accumulation)
int conv_igemm_tc(const void *x, const void *w, const float *bias, void *out_any, int N, int C, int IH, int IW, int F, int kh,
                  int kw, int OH, int OW, int pad_h, int pad_w, int act, int accumulate, bool bf16, double *stats, cudaStream_t st,
                  bool out_bf16) {
    float *out = static_cast<float *>(out_any);
    // filters resident in shared memory + row-shifted activation copies (conv_rows_tc.cu) where the layer allows
    if (bf16 && !accumulate && g_igemm_pair_mode != 2 /* pairs forced: A/B measurements, tests */ && (!bias || (((uintptr_t)bias) & 15) == 0) && ((((uintptr_t)x) | ((uintptr_t)w) | ((uintptr_t)out_any)) & 15) == 0 &&
        conv_rows_eligible(N, C, IH, IW, F, kh, kw, OH, OW, pad_h, pad_w, out_bf16))
        return conv_rows_igemm_tc(x, w, bias, out_any, N, C, IH, IW, F, kh, kw, OH, OW, pad_h, pad_w, act, stats, out_bf16, st);
    if (out_bf16 && !(bf16 && !accumulate && F > 32 && (F & 7) == 0)) {
        set_error("conv_igemm_tc: a bf16 output map needs bf16 operands, F %% 8 == 0, F > 32 and no accumulation (F = %d)", F);
        return SN_ERR_UNSUPPORTED;
    }
    TileGeom g;
    if (!tile_geometry(N, OH, OW, g)) { set_error("conv_igemm_tc: unsupported output geometry %dx%d", OH, OW); return SN_ERR_UNSUPPORTED; }
    if (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) { set_error("conv_igemm_tc: tensors must be 16-byte aligned"); return SN_ERR_INVALID_ARG; }
    int BN = F > 128 ? 256 : (F > 64 ? 128 : (F > 32 ? 64 : (F > 16 ? 32 : 16)));
    // few pixel tiles (deep, small maps): halve the filter tile so that more of the 148 SMs get one
    if (BN == 256 && ((long long)N * OH * OW + BM - 1) / BM * ((F + 255) / 256) < (3 * num_sms()) / 4) BN = 128;
    CUtensorMap tmA, tmB;
    long long adims[4] = {C, IW, IH, N};
    long long astr[4] = {1, C, (long long)IW * C, (long long)IH * IW * C};
    const int bke = bf16 ? 64 : 32;
    int abox[4] = {bke, OW, g.th, g.tn};
    int e = make_tmap(&tmA, x, 4, adims, astr, abox, false, bf16);
    if (e) return e;
    long long ktot = (long long)kh * kw * C;
    long long bdims[2] = {ktot, F};
    long long bstr[2] = {1, ktot};
    // CTA pairs for the wide tiles (each CTA loads half of the filter tile): needs an even number of pixel tiles
    const int tiles_m = (int)(((long long)N * OH * OW + BM - 1) / BM);
    const bool pair = igemm_use_pairs(BN, tiles_m, (F + BN - 1) / BN);
    int bbox[2] = {bke, pair ? BN / 2 : BN};
    e = make_tmap(&tmB, w, 2, bdims, bstr, bbox, false, bf16);
    if (e) return e;
    IgemmParams p{};
    p.out = out; p.bias = bias; p.rows = N * OH * OW; p.F = F; p.taps_w = kw; p.ntaps = kh * kw; p.C = C;
    p.cblocks = (C + bke - 1) / bke;
    p.tiles_m = (p.rows + BM - 1) / BM; p.tiles_n = (F + BN - 1) / BN;
    p.th = g.th; p.tn = g.tn; p.bands_per_image = g.bands_per_image;
    p.pad_h = pad_h; p.pad_w = pad_w; p.act = act; p.accumulate = accumulate;
    // TMA-store epilogue: full-line writes instead of 16-byte pieces of 32 different rows per instruction
    CUtensorMap tmO = tmB;
    p.tma_store = (!accumulate && BN >= 32 && (F & 3) == 0 && (!bias || (((uintptr_t)bias) & 15) == 0)) ? 1 : 0;
    if (out_bf16 && !p.tma_store) { set_error("conv_igemm_tc: bf16 output needs the TMA-store epilogue"); return SN_ERR_UNSUPPORTED; }
    if (p.tma_store) {
        long long odims[2] = {F, p.rows};
        long long ostr[2] = {1, F};
        int obox[2] = {32, 32};
        e = out_bf16 ? make_tmap_bf16_sw64(&tmO, out, 2, odims, ostr, obox) : make_tmap_f32(&tmO, out, 2, odims, ostr, obox);
        if (e) return e;
    }
    if (stats) {
        if (!p.tma_store || F > STAT_MAX_F) { set_error("conv_igemm_tc: no BatchNorm statistics for this shape (F = %d)", F); return SN_ERR_UNSUPPORTED; }
        p.stats = stats;          // zero on entry (caller's contract): no memset node in the step graph
    }
    switch (BN) {
        case 16: return launch_igemm<16>(tmA, tmB, tmO, p, bf16, false, false, st);
        case 32: return launch_igemm<32>(tmA, tmB, tmO, p, bf16, false, false, st);
        case 64: return launch_igemm<64>(tmA, tmB, tmO, p, bf16, false, out_bf16, st);
        case 128: return launch_igemm<128>(tmA, tmB, tmO, p, bf16, pair, out_bf16, st);
        default: return launch_igemm<256>(tmA, tmB, tmO, p, bf16, pair, out_bf16, st);
    }
}

bool wgrad_tc_eligible(int N, int C, int OH, int OW, int F, int stride) {
    // C need not be a multiple of the 32-channel TMA box: the box that straddles the end of the tensor is
    // zero-filled by the TMA and the epilogue skips those columns (Dense 784 -> 500: K = 784 = 24.5 boxes)
    if (stride != 1 || (C & 3) || C < 32 || (F & 3)) return false;
    if (OW <= 0 || OW > PB || (PB % OW)) return false;
    const int S = OH * OW;
    if (S >= PB ? (S % PB) != 0 : (PB % S) != 0) return false;
    long long flops = 2ll * N * OH * OW * F * C;
    return flops >= (1ll << 24);
}

// x: NHWC (N, IH, IW, C); dy: (N*OH*OW, F); dw: [F][kh*kw][C], accumulated into (+=).
int conv_wgrad_tc(const float *x, const float *dy, float *dw, int N, int C, int IH, int IW, int F, int kh, int kw, int OH,
                  int OW, int pad_h, int pad_w, cudaStream_t st) {
    if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) { set_error("conv_wgrad_tc: tensors must be 16-byte aligned"); return SN_ERR_INVALID_ARG; }
    const int S = OH * OW;
    const long long P = (long long)N * S;
    const int TAPS = (kh * kw) % 3 == 0 ? 3 : 1;
    int CN = (C % 256 == 0) ? 256 : ((C % 128 == 0) ? 128 : ((C % 64 == 0) ? 64 : 32));
    if (C % 32) CN = C >= 128 ? 128 : (C >= 64 ? 64 : 32);       // ragged channel count: the last tile hangs over (zero-filled)
    if (TAPS == 3 && CN > 128) CN = 128;                  // 3 accumulators of CN columns must fit the 512 TMEM columns
    WgradParams p{};
    p.dw = dw; p.F = F; p.C = C; p.ntaps = kh * kw; p.taps_w = kw;
    p.c_tiles = (C + CN - 1) / CN; p.f_tiles = (F + 127) / 128;
    p.pblocks = (int)((P + PB - 1) / PB);
    int items = p.f_tiles * p.c_tiles * (p.ntaps / TAPS);
    int splits = num_sms() / items;
    if (splits < 1) splits = 1;
    if (splits > p.pblocks) splits = p.pblocks;
    p.pblocks_per_split = (p.pblocks + splits - 1) / splits;
    p.splits = (p.pblocks + p.pblocks_per_split - 1) / p.pblocks_per_split;
    p.S = S; p.OW = OW;
    if (S >= PB) { p.th = PB / OW; p.tn = 1; } else { p.th = OH; p.tn = PB / S; }
    p.pad_h = pad_h; p.pad_w = pad_w;
    CUtensorMap tmDY, tmX;
    long long ddims[2] = {F, P};
    long long dstr[2] = {1, F};
    int dbox[2] = {32, PB};
    int e = make_tmap_f32(&tmDY, dy, 2, ddims, dstr, dbox, true);
    if (e) return e;
    long long xdims[4] = {C, IW, IH, N};
    long long xstr[4] = {1, C, (long long)IW * C, (long long)IH * IW * C};
    int xbox[4] = {32, OW, p.th, p.tn};
    e = make_tmap_f32(&tmX, x, 4, xdims, xstr, xbox, true);
    if (e) return e;
    static const bool force_redg = [] { const char *v = getenv("SN_WGRAD_REDG"); return v && v[0] == '1'; }();
    p.tma_reduce = force_redg ? 0 : 1;
    CUtensorMap tmDW;
    long long wdims[3] = {C, p.ntaps, F};
    long long wstr[3] = {1, C, (long long)p.ntaps * C};
    int wbox[3] = {32, 1, 32};
    e = make_tmap_f32(&tmDW, dw, 3, wdims, wstr, wbox, false);
    if (e) return e;
    if (TAPS == 3) {
        switch (CN) {
            case 32: return launch_wgrad<32, 3>(tmDY, tmX, tmDW, p, st);
            case 64: return launch_wgrad<64, 3>(tmDY, tmX, tmDW, p, st);
            default: return launch_wgrad<128, 3>(tmDY, tmX, tmDW, p, st);
        }
    }
    switch (CN) {
        case 32: return launch_wgrad<32, 1>(tmDY, tmX, tmDW, p, st);
        case 64: return launch_wgrad<64, 1>(tmDY, tmX, tmDW, p, st);
        case 128: return launch_wgrad<128, 1>(tmDY, tmX, tmDW, p, st);
        default: return launch_wgrad<256, 1>(tmDY, tmX, tmDW, p, st);
    }
}

int flip_transpose_filters_bf16(const float *w, void *wt, int F, int C, int kh, int kw, cudaStream_t st) {
    long long total = (long long)F * kh * kw * C;
    flip_transpose_filters_bf16_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, (__nv_bfloat16 *)wt, F, C, kh, kw);
    return after_launch("flip_transpose_filters_bf16_kernel");
}

int prepare_weights_bf16(const float *arena, void *shadow, long long n, const long long *tasks, int ntasks, int flip_units, cudaStream_t st) {
    const int cast_units = (int)((n + 1023) / 1024);
    if (cast_units + flip_units == 0) return SN_OK;
    SN_CUDA(launch_dependent(prepare_weights_bf16_kernel, dim3(cast_units + flip_units), dim3(256), 0, st, arena, (__nv_bfloat16 *)shadow, n, cast_units, tasks, ntasks));
    return after_launch("prepare_weights_bf16_kernel");
}

int split_tf32_lo(const float *src, float *lo, long long n, cudaStream_t st) {
    if (n <= 0) return SN_OK;
    split_tf32_lo_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, st>>>(src, lo, n);
    return after_launch("split_tf32_lo_kernel");
}

int cast_f32_to_bf16(const float *src, void *dst, long long n, cudaStream_t st) {
    if (n <= 0) return SN_OK;
    SN_CUDA(launch_dependent(cast_f32_bf16_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), 0, st, src, (__nv_bfloat16 *)dst, n));
    return after_launch("cast_f32_bf16_kernel");
}

int flip_transpose_filters(const float *w, float *wt, int F, int C, int kh, int kw, cudaStream_t st) {
    long long total = (long long)F * kh * kw * C;
    flip_transpose_filters_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, wt, F, C, kh, kw);
    return after_launch("flip_transpose_filters_kernel");
}

}  // namespace sn
