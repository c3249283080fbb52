// Tiny-C convolution (C <= 4, kh*kw <= 9: the RGB stem of configs 3 / 4, the README CNN of config 1) with
// the column matrix gathered BY THE TMA UNIT: cp.async.bulk.tensor ... im2col.
//
// conv_stem_tc.cu assembles the column tile with 256 builder threads per CTA (index arithmetic, 16 LDS and
// 4 STS per thread and tile) and is bound by instruction issue: ncu shows 55 % SM throughput at 5 % of DRAM
// bandwidth, 49 us for a map that takes 10 us to write.  Here nothing but the TMA touches the operands:
//
//   * the input is held as NHWC4 (channels padded to 4 floats = 16 bytes per pixel, sn_pad_channels4);
//   * for each filter tap (u, v) ONE im2col TMA load brings [128 output pixels][4 channels] — the halo,
//     the padding and the wrap from one image row / image to the next are the TMA's business (bounding-box
//     corners + out-of-bounds zero fill) — into 2 KB of shared memory, pixel p at byte 16 p;
//   * that IS the canonical no-swizzle K-major UMMA operand: core matrix = 8 pixels x 16 bytes = 128
//     contiguous bytes, 8-pixel groups 128 bytes apart (SBO), the two 16-byte K chunks of one
//     tcgen05.mma kind::tf32 (K = 8) = two taps = 2048 bytes apart (LBO).  Five MMAs cover ten tap slots:
//     nine taps and one slot of zeros (zero weights against finite zeros);
//   * the filters sit in shared memory in the same form, [tap][f][4 channels];
//   * warp 0 issues the TMA loads, warp 1 the MMAs, 24 epilogue warps in three sets take alternate tiles
//     (three TMEM accumulators): TMEM -> bias / ReLU -> staging tile -> TMA store, plus the BatchNorm column
//     sums, with fp32 or bf16 output maps like the other forward kernels.
//
// wgrad stays with conv_stem_tc.cu (builder threads): its contraction runs over PIXELS, so both operands are
// MN-major, and for 32-bit types the tensor core reads MN-major operands only in the 128-byte-swizzle /
// 32-byte-atom layout (128 bytes of k per pixel row) — measured here: a no-swizzle MN-major descriptor over the
// TMA-gathered [pixel][16 B] tap slots yields an all-zero accumulator.  The TMA cannot assemble 128-byte rows
// out of nine 16-byte taps, so that operand has to be built by threads.
#include "common.cuh"
#include "tc_common.cuh"
#include "bn_common.cuh"

using namespace sn;
using namespace sn::tc;

namespace {

constexpr int TAP_SLOTS = 10;                 // 9 taps + 1 constant slot of zeros
constexpr int TAPS_MAX = 9;
constexpr int TAP_BYTES = 128 * 16;           // one tap of one 128-pixel tile
constexpr int A_STAGE_BYTES = TAPS_MAX * TAP_BYTES;      // the constant slot is ONE block shared by all stages (reached through the descriptor's LBO)
constexpr int FP_SETS = 3;                    // epilogue warp sets = TMEM accumulators
constexpr int FP_EPI_WARPS = 8 * FP_SETS;
constexpr int FP_THREADS = 64 + 32 * FP_EPI_WARPS;
// Ring depth: a stage is one whole tile (9 taps = 18 KB, ~1 us of L2 round trip per TMA load), so the depth IS
// the number of tiles in flight; measured with 4 stages the kernel was bound by that latency (0.6 us per tile
// with every other part switched off).  As deep as shared memory allows, at most 8.
template <int BN, bool OBF16> struct FpCfg {
    static constexpr int EPI_TILE = OBF16 ? 2048 : 4096;
    static constexpr int FIXED = TAP_BYTES + TAP_SLOTS * BN * 16 + FP_EPI_WARPS * EPI_TILE + FP_SETS * 4 * 2 * BN * 4 + 1024 + 256;
    static constexpr int FIT = (232448 - FIXED) / A_STAGE_BYTES;
    static constexpr int STAGES = FIT > 8 ? 8 : FIT;
    static_assert(STAGES >= 3, "ring too shallow");
    static constexpr int SMEM = FIXED + STAGES * A_STAGE_BYTES;
};

// Column sums of one staged epilogue tile ACCUMULATED IN REGISTERS over all tiles of a warp (the per-tile
// read-modify-write of shared tables plus four shuffles cost ~700 cycles per tile): lane l keeps {sum, sum of
// squares} of one column pair over its half of the rows; the halves are joined once at the end of the kernel.
__device__ __forceinline__ void colsums_acc_bf16(const uint8_t *st, int lane, int nvalid, float2 &a, float2 &b) {
    const int pr = lane & 15, rsel = lane >> 4;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = 2 * i + rsel;
        const uint32_t w = *reinterpret_cast<const uint32_t *>(st + r * 64 + (((pr >> 2) ^ ((r >> 1) & 3)) * 16) + (pr & 3) * 4);
        float2 v = make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
        if (r >= nvalid) v = make_float2(0.f, 0.f);
        a = __fadd2_rn(a, v);
        b = __ffma2_rn(v, v, b);
    }
}
__device__ __forceinline__ void colsums_acc_f32(const uint8_t *st, int lane, int nvalid, float2 &a, float2 &b) {
    const int cp = lane & 15, r0 = (lane >> 4) * 16;
    const uint8_t *base = st + (cp & 1) * 8;
    const int chunk = cp >> 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = r0 + i;
        float2 v = *reinterpret_cast<const float2 *>(base + r * 128 + ((chunk ^ (r & 7)) * 16));
        if (r >= nvalid) v = make_float2(0.f, 0.f);
        a = __fadd2_rn(a, v);
        b = __ffma2_rn(v, v, b);
    }
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const int *, const int *, cuuint32_t, cuuint32_t, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeIm2colFn get_encode_im2col() {
    static EncodeIm2colFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeIm2colFn)p;
    }
    return fn;
}

// x4: NHWC4 fp32 (N, H, W, 4).  Base positions of the filter window run over [-pad, dim + pad - (k - 1)), i.e.
// one per output pixel; a load of `pixels` consecutive base positions (W fastest, then H, then N) at filter
// offset (v, u) lands as [pixels][4] floats.
int make_tmap_im2col(CUtensorMap *out, const float *x4, int N, int H, int W, int kh, int kw, int ph, int pw, int stride, int pixels) {
    EncodeIm2colFn enc = get_encode_im2col();
    if (!enc) { set_error("cuTensorMapEncodeIm2col is not available from the driver"); return SN_ERR_UNSUPPORTED; }
    cuuint64_t gdim[4] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    int lower[2] = {-pw, -ph};
    int upper[2] = {pw - (kw - 1), ph - (kh - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x4), gdim, gstr, lower, upper, 4, (cuuint32_t)pixels, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col failed (%d): N %d H %d W %d k %dx%d pad %d,%d stride %d", (int)r, N, H, W, kh, kw, ph, pw, stride);
        return SN_ERR_UNSUPPORTED;
    }
    return SN_OK;
}

__device__ __forceinline__ void tma_load_im2col(void *dst, const CUtensorMap *m, uint64_t *bar, int w, int h, int n, int off_w, int off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(0), "r"(w), "r"(h), "r"(n), "h"((uint16_t)off_w), "h"((uint16_t)off_h)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d_rows(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

struct Im2colP {
    const float *w, *bias;
    int N, C, H, W, F, kh, kw, stride, ph, pw, oh, ow;
    int taps;            // kh*kw <= 9
    int P;               // N*oh*ow
    int act, tiles;
    double *stats;
    float inv_ow, inv_oh;
};

__device__ __forceinline__ int fdiv(int x, int d, float inv, int &rem) {      // exact for 0 <= x < 2^24
    int q = __float2int_rz(__int2float_rz(x) * inv);
    rem = x - q * d;
    if (rem < 0) { rem += d; --q; } else if (rem >= d) { rem -= d; ++q; }
    return q;
}

// all taps of the `pixels`-pixel tile that starts at output pixel pix0 -> stage `dst` (tap t at dst + t * tap_bytes)
__device__ __forceinline__ void load_taps(const Im2colP &p, const CUtensorMap *tm, uint8_t *dst, int tap_bytes, uint64_t *bar, int pix0) {
    int j, i;
    const int t = fdiv(pix0, p.ow, p.inv_ow, j), n = fdiv(t, p.oh, p.inv_oh, i);
    const int w0 = j * p.stride - p.pw, h0 = i * p.stride - p.ph;
    int u = 0, v = 0;
    for (int tap = 0; tap < p.taps; ++tap) {
        if (elect_one()) tma_load_im2col(dst + tap * tap_bytes, tm, bar, w0, h0, n, v, u);
        if (++v == p.kw) { v = 0; ++u; }
    }
}

// ============================================================================== fprop
template <int BN, bool OBF16>
__global__ void __launch_bounds__(FP_THREADS, 1) stem_im2col_fprop_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                                                                          const Im2colP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using Cfg = FpCfg<BN, OBF16>;
    constexpr int FP_STAGES = Cfg::STAGES;
    uint8_t *sA = smem;                                              // FP_STAGES x [9 taps][128 px][16 B]
    uint8_t *sZ = sA + FP_STAGES * A_STAGE_BYTES;                    // the constant tap slot (zeros), shared by all stages
    uint8_t *sB = sZ + TAP_BYTES;                                    // [TAP_SLOTS][BN][16 B]
    uint8_t *epi_stage = sB + TAP_SLOTS * BN * 16;                   // per epilogue warp: 32 rows x 128 B (fp32) or x 64 B (bf16)
    constexpr int EPI_TILE = Cfg::EPI_TILE;
    float *sstat = reinterpret_cast<float *>(epi_stage + FP_EPI_WARPS * EPI_TILE);       // [set][quarter][{sum, sumsq}][BN]
    uint64_t *full = reinterpret_cast<uint64_t *>(sstat + FP_SETS * 4 * 2 * BN);
    uint64_t *empty = full + FP_STAGES;
    uint64_t *t_full = empty + FP_STAGES;
    uint64_t *t_empty = t_full + FP_SETS;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(t_empty + FP_SETS);
    constexpr int TMEM_COLS = FP_SETS * BN <= 32 ? 32 : (FP_SETS * BN <= 64 ? 64 : (FP_SETS * BN <= 128 ? 128 : (FP_SETS * BN <= 256 ? 256 : 512)));

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < FP_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < FP_SETS; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmO);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
    // the constant slot, and the tap slots of every stage that are never loaded (filters with fewer than 9 taps): zeros
    for (int i = threadIdx.x; i < TAP_BYTES / 16; i += FP_THREADS) *reinterpret_cast<float4 *>(sZ + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.taps < TAPS_MAX) {
        const int per = (TAPS_MAX - p.taps) * (TAP_BYTES / 16);
        for (int i = threadIdx.x; i < FP_STAGES * per; i += FP_THREADS) {
            const int s = i / per, r = i - s * per;
            *reinterpret_cast<float4 *>(sA + s * A_STAGE_BYTES + p.taps * TAP_BYTES + r * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (int i = threadIdx.x; i < FP_SETS * 4 * 2 * BN; i += FP_THREADS) sstat[i] = 0.f;
    pdl_wait();        // (the weights below are global memory the previous kernels may have written)
    // filters -> [tap][f][4]: w is [F][taps][C]
    for (int i = threadIdx.x; i < TAP_SLOTS * BN; i += FP_THREADS) {
        const int tap = i / BN, f = i - tap * BN;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tap < p.taps && f < p.F) {
            const float *src = p.w + ((long long)f * p.taps + tap) * p.C;
            v.x = __ldg(src);
            if (p.C > 1) v.y = __ldg(src + 1);
            if (p.C > 2) v.z = __ldg(src + 2);
            if (p.C > 3) v.w = __ldg(src + 3);
        }
        *reinterpret_cast<float4 *>(sB + i * 16) = v;
    }
    fence_proxy_async();            // generic-proxy writes (filters, zero slots) before the tensor core reads them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) mbar_expect_tx(&full[stage], (uint32_t)(p.taps * TAP_BYTES));
            __syncwarp();
            load_taps(p, &tmX, sA + stage * A_STAGE_BYTES, TAP_BYTES, &full[stage], tile * 128);
            __syncwarp();
            if (++stage == FP_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = make_idesc(2, 128, BN, 0, 0);
        int stage = 0; uint32_t phase = 0; int iter = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
            const int acc = iter % FP_SETS;
            mbar_wait(&t_empty[acc], ((iter / FP_SETS) & 1) ^ 1);
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sA + stage * A_STAGE_BYTES), b0 = smem_u32(sB);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < TAP_SLOTS / 2; ++k) {
                    // K = 8: the 16-byte chunks of tap slots 2k and 2k+1 (LBO apart); 8-row groups 128 B apart (SBO).
                    // Slot 9 is the shared block of zeros behind the ring: the last descriptor's LBO reaches it.
                    const uint32_t lbo = (k == TAP_SLOTS / 2 - 1) ? (smem_u32(sZ) - (a0 + (TAP_SLOTS - 2) * TAP_BYTES)) : (uint32_t)TAP_BYTES;
                    const uint64_t adesc = make_smem_desc(a0 + 2 * k * TAP_BYTES, lbo, 128, 0);
                    const uint64_t bdesc = make_smem_desc(b0 + 2 * k * BN * 16, BN * 16, 128, 0);
                    mma_tf32(tmem_base + acc * BN, adesc, bdesc, idesc, k ? 1u : 0u);
                }
                mma_commit(&empty[stage]);
                mma_commit(&t_full[acc]);
            }
            __syncwarp();
            if (++stage == FP_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ---------------- epilogue: three sets of eight warps, set s takes tiles s, s + 3, ... ----------------
        const int q = warp & 3;                              // TMEM lane quarter this warp may read
        const int ew = warp - 2;                             // 0..23
        const int set = ew >> 3;
        const int half = (ew >> 2) & 1;                      // which 32-column chunks
        float *mystat = sstat + (set * 4 + q) * 2 * BN;      // shared with the other half's warp only (disjoint columns)
        uint8_t *st = epi_stage + ew * EPI_TILE;
        constexpr int CH = BN >= 32 ? 32 : 16;
        constexpr int NCH = BN >= 64 ? BN / 64 : 1;          // 32-column chunks this warp owns per tile
        float2 cs[NCH], cq[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) { cs[i] = make_float2(0.f, 0.f); cq[i] = make_float2(0.f, 0.f); }
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
            if (iter % FP_SETS != set) continue;
            mbar_wait(&t_full[set], (iter / FP_SETS) & 1);
            tc_fence_after();
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = half * CH + ci * 2 * CH;
                if (c0 >= BN) break;
                uint32_t rr[32];
                if constexpr (CH == 32) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * BN + c0), rr);
                else {
                    uint32_t r16[16];
                    tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * BN + c0), r16);
#pragma unroll
                    for (int j = 0; j < 16; ++j) { rr[j] = r16[j]; rr[j + 16] = 0u; }
                }
                tmem_ld_wait();
                if (c0 >= p.F) continue;
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
                if (lane == 0) tma_store_wait_read();               // this warp's previous store has drained the tile
                __syncwarp();
                if constexpr (OBF16) {
                    stage_row_bf16(st, lane, v);
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; jj += 4)
                        *reinterpret_cast<float4 *>(st + lane * 128 + (((jj >> 2) ^ (lane & 7)) * 16)) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_2d(&tmO, st, c0, tile * 128 + q * 32); tma_store_commit(); }
                if (p.stats) {
                    if constexpr (OBF16) colsums_acc_bf16(st, lane, p.P - (tile * 128 + q * 32), cs[ci], cq[ci]);
                    else colsums_acc_f32(st, lane, p.P - (tile * 128 + q * 32), cs[ci], cq[ci]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[set]);
        }
        if (lane == 0) tma_store_wait_all();
        if (p.stats) {
            // join the two row halves and leave this warp's sums in its table (columns c0 + 2*(lane & 15), +1)
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = half * CH + ci * 2 * CH;
                float2 a = cs[ci], b = cq[ci];
                a.x += __shfl_down_sync(0xffffffffu, a.x, 16); a.y += __shfl_down_sync(0xffffffffu, a.y, 16);
                b.x += __shfl_down_sync(0xffffffffu, b.x, 16); b.y += __shfl_down_sync(0xffffffffu, b.y, 16);
                if (lane < 16 && c0 < BN) {
                    const int c = c0 + 2 * lane;
                    mystat[c] = a.x; mystat[c + 1] = a.y;
                    mystat[BN + c] = b.x; mystat[BN + c + 1] = b.y;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
    if (p.stats) {
        double *slot = p.stats + (size_t)(blockIdx.x % BN_SLOTS) * 2 * p.F;
        for (int f = threadIdx.x; f < p.F; f += FP_THREADS) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int e = 0; e < FP_SETS * 4; ++e) { s1 += (double)sstat[e * 2 * BN + f]; s2 += (double)sstat[(e * 2 + 1) * BN + f]; }
            atomicAdd(slot + f, s1);
            atomicAdd(slot + p.F + f, s2);
        }
    }
}

// ============================================================================== fprop, whole-row tiles
// The im2col-mode kernel above moves 1152 16-byte rows per tile through the TMA unit, and the TMA unit is
// what bounds it (measured: ~1.4 cycles per 16-byte row; loads and stores serialise in it).  When a tile is a
// whole number of image rows (128 % OW == 0, which every CIFAR / MNIST-style stem satisfies) the same operand
// comes from kw plain tiled loads of FAT rows instead: copy v holds image rows i0-ph .. i0-ph+R-1 (R =
// rows per tile + kh - 1) of the input shifted horizontally by v - pw, OW pixels wide — the shift, the halo
// columns and the padding rows are the box start plus out-of-bounds zero fill.  Row pitch = OW * 16 bytes, so
// output pixel m = i * OW + j of tap (u, v) sits at copy_v + u * pitch + 16 m: LINEAR in m, i.e. the canonical
// no-swizzle K-major operand again, and every tap is just a different start address inside the same few KB.
// Per tile: kw loads of R rows of OW*16 bytes (cfg 3: 3 loads, 18 rows of 512 B) instead of 9 x 128 rows of 16 B.
// Taps are taken in address order (v major) and paired; each pair's descriptor carries its own LBO, the odd
// last tap pairs with a block of zeros behind the ring.
// Epilogue: six sets of four warps (six TMEM accumulators); a warp owns ALL BN columns of its 32 pixels, stages
// them as 128-byte rows (64 bf16 or 32 fp32 columns per TMA store: half as many store operations), and keeps
// its BatchNorm column sums in registers until the end of the kernel.
constexpr int RW_SETS = 6;
constexpr int RW_EPI_WARPS = 4 * RW_SETS;
constexpr int RW_THREADS = 64 + 32 * RW_EPI_WARPS;
constexpr int RW_STAGES = 8;

struct RowsP {
    Im2colP g;
    int rows_per_tile, tiles_per_image, R, pitch, copy_bytes, stage_bytes;
    int tap_off[TAP_SLOTS];       // byte offset of tap slot s inside a stage, address order; slots >= taps: -1 (the zero block)
    int tap_w[TAP_SLOTS];         // which tap (u*kw + v) of the filter slot s is, or -1
};

template <int BN, bool OBF16>
__global__ void __launch_bounds__(RW_THREADS, 1) stem_rows_fprop_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                                                                        const RowsP rp) {
    const Im2colP &p = rp.g;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                                              // RW_STAGES x stage_bytes
    uint8_t *sZ = sA + RW_STAGES * rp.stage_bytes;                   // 2 KB of zeros
    uint8_t *sB = sZ + TAP_BYTES;                                    // [TAP_SLOTS][BN][16 B]
    uint8_t *epi_stage = sB + TAP_SLOTS * BN * 16;                   // per epilogue warp: 32 rows x 128 B
    float *sstat = reinterpret_cast<float *>(epi_stage + RW_EPI_WARPS * 4096);           // [set][quarter][{sum, sumsq}][BN]
    uint64_t *full = reinterpret_cast<uint64_t *>(sstat + RW_SETS * 4 * 2 * BN);
    uint64_t *empty = full + RW_STAGES;
    uint64_t *t_full = empty + RW_STAGES;
    uint64_t *t_empty = t_full + RW_SETS;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(t_empty + RW_SETS);
    constexpr int TMEM_COLS = RW_SETS * BN <= 256 ? 256 : 512;
    static_assert(RW_SETS * BN <= 512, "accumulators exceed TMEM");

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < RW_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < RW_SETS; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmO);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
    for (int i = threadIdx.x; i < TAP_BYTES / 16; i += RW_THREADS) *reinterpret_cast<float4 *>(sZ + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < RW_SETS * 4 * 2 * BN; i += RW_THREADS) sstat[i] = 0.f;
    pdl_wait();        // (the weights below are global memory the previous kernels may have written)
    // filters -> [slot][f][4] in the slot (= address) order of the taps: w is [F][taps][C]
    for (int i = threadIdx.x; i < TAP_SLOTS * BN; i += RW_THREADS) {
        const int slot = i / BN, f = i - slot * BN, tap = rp.tap_w[slot];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tap >= 0 && f < p.F) {
            const float *src = p.w + ((long long)f * p.taps + tap) * p.C;
            v.x = __ldg(src);
            if (p.C > 1) v.y = __ldg(src + 1);
            if (p.C > 2) v.z = __ldg(src + 2);
            if (p.C > 3) v.w = __ldg(src + 3);
        }
        *reinterpret_cast<float4 *>(sB + i * 16) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

    if (warp == 0) {
        // ---------------- TMA producer: kw fat-row loads per tile ----------------
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int n = tile / rp.tiles_per_image, i0 = (tile - n * rp.tiles_per_image) * rp.rows_per_tile;
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) mbar_expect_tx(&full[stage], (uint32_t)(p.kw * rp.copy_bytes));
            __syncwarp();
            for (int v = 0; v < p.kw; ++v)
                if (elect_one())
                    tma_load_3d_rows(sA + stage * rp.stage_bytes + v * rp.copy_bytes, &tmX, &full[stage], (v - p.pw) * 4, i0 - p.ph, n);
            __syncwarp();
            if (++stage == RW_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        constexpr uint32_t idesc = make_idesc(2, 128, BN, 0, 0);
        int stage = 0; uint32_t phase = 0; int iter = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
            const int acc = iter % RW_SETS;
            mbar_wait(&t_empty[acc], ((iter / RW_SETS) & 1) ^ 1);
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a0 = smem_u32(sA + stage * rp.stage_bytes), b0 = smem_u32(sB), z0 = smem_u32(sZ);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < TAP_SLOTS / 2; ++k) {
                    // the two 16-byte K chunks of this MMA = tap slots 2k and 2k+1: start at the first, LBO to the second
                    // (a later tap of the same stage, or the block of zeros); 8-pixel groups 128 B apart (SBO)
                    const int o0 = rp.tap_off[2 * k], o1 = rp.tap_off[2 * k + 1];
                    if (o0 < 0) break;
                    const uint32_t s0 = a0 + (uint32_t)o0, s1 = o1 >= 0 ? a0 + (uint32_t)o1 : z0;
                    const uint64_t adesc = make_smem_desc(s0, s1 - s0, 128, 0);
                    const uint64_t bdesc = make_smem_desc(b0 + 2 * k * BN * 16, BN * 16, 128, 0);
                    mma_tf32(tmem_base + acc * BN, adesc, bdesc, idesc, k ? 1u : 0u);
                }
                mma_commit(&empty[stage]);
                mma_commit(&t_full[acc]);
            }
            __syncwarp();
            if (++stage == RW_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ---------------- epilogue: six sets of four warps, set s takes tiles s, s + 6, ... ----------------
        const int q = warp & 3;                              // TMEM lane quarter this warp may read
        const int ew = warp - 2;                             // 0..23
        const int set = ew >> 2;
        float *mystat = sstat + (set * 4 + q) * 2 * BN;
        uint8_t *st = epi_stage + ew * 4096;
        constexpr int CW = OBF16 ? 64 : 32;                  // columns per 128-byte staging row = per TMA store
        constexpr int NCH = (BN + CW - 1) / CW;
        float2 cs[NCH], cq[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) { cs[i] = make_float2(0.f, 0.f); cq[i] = make_float2(0.f, 0.f); }
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
            if (iter % RW_SETS != set) continue;
            mbar_wait(&t_full[set], (iter / RW_SETS) & 1);
            tc_fence_after();
            const int nvalid = p.P - (tile * 128 + q * 32);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = ci * CW;
                if (c0 >= p.F) break;
                if (lane == 0) tma_store_wait_read();               // this warp's previous store has drained the tile
                __syncwarp();
#pragma unroll
                for (int hh = 0; hh < CW / 32; ++hh) {
                    uint32_t rr[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * BN + c0 + 32 * hh), rr);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int jj = 0; jj < 32; jj += 4) {
                        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias && c0 + 32 * hh + jj < p.F) b = __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + 32 * hh + jj));
                        v[jj] = __uint_as_float(rr[jj]) + b.x; v[jj + 1] = __uint_as_float(rr[jj + 1]) + b.y;
                        v[jj + 2] = __uint_as_float(rr[jj + 2]) + b.z; v[jj + 3] = __uint_as_float(rr[jj + 3]) + b.w;
                        if (p.act == SN_ACT_RELU) {
                            v[jj] = relu_keep_sign(v[jj]); v[jj + 1] = relu_keep_sign(v[jj + 1]);
                            v[jj + 2] = relu_keep_sign(v[jj + 2]); v[jj + 3] = relu_keep_sign(v[jj + 3]);
                        }
                    }
                    if constexpr (OBF16) {
                        // 128-byte rows of 64 bf16 in the 128-byte swizzle: 16-byte chunk j of row r at (j ^ (r & 7)) * 16
#pragma unroll
                        for (int jc = 0; jc < 4; ++jc) {
                            uint4 u;
                            u.x = pack_bf16x2(v[8 * jc + 0], v[8 * jc + 1]); u.y = pack_bf16x2(v[8 * jc + 2], v[8 * jc + 3]);
                            u.z = pack_bf16x2(v[8 * jc + 4], v[8 * jc + 5]); u.w = pack_bf16x2(v[8 * jc + 6], v[8 * jc + 7]);
                            *reinterpret_cast<uint4 *>(st + lane * 128 + (((4 * hh + jc) ^ (lane & 7)) * 16)) = u;
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 4)
                            *reinterpret_cast<float4 *>(st + lane * 128 + (((jj >> 2) ^ (lane & 7)) * 16)) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_2d(&tmO, st, c0, tile * 128 + q * 32); tma_store_commit(); }
                if (p.stats) {
                    if constexpr (OBF16) {
                        // lane l owns the column pair (2l, 2l+1) of this 64-column chunk over all 32 rows: one conflict-free
                        // 4-byte load per row (a row is 128 contiguous bytes up to the swizzle of its 16-byte chunks)
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const uint32_t w = *reinterpret_cast<const uint32_t *>(st + r * 128 + (((lane >> 2) ^ (r & 7)) * 16) + (lane & 3) * 4);
                            float2 x2 = make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
                            if (r >= nvalid) x2 = make_float2(0.f, 0.f);
                            cs[ci] = __fadd2_rn(cs[ci], x2);
                            cq[ci] = __ffma2_rn(x2, x2, cq[ci]);
                        }
                    } else {
                        colsums_acc_f32(st, lane, nvalid, cs[ci], cq[ci]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[set]);
        }
        if (lane == 0) tma_store_wait_all();
        if (p.stats) {
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = ci * CW;
                float2 a = cs[ci], b = cq[ci];
                if constexpr (OBF16) {
                    const int c = c0 + 2 * lane;
                    if (c < BN) { mystat[c] = a.x; mystat[c + 1] = a.y; mystat[BN + c] = b.x; mystat[BN + c + 1] = b.y; }
                } else {
                    a.x += __shfl_down_sync(0xffffffffu, a.x, 16); a.y += __shfl_down_sync(0xffffffffu, a.y, 16);
                    b.x += __shfl_down_sync(0xffffffffu, b.x, 16); b.y += __shfl_down_sync(0xffffffffu, b.y, 16);
                    const int c = c0 + 2 * lane;
                    if (lane < 16 && c < BN) { mystat[c] = a.x; mystat[c + 1] = a.y; mystat[BN + c] = b.x; mystat[BN + c + 1] = b.y; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
    if (p.stats) {
        double *slot = p.stats + (size_t)(blockIdx.x % BN_SLOTS) * 2 * p.F;
        for (int f = threadIdx.x; f < p.F; f += RW_THREADS) {
            double s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int e = 0; e < RW_SETS * 4; ++e) { s1 += (double)sstat[e * 2 * BN + f]; s2 += (double)sstat[(e * 2 + 1) * BN + f]; }
            atomicAdd(slot + f, s1);
            atomicAdd(slot + p.F + f, s2);
        }
    }
}

__global__ void pad_channels4_kernel(const float *__restrict__ x, float4 *__restrict__ x4, long long npix, int C) {
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        const float *s = x + i * C;
        float4 v = make_float4(__ldg(s), 0.f, 0.f, 0.f);
        if (C > 1) v.y = __ldg(s + 1);
        if (C > 2) v.z = __ldg(s + 2);
        if (C > 3) v.w = __ldg(s + 3);
        x4[i] = v;
    }
}

void fill(Im2colP &p, int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw) {
    p.N = N; p.C = C; p.H = H; p.W = W; p.F = F; p.kh = kh; p.kw = kw; p.stride = stride; p.ph = ph; p.pw = pw;
    p.oh = (H + 2 * ph - kh) / stride + 1; p.ow = (W + 2 * pw - kw) / stride + 1;
    p.taps = kh * kw; p.P = N * p.oh * p.ow;
    p.inv_ow = 1.0f / (float)p.ow; p.inv_oh = 1.0f / (float)p.oh;
}

template <int BN, bool OBF16>
int launch_fprop(const float *x4, void *y, const Im2colP &p, cudaStream_t st) {
    const int smem = FpCfg<BN, OBF16>::SMEM;
    CUtensorMap tmX, tmO;
    int e = make_tmap_im2col(&tmX, x4, p.N, p.H, p.W, p.kh, p.kw, p.ph, p.pw, p.stride, 128);
    if (e) return e;
    long long odims[2] = {p.F, p.P};
    long long ostr[2] = {1, p.F};
    int obox[2] = {BN >= 32 ? 32 : 16, 32};
    e = OBF16 ? make_tmap_bf16_sw64(&tmO, y, 2, odims, ostr, obox) : make_tmap_f32(&tmO, y, 2, odims, ostr, obox);
    if (e) return e;
    static bool configured_dev[16] = {};
    bool &configured = configured_dev[current_device_slot()];
    if (!configured) {
        SN_CUDA(cudaFuncSetAttribute(stem_im2col_fprop_kernel<BN, OBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    const int grid = p.tiles < num_sms() ? p.tiles : num_sms();
    SN_CUDA(launch_dependent(stem_im2col_fprop_kernel<BN, OBF16>, dim3(grid), dim3(FP_THREADS), (size_t)smem, st, tmX, tmO, p));
    return after_launch("stem_im2col_fprop_kernel");
}


bool rows_geometry(const Im2colP &p, RowsP &rp) {
    if (p.stride != 1 || p.ow < 8 || p.ow > 64 || (128 % p.ow) != 0) return false;
    rp.rows_per_tile = 128 / p.ow;
    if (p.oh % rp.rows_per_tile) return false;
    rp.tiles_per_image = p.oh / rp.rows_per_tile;
    rp.R = rp.rows_per_tile + p.kh - 1;
    rp.pitch = p.ow * 16;
    rp.copy_bytes = rp.R * rp.pitch;
    rp.stage_bytes = (p.kw * rp.copy_bytes + 1023) / 1024 * 1024;
    // tap slots in address order: v major, u minor
    int s = 0;
    for (int v = 0; v < p.kw; ++v)
        for (int u = 0; u < p.kh; ++u) { rp.tap_off[s] = v * rp.copy_bytes + u * rp.pitch; rp.tap_w[s] = u * p.kw + v; ++s; }
    for (; s < TAP_SLOTS; ++s) { rp.tap_off[s] = -1; rp.tap_w[s] = -1; }
    return true;
}

template <int BN, bool OBF16>
int launch_rows_fprop(const float *x4, void *y, RowsP &rp, cudaStream_t st) {
    const Im2colP &p = rp.g;
    const int smem = RW_STAGES * rp.stage_bytes + TAP_BYTES + TAP_SLOTS * BN * 16 + RW_EPI_WARPS * 4096 + RW_SETS * 4 * 2 * BN * 4 + 1024 + 256;
    if (smem > 232448) { set_error("stem_rows_fprop: %d bytes of shared memory", smem); return SN_ERR_UNSUPPORTED; }
    CUtensorMap tmX, tmO;
    long long xdims[3] = {(long long)p.W * 4, p.H, p.N};
    long long xstr[3] = {1, (long long)p.W * 4, (long long)p.H * p.W * 4};
    int xbox[3] = {p.ow * 4, rp.R, 1};
    int e = make_tmap_linear_f32(&tmX, x4, 3, xdims, xstr, xbox);
    if (e) return e;
    long long odims[2] = {p.F, p.P};
    long long ostr[2] = {1, p.F};
    int obox[2] = {OBF16 ? 64 : 32, 32};
    if (obox[0] > BN) obox[0] = BN;
    e = OBF16 ? make_tmap(&tmO, y, 2, odims, ostr, obox, false, true) : make_tmap_f32(&tmO, y, 2, odims, ostr, obox);
    if (e) return e;
    static int configured_dev[16] = {};
    int &configured = configured_dev[current_device_slot()];
    if (configured < smem) {
        SN_CUDA(cudaFuncSetAttribute(stem_rows_fprop_kernel<BN, OBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    const int grid = p.tiles < num_sms() ? p.tiles : num_sms();
    SN_CUDA(launch_dependent(stem_rows_fprop_kernel<BN, OBF16>, dim3(grid), dim3(RW_THREADS), (size_t)smem, st, tmX, tmO, rp));
    return after_launch("stem_rows_fprop_kernel");
}

}  // namespace

namespace sn {

bool stem_im2col_eligible(int N, int C, int H, int W, int F, int kh, int kw, int stride, int ph, int pw) {
    if (C < 1 || C > 4 || kh * kw > 9 || kh < 1 || kw < 1 || stride != 1) return false;
    if (F < 32 || F > 128 || (F & 31)) return false;           // 32-column epilogue chunks, three accumulators in 512 TMEM columns
    const long long oh = (H + 2 * ph - kh) / stride + 1, ow = (W + 2 * pw - kw) / stride + 1;
    if (oh <= 0 || ow <= 0 || ph >= kh || pw >= kw) return false;
    if ((long long)N * oh * ow >= (1ll << 24) || W > 65535 || H > 65535) return false;
    return get_encode_im2col() != nullptr;
}

int pad_channels4(const float *x, float *x4, long long npix, int C, cudaStream_t st) {
    if (npix <= 0) return SN_OK;
    SN_CUDA(launch_dependent(pad_channels4_kernel, dim3(grid_for(npix, 256)), dim3(256), 0, st, x, reinterpret_cast<float4 *>(x4), npix, C));
    return after_launch("pad_channels4_kernel");
}

int conv_stem_im2col_fprop(const float *x4, const float *w, const float *bias, void *y, int N, int C, int H, int W, int F, int kh, int kw,
                           int stride, int ph, int pw, int act, double *stats, cudaStream_t st, bool out_bf16) {
    Im2colP p{};
    fill(p, N, C, H, W, F, kh, kw, stride, ph, pw);
    p.w = w; p.bias = bias; p.act = act; p.stats = stats;
    p.tiles = (p.P + 127) / 128;
    {
        // whole-row tiles (fat TMA rows) when the geometry allows and the filter tile is at most 64 wide (six TMEM accumulators)
        RowsP rp{};
        rp.g = p;
        const char *env = getenv("SN_STEM_ROWS");
        if (!(env && env[0] == '0') && F <= 64 && (!out_bf16 || F == 64) && rows_geometry(p, rp)) {
            if (out_bf16) return launch_rows_fprop<64, true>(x4, y, rp, st);
            if (F <= 32) return launch_rows_fprop<32, false>(x4, y, rp, st);
            return launch_rows_fprop<64, false>(x4, y, rp, st);
        }
    }
    if (out_bf16) {
        if (F <= 32) return launch_fprop<32, true>(x4, y, p, st);
        if (F <= 64) return launch_fprop<64, true>(x4, y, p, st);
        return launch_fprop<128, true>(x4, y, p, st);
    }
    if (F <= 32) return launch_fprop<32, false>(x4, y, p, st);
    if (F <= 64) return launch_fprop<64, false>(x4, y, p, st);
    return launch_fprop<128, false>(x4, y, p, st);
}

}  // namespace sn
