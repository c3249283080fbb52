// Implicit-GEMM convolution with the FILTERS RESIDENT in shared memory and the activation taps taken from
// row-shifted "fat" copies — bf16 operands, stride 1, pixel tiles that lie inside one image (16x16 maps and
// larger: config 3's conv2 forward and dgrad, config 4's 64-channel blocks).
//
// Why: conv_igemm_kernel (conv_tc.cu) loads one 16 KB A tile and one BN x 128 B filter tile per (tap, channel
// block): 128 + BN rows of 128 bytes per 256 (BN = 128) MMA clocks, against an SM that ingests one such row
// every ~2 clocks (~64 B/clk from L2).  At batch 512 that — not the tensor pipe — is what the 128-wide tiles
// run at (tensor pipe active 41-47 % in ncu).  Two things remove most of those rows:
//   * the whole filter matrix [F][kh*kw*C] of these layers is <= 147 KB in bf16: it is loaded ONCE per CTA and
//     stays; no filter traffic per tile at all;
//   * the kh*kw activation tiles of a pixel tile overlap almost entirely.  Copy v (one per filter column) holds
//     image rows i0-ph .. i0-ph+R-1 (R = rows per tile + kh - 1), OW pixels wide, shifted horizontally by v - pw
//     — the shift, the halo columns and the padding rows are the TMA box start plus out-of-bounds zero fill.
//     Row pitch = OW pixels, so output pixel m of tap (u, v) is row u*OW + m of copy v: linear in m, i.e. the
//     same 128-byte-swizzled K-major operand as before, started u*OW rows further down (OW is a multiple of 8,
//     so the start stays aligned to the 1024-byte swizzle atom).  kw loads of R*OW rows per channel block
//     instead of kh*kw loads of 128 rows (conv2: 480 rows instead of 1152).
// One pipeline stage = one (filter column, channel block) copy = kh taps x 4 MMAs (K = 16 each).
// Roles, TMEM double buffering and the epilogue (bias / ReLU, fp32 or bf16 output through TMA stores, BatchNorm
// column sums) are those of conv_igemm_kernel; the column sums are kept in registers until the end.
#include "common.cuh"
#include "tc_common.cuh"
#include "bn_common.cuh"
#include <cuda_bf16.h>

using namespace sn;
using namespace sn::tc;

namespace {

constexpr int RI_EPI_WARPS = 8;
constexpr int RI_THREADS = 64 + 32 * RI_EPI_WARPS;

struct RowsIgemmP {
    const float *bias;
    double *stats;
    int rows, F, C, kh, kw, cblocks, OW, OH;
    int rpt, tiles_per_image, R, copy_bytes, tiles, stages, kb_total;
    int pad_h, pad_w, act;
};

template <int BN, bool OBF16>
__global__ void __launch_bounds__(RI_THREADS, 1)
conv_rows_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmO, const RowsIgemmP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int B_KB_BYTES = BN * 128;                              // one (tap, channel block) of the filter matrix
    constexpr int EPI_TILE = OBF16 ? 2048 : 4096;
    uint8_t *sB = smem;                                               // kb_total x [BN][128 B], SW128 K-major
    uint8_t *sA = sB + p.kb_total * B_KB_BYTES;                       // stages x copy_bytes
    uint8_t *epi_stage = sA + p.stages * p.copy_bytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(epi_stage + RI_EPI_WARPS * EPI_TILE);
    uint64_t *empty_bar = full_bar + 8;
    uint64_t *b_full = empty_bar + 8;
    uint64_t *tmem_full = b_full + 1;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *sbias = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(full_bar) + 256);          // (16-byte aligned) the bias, once: an __ldg per chunk and tile was eight L2 round trips in the epilogue's chain

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], RI_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_ptr, 2 * BN < 32 ? 32 : 2 * BN); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    pdl_wait();
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < BN; i += 32 * RI_EPI_WARPS) sbias[i] = (p.bias && i < p.F) ? __ldg(p.bias + i) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * RI_EPI_WARPS) : "memory");
    }

    if (warp == 0) {
        // ===================== TMA producer =====================
        // the filter matrix, once: kb_total boxes of BN rows x 64 elements, K ordered (tap, channel)
        if (elect_one()) mbar_expect_tx(b_full, (uint32_t)(p.kb_total * B_KB_BYTES));
        __syncwarp();
        for (int kb = 0; kb < p.kb_total; ++kb) {
            const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
            if (elect_one()) tma_load_2d(sB + kb * B_KB_BYTES, &tmB, b_full, tap * p.C + cb * 64, 0);
        }
        __syncwarp();
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const int n = tile / p.tiles_per_image, i0 = (tile - n * p.tiles_per_image) * p.rpt;
            for (int v = 0; v < p.kw; ++v)
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], (uint32_t)p.copy_bytes);
                        tma_load_4d(sA + stage * p.copy_bytes, &tmA, &full_bar[stage], cb * 64, v - p.pad_w, i0 - p.pad_h, n);
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc(1, 128, BN, 0, 0);
        mbar_wait(b_full, 0);
        tc_fence_after();
        int stage = 0; uint32_t phase = 0;
        int iter = 0;
        const uint32_t b0 = smem_u32(sB);
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
            const int acc = iter & 1;
            mbar_wait(&tmem_empty[acc], ((iter >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            uint32_t accum = 0;
            for (int v = 0; v < p.kw; ++v)
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + stage * p.copy_bytes);
                    if (elect_one()) {
                        for (int u = 0; u < p.kh; ++u) {
                            // tap (u, v): the same copy, u image rows (u*OW pixel rows of 128 B) further down
                            const uint64_t adesc = make_smem_desc(a0 + u * p.OW * 128, 0, 1024);
                            const uint64_t bdesc = make_smem_desc(b0 + ((u * p.kw + v) * p.cblocks + cb) * B_KB_BYTES, 0, 1024);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
                                accum = 1u;
                            }
                        }
                        mma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    accum = 1u;
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            if (elect_one()) mma_commit(&tmem_full[acc]);
            __syncwarp();
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int WPQ = RI_EPI_WARPS / 4;          // warps per TMEM lane quarter
        constexpr int NCH = (BN / 32 + WPQ - 1) / WPQ; // 32-column chunks this warp owns per tile
        float2 cs[NCH], cq[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) { cs[i] = make_float2(0.f, 0.f); cq[i] = make_float2(0.f, 0.f); }
        uint8_t *st = epi_stage + (warp - 2) * EPI_TILE;
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++iter) {
            const int acc = iter & 1;
            mbar_wait(&tmem_full[acc], (iter >> 1) & 1);
            tc_fence_after();
            const int nvalid = p.rows - (tile * 128 + q * 32);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = half * 32 + ci * 32 * WPQ;
                if (c0 >= BN) break;
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
                tmem_ld_wait();
                if (c0 >= p.F) continue;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = *reinterpret_cast<const float4 *>(sbias + c0 + j);
                    v[j] = __uint_as_float(r[j]) + b.x; v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
                    v[j + 2] = __uint_as_float(r[j + 2]) + b.z; v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
                    if (p.act == SN_ACT_RELU) {
                        v[j] = relu_keep_sign(v[j]); v[j + 1] = relu_keep_sign(v[j + 1]);
                        v[j + 2] = relu_keep_sign(v[j + 2]); v[j + 3] = relu_keep_sign(v[j + 3]);
                    }
                }
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
                if constexpr (OBF16) {
                    stage_row_bf16(st, lane, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(st + lane * 128 + (((j >> 2) ^ (lane & 7)) * 16)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_2d(&tmO, st, c0, tile * 128 + q * 32); tma_store_commit(); }
                if (p.stats) {
                    if constexpr (OBF16) {
                        const int pr = lane & 15, rsel = lane >> 4;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int rr = 2 * i + rsel;
                            const uint32_t w = *reinterpret_cast<const uint32_t *>(st + rr * 64 + (((pr >> 2) ^ ((rr >> 1) & 3)) * 16) + (pr & 3) * 4);
                            float2 x2 = make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
                            if (rr >= nvalid) x2 = make_float2(0.f, 0.f);
                            cs[ci] = __fadd2_rn(cs[ci], x2);
                            cq[ci] = __ffma2_rn(x2, x2, cq[ci]);
                        }
                    } else {
                        const int cp = lane & 15, r0 = (lane >> 4) * 16;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int rr = r0 + i;
                            float2 x2 = *reinterpret_cast<const float2 *>(st + (cp & 1) * 8 + rr * 128 + (((cp >> 1) ^ (rr & 7)) * 16));
                            if (rr >= nvalid) x2 = make_float2(0.f, 0.f);
                            cs[ci] = __fadd2_rn(cs[ci], x2);
                            cq[ci] = __ffma2_rn(x2, x2, cq[ci]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (lane == 0) tma_store_wait_all();
        __syncwarp();
        if (p.stats) {
            // every tile of this CTA is done (its accumulators were read), so the activation ring is free: the per-warp
            // tables go there.  table[quarter][{sum, sumsq}][BN]; the two warps of a quarter own disjoint columns
            float *tab = reinterpret_cast<float *>(sA);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c0 = half * 32 + ci * 32 * WPQ;
                if (c0 >= BN) break;
                float2 a = cs[ci], b = cq[ci];
                a.x += __shfl_down_sync(0xffffffffu, a.x, 16); a.y += __shfl_down_sync(0xffffffffu, a.y, 16);
                b.x += __shfl_down_sync(0xffffffffu, b.x, 16); b.y += __shfl_down_sync(0xffffffffu, b.y, 16);
                if (lane < 16) {
                    const int c = c0 + 2 * lane;
                    tab[(q * 2) * BN + c] = a.x; tab[(q * 2) * BN + c + 1] = a.y;
                    tab[(q * 2 + 1) * BN + c] = b.x; tab[(q * 2 + 1) * BN + c + 1] = b.y;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * RI_EPI_WARPS) : "memory");           // the epilogue warps only
            double *slot = p.stats + (size_t)(blockIdx.x % BN_SLOTS) * 2 * p.F;
            for (int f = threadIdx.x - 64; f < p.F; f += 32 * RI_EPI_WARPS) {
                double s1 = 0.0, s2 = 0.0;
#pragma unroll
                for (int wq = 0; wq < 4; ++wq) { s1 += (double)tab[(wq * 2) * BN + f]; s2 += (double)tab[(wq * 2 + 1) * BN + f]; }
                atomicAdd(slot + f, s1);
                atomicAdd(slot + p.F + f, s2);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN < 32 ? 32 : 2 * BN); }
}

bool rows_geom(int N, int C, int OH, int OW, int F, int kh, int kw, int pad_h, int pad_w, int IH, int IW, bool out_bf16, RowsIgemmP &p, int &smem,
               int &BN) {
    if (OW < 16 || OW > 64 || (128 % OW) != 0 || (OW & 7)) return false;
    const int rpt = 128 / OW;
    if (OH % rpt) return false;
    if ((C & 63) || F > 128 || (F & 31) || kh < 1 || kw < 1 || kh > 7 || kw > 7) return false;
    if (OH != IH + 2 * pad_h - kh + 1 || OW != IW + 2 * pad_w - kw + 1) return false;
    BN = F <= 64 ? 64 : 128;
    p.cblocks = C / 64;
    p.kb_total = kh * kw * p.cblocks;
    const long long b_bytes = (long long)p.kb_total * BN * 128;
    p.rpt = rpt; p.tiles_per_image = OH / rpt; p.R = rpt + kh - 1;
    if (p.R > 256) return false;
    p.copy_bytes = p.R * OW * 128;                       // a multiple of 1024 (OW is a multiple of 8)
    const int epi = RI_EPI_WARPS * (out_bf16 ? 2048 : 4096);
    const long long room = 232448 - 1024 - 1024 - epi - b_bytes;
    if (room < 3ll * p.copy_bytes) return false;
    p.stages = (int)(room / p.copy_bytes);
    if (p.stages > 8) p.stages = 8;
    if (p.stages < 3) return false;                              // measured: with a two-deep ring (conv2's dgrad: fp32 staging tiles leave
                                                                 // 50 KB) the kernel waits on load latency: 36.7 us against 33.7 for conv_igemm_kernel
    if (8 * BN * 4 > p.stages * p.copy_bytes) return false;      // the statistics tables reuse the ring
    smem = (int)(b_bytes + (long long)p.stages * p.copy_bytes + epi + 1024 + 1024);
    p.tiles = N * p.tiles_per_image;
    p.rows = N * OH * OW; p.F = F; p.C = C; p.kh = kh; p.kw = kw; p.OW = OW; p.OH = OH; p.pad_h = pad_h; p.pad_w = pad_w;
    return true;
}

template <int BN, bool OBF16>
int launch_rows(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmO, const RowsIgemmP &p, int smem, cudaStream_t st) {
    static int configured_dev[16] = {};
    int &configured = configured_dev[current_device_slot()];
    if (configured < smem) {
        SN_CUDA(cudaFuncSetAttribute(conv_rows_igemm_kernel<BN, OBF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    const int grid = p.tiles < num_sms() ? p.tiles : num_sms();
    SN_CUDA(launch_dependent(conv_rows_igemm_kernel<BN, OBF16>, dim3(grid), dim3(RI_THREADS), (size_t)smem, st, tmA, tmB, tmO, p));
    return after_launch("conv_rows_igemm_kernel");
}

}  // namespace

namespace sn {

bool conv_rows_eligible(int N, int C, int IH, int IW, int F, int kh, int kw, int OH, int OW, int pad_h, int pad_w, bool out_bf16) {
    const char *e = getenv("SN_IGEMM_ROWS");
    if (e && e[0] == '0') return false;
    RowsIgemmP p{};
    int smem = 0, BN = 0;
    if (!rows_geom(N, C, OH, OW, F, kh, kw, pad_h, pad_w, IH, IW, out_bf16, p, smem, BN)) return false;
    return (long long)N * OH * OW >= 128ll * 148;          // at least one tile per SM: below that the resident filter load is not amortised
}

// x: bf16 NHWC (N, IH, IW, C); w: bf16 [F][kh*kw][C]; out: (N*OH*OW, F) fp32 or bf16
int conv_rows_igemm_tc(const void *x, const void *w, const float *bias, void *out, int N, int C, int IH, int IW, int F, int kh, int kw,
                       int OH, int OW, int pad_h, int pad_w, int act, double *stats, bool out_bf16, cudaStream_t st) {
    RowsIgemmP p{};
    int smem = 0, BN = 0;
    if (!rows_geom(N, C, OH, OW, F, kh, kw, pad_h, pad_w, IH, IW, out_bf16, p, smem, BN)) { set_error("conv_rows_igemm_tc: unsupported geometry"); return SN_ERR_UNSUPPORTED; }
    p.bias = bias; p.stats = stats; p.act = act;
    CUtensorMap tmA, tmB, tmO;
    long long adims[4] = {C, IW, IH, N};
    long long astr[4] = {1, C, (long long)IW * C, (long long)IH * IW * C};
    int abox[4] = {64, OW, p.R, 1};
    int e = make_tmap(&tmA, x, 4, adims, astr, abox, false, true);
    if (e) return e;
    const long long ktot = (long long)kh * kw * C;
    long long bdims[2] = {ktot, F};
    long long bstr[2] = {1, ktot};
    int bbox[2] = {64, BN};
    e = make_tmap(&tmB, w, 2, bdims, bstr, bbox, false, true);
    if (e) return e;
    long long odims[2] = {F, p.rows};
    long long ostr[2] = {1, F};
    int obox[2] = {32, 32};
    e = out_bf16 ? make_tmap_bf16_sw64(&tmO, out, 2, odims, ostr, obox) : make_tmap_f32(&tmO, out, 2, odims, ostr, obox);
    if (e) return e;
    if (BN == 64) return out_bf16 ? launch_rows<64, true>(tmA, tmB, tmO, p, smem, st) : launch_rows<64, false>(tmA, tmB, tmO, p, smem, st);
    return out_bf16 ? launch_rows<128, true>(tmA, tmB, tmO, p, smem, st) : launch_rows<128, false>(tmA, tmB, tmO, p, smem, st);
}

}  // namespace sn
