// Weight gradient of a kh x kw (kw > 1) stride-1 convolution on the tensor pipe with HALO REUSE:
//
//     dW[f][(u,v)][c] += sum_p dY[p][f] * X[p shifted by (u,v)][c]
//
// The kw taps of one filter row u read the SAME input pixels, shifted by one column each.  A CTA
// owns (128 filters) x (CN channels) x (filter row u); per block of 64 output pixels it loads ONE
// dY tile and ONE halo tile of X — th image rows of OW+kw-1 pixels, a TMA box whose out-of-bounds
// zero fill is the padding — and issues the MMAs of all kw taps against them: the B descriptor of
// tap v starts v rows further down the halo tile (UMMA applies the swizzle to the final shared-
// memory address, so a row-shifted start stays consistent with what TMA wrote).
//
// Why this shape (measured on B200, scratch microbenchmarks quoted in DESIGN.md):
//   * a tcgen05.mma with M = 128 costs max(54, N/2) cycles whatever its kind or operand major, so
//     only N >= 128 runs the tensor pipe at full rate and N = 32 caps it at 30 %; here N = CN;
//   * one SM pulls ~50-64 B/clk through TMA, so fp32 operands must be reused: kw taps share dY;
//   * the single issuing thread must not spend more than ~50 cycles per MMA on address arithmetic:
//     every descriptor below is a 32-bit add on a precomputed base.
//
// Both operands are MN-major fp32 in the 128B-swizzle/32B-atom layout (see conv_tc.cu).  Split over
// pixel blocks across CTAs; the partial tiles are added into dW (= the reference's `+=`) by TMA
// REDUCE-ADD stores: an epilogue warp stages its 32 filters x kw taps x 32 channels in the (now idle)
// operand ring, 128-byte swizzled, and one cp.reduce.async.bulk.tensor hands the 12 KB box to the L2
// adders.  (Per-lane red.global.add.v4.f32 — the fallback when the ring is too small — kept the four
// epilogue warps in the LSU for as long as the whole main loop took: ncu, 35 % of all warp samples
// behind the REDG queue.)
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

using namespace sn;
using namespace sn::tc;

namespace {

constexpr int HB_PB = 64;                       // output pixels per stage
constexpr int HB_ATOM = HB_PB * 128;            // one atom column: 64 pixel rows x 128 B (32 fp32 / 64 bf16 filters or channels)
constexpr int HB_THREADS = 192;
constexpr int HB_MAX_KW = 5;
// fp32 (tf32) operands: 4 atom columns of dY, 8-pixel K-steps, K groups of 4 rows (32-byte-atom swizzle)
// bf16 operands       : 2 atom columns of dY, 16-pixel K-steps, K groups of 8 rows (plain 128-byte swizzle)
template <bool BF16> struct HaloT {
    static constexpr int A_ATOMS = BF16 ? 2 : 4;
    static constexpr int A_BYTES = A_ATOMS * HB_ATOM;
    static constexpr int KSTEP_PX = BF16 ? 16 : 8;
    static constexpr int KSTEPS = HB_PB / KSTEP_PX;
    static constexpr int ELEMS = BF16 ? 64 : 32;          // elements per 128-byte atom row
    static constexpr uint32_t LAYOUT = BF16 ? LAYOUT_SW128 : LAYOUT_SW128_BASE32B;
    static constexpr uint32_t KGROUP_BYTES = BF16 ? 1024 : 512;
};

struct HaloP {
    float *dw;
    int F, C, ntaps, kw, kh;
    int c_tiles, f_tiles, splits, blocks_per_split, blocks;   // blocks = pixel blocks of 64
    int OW, S;                 // S = OH*OW
    int th, tn;                // pixel block = tn images x th rows x OW
    int Wp;                    // halo row width OW + kw - 1
    int b_atom_bytes;          // tn*th*Wp*128, rounded up to 1024
    int stage_bytes, stages;
    int pad_h, pad_w;
    int krow[HB_PB / 8];       // halo row index of the first pixel of each K-step
    unsigned b_sbo;            // stride between the two K groups of one K-step in the halo tile
    int red_bufs;              // staging buffers (kw*32 rows x 128 B) per epilogue warp for the TMA reduce-add epilogue; 0: red.global
};

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// dW as (C, taps, F): box (32 channels, kw taps, 32 filters), 128-byte swizzle; filters past F are clipped by the TMA
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *m, const void *src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <int CN, bool BF16>
__global__ void __launch_bounds__(HB_THREADS, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                       const __grid_constant__ CUtensorMap tmDW, const HaloP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + p.stages * p.stage_bytes);
    uint64_t *empty_bar = full_bar + 8;
    uint64_t *acc_full = empty_bar + 8;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_full + 1);
    using T = HaloT<BF16>;
    constexpr int ATOMS = CN / T::ELEMS;
    constexpr uint32_t TMEM_COLS = 512;

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int split = blockIdx.x % p.splits;
    int item = blockIdx.x / p.splits;
    const int u = item % p.kh; item /= p.kh;
    const int ct = item % p.c_tiles;
    const int ft = item / p.c_tiles;
    const int b0 = split * p.blocks_per_split;
    const int b1 = min(p.blocks, b0 + p.blocks_per_split);
    const int nst = b1 - b0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmDY);
        tma_prefetch_desc(&tmX);
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(tmem_ptr, TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);     // shuffle from lane 0: provably warp-uniform
    pdl_wait();

    if (nst > 0) {
        if (warp == 0) {
            {   // warp-uniform control flow, lane 0 issues (TMA / MMA operands live in uniform registers)
                const uint32_t tx_bytes = T::A_BYTES + ATOMS * p.tn * p.th * p.Wp * 128;
                int stage = 0; uint32_t phase = 0;
                int pix = b0 * HB_PB;
                int n0 = pix / p.S, i0 = (p.S >= HB_PB) ? (pix - n0 * p.S) / p.OW : 0;
                for (int b = b0; b < b1; ++b) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t *sa = smem + stage * p.stage_bytes;
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], tx_bytes);
#pragma unroll
                        for (int a = 0; a < T::A_ATOMS; ++a)
                            tma_load_2d(sa + a * HB_ATOM, &tmDY, &full_bar[stage], ft * 128 + a * T::ELEMS, pix);
#pragma unroll
                        for (int a = 0; a < ATOMS; ++a)
                            tma_load_4d(sa + T::A_BYTES + a * p.b_atom_bytes, &tmX, &full_bar[stage], ct * CN + a * T::ELEMS,
                                        -p.pad_w, i0 + u - p.pad_h, n0);
                    }
                    __syncwarp();
                    pix += HB_PB;
                    if (p.S >= HB_PB) { i0 += p.th; if (i0 * p.OW >= p.S) { i0 = 0; ++n0; } } else { n0 += p.tn; }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            {
                constexpr uint32_t idesc = make_idesc(BF16 ? 1 : 2, 128, CN, 1, 1);      // both operands MN-major
                // descriptor templates: everything but the 14-bit start address
                const uint64_t a_hi = make_smem_desc(0, HB_ATOM, T::KGROUP_BYTES, T::LAYOUT);
                const uint64_t b_hi = make_smem_desc(0, p.b_atom_bytes, p.b_sbo, T::LAYOUT);
                uint32_t krow8[T::KSTEPS];                                   // K-step start rows, in 16-byte units
#pragma unroll
                for (int k = 0; k < T::KSTEPS; ++k) krow8[k] = (uint32_t)p.krow[k] * 8u;
                const int kw = p.kw;
                int stage = 0; uint32_t phase = 0;
                uint32_t accum = 0;
                for (int s = 0; s < nst; ++s) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa16 = smem_u32(smem + stage * p.stage_bytes) >> 4;
                    const uint32_t sb16 = sa16 + (T::A_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < T::KSTEPS; ++k) {
                        const uint64_t adesc = a_hi | (uint64_t)(sa16 + k * (T::KSTEP_PX * 8));   // K-step = KSTEP_PX rows of 128 B
                        const uint32_t brow = sb16 + krow8[k];
#pragma unroll
                        for (int v = 0; v < HB_MAX_KW; ++v) {
                            if (v < kw && elect_one()) {
                                if constexpr (BF16) mma_bf16(tmem_base + v * CN, adesc, b_hi | (uint64_t)(brow + v * 8), idesc, accum);
                                else mma_tf32(tmem_base + v * CN, adesc, b_hi | (uint64_t)(brow + v * 8), idesc, accum);
                            }
                        }
                        accum = 1;
                    }
                    if (elect_one()) mma_commit(&empty_bar[stage]);
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) mma_commit(acc_full);
                __syncwarp();
            }
        } else {
            const int q = warp & 3;
            mbar_wait_sleep(acc_full, 0);
            tc_fence_after();
            const int f = ft * 128 + q * 32 + lane;
            if (p.red_bufs > 0) {
                // every MMA has completed, so every operand stage has been consumed: the ring is staging space now
                const int buf_bytes = p.kw * 32 * 128;
                uint8_t *mine = smem + q * p.red_bufs * buf_bytes;
                int nb = 0;
#pragma unroll 1
                for (int c0 = 0; c0 < CN; c0 += 32, ++nb) {
                    uint8_t *st = mine + (nb % p.red_bufs) * buf_bytes;
                    if (nb >= p.red_bufs) {                   // the box that used this buffer must have been read
                        if (lane == 0) { if (p.red_bufs == 1) tma_store_wait_read(); else tma_store_wait_read1(); }
                        __syncwarp();
                    }
#pragma unroll 1
                    for (int v = 0; v < p.kw; ++v) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(v * CN + c0), r);
                        tmem_ld_wait();
                        const int row = lane * p.kw + v;      // box order: channel fastest, then tap, then filter
                        uint8_t *dst = st + row * 128;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<uint4 *>(dst + ((j ^ (row & 7)) * 16)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_3d(&tmDW, st, ct * CN + c0, u * p.kw, ft * 128 + q * 32);
                        tma_store_commit();
                    }
                }
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            } else
#pragma unroll 1
            for (int v = 0; v < p.kw; ++v) {
                float *row = p.dw + ((long long)f * p.ntaps + u * p.kw + v) * p.C + ct * CN;
#pragma unroll 1
                for (int c0 = 0; c0 < CN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(v * CN + c0), r);
                    tmem_ld_wait();
                    if (f < p.F) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            red_add_v4(row + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                       __uint_as_float(r[j + 3]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// a K-step (8 px fp32 / 16 px bf16) is two K groups (4 / 8 halo rows each): either both lie in one
// image row (contiguous), or each is exactly one image row (OW == group size) and they are Wp rows apart
bool block_geometry(int OH, int OW, bool bf16, int &th, int &tn) {
    const int group = bf16 ? 8 : 4;
    if (OW != group && OW != 2 * group && OW != 4 * group && OW != 8 * group && OW != 16 * group) return false;
    if (OW > 64) return false;
    const int S = OH * OW;
    if (S >= HB_PB) { if (S % HB_PB) return false; th = HB_PB / OW; tn = 1; }
    else { if (HB_PB % S) return false; th = OH; tn = HB_PB / S; }
    if (OW == group && (th & 1)) return false;
    return true;
}

int cn_for(int C, int kw) {
    if (C % 128 == 0 && kw * 128 <= 512) return 128;
    if (C % 64 == 0 && kw * 64 <= 512) return 64;
    return 0;
}

template <int CN, bool BF16>
int launch_halo(const CUtensorMap &tmDY, const CUtensorMap &tmX, const CUtensorMap &tmDW, const HaloP &p, int grid, int smem, cudaStream_t st) {
    static int configured_dev[16] = {};
    int &configured = configured_dev[current_device_slot()];
    if (configured < smem) {
        SN_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel<CN, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    SN_CUDA(launch_dependent(conv_wgrad_halo_kernel<CN, BF16>, dim3(grid), dim3(HB_THREADS), (size_t)smem, st, tmDY, tmX, tmDW, p));
    return after_launch("conv_wgrad_halo_kernel");
}

}  // namespace

namespace sn {

bool wgrad_halo_eligible(int N, int C, int OH, int OW, int F, int kh, int kw, int stride, bool bf16) {
    int th, tn;
    if (stride != 1 || kw < 2 || kw > HB_MAX_KW || kh > 16 || (F & (bf16 ? 7 : 3))) return false;
    if (!cn_for(C, kw) || !block_geometry(OH, OW, bf16, th, tn)) return false;
    if (OW + kw - 1 > 256 || th > 256 || tn > 256) return false;
    const int CN = cn_for(C, kw);
    const int b_atom = ((tn * th * (OW + kw - 1) * 128) + 1023) / 1024 * 1024;
    const int a_bytes = bf16 ? HaloT<true>::A_BYTES : HaloT<false>::A_BYTES;
    if (a_bytes + (CN / (bf16 ? 64 : 32)) * b_atom > 100 * 1024) return false;          // at least two stages
    long long flops = 2ll * N * OH * OW * F * C * kh * kw;
    return flops >= (1ll << 26);
}

// x: NHWC (N, IH, IW, C); dy: (N*OH*OW, F); dw: [F][kh*kw][C] accumulated into (+=).
int conv_wgrad_halo_tc(const void *x, const void *dy, float *dw, int N, int C, int IH, int IW, int F, int kh, int kw, int OH,
                       int OW, int pad_h, int pad_w, bool bf16, cudaStream_t st) {
    if (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) { set_error("conv_wgrad_halo_tc: tensors must be 16-byte aligned"); return SN_ERR_INVALID_ARG; }
    HaloP p{};
    const int CN = cn_for(C, kw);
    p.dw = dw; p.F = F; p.C = C; p.ntaps = kh * kw; p.kw = kw; p.kh = kh;
    p.OW = OW; p.S = OH * OW;
    if (!CN || !block_geometry(OH, OW, bf16, p.th, p.tn)) { set_error("conv_wgrad_halo_tc: unsupported geometry"); return SN_ERR_UNSUPPORTED; }
    const int elems = bf16 ? 64 : 32, kstep_px = bf16 ? 16 : 8, group = bf16 ? 8 : 4;
    const int a_bytes = bf16 ? HaloT<true>::A_BYTES : HaloT<false>::A_BYTES;
    p.Wp = OW + kw - 1;
    p.b_atom_bytes = ((p.tn * p.th * p.Wp * 128) + 1023) / 1024 * 1024;
    p.stage_bytes = a_bytes + (CN / elems) * p.b_atom_bytes;
    p.stages = (200 * 1024) / p.stage_bytes;
    if (p.stages > 6) p.stages = 6;
    p.pad_h = pad_h; p.pad_w = pad_w;
    p.b_sbo = (OW >= 2 * group) ? (unsigned)group * 128u : (unsigned)p.Wp * 128u;
    for (int k = 0; k < HB_PB / kstep_px; ++k) {
        const int pl = k * kstep_px;
        const int nl = pl / (p.th * OW), rem = pl % (p.th * OW);
        p.krow[k] = (nl * p.th + rem / OW) * p.Wp + rem % OW;
    }
    const long long P = (long long)N * p.S;
    p.blocks = (int)((P + HB_PB - 1) / HB_PB);
    p.c_tiles = C / CN; p.f_tiles = (F + 127) / 128;
    const int items = p.c_tiles * p.f_tiles * kh;
    int splits = num_sms() / items;
    if (splits < 1) splits = 1;
    if (splits > p.blocks) splits = p.blocks;
    p.blocks_per_split = (p.blocks + splits - 1) / splits;
    p.splits = (p.blocks + p.blocks_per_split - 1) / p.blocks_per_split;
    CUtensorMap tmDY, tmX;
    long long ddims[2] = {F, P};
    long long dstr[2] = {1, F};
    int dbox[2] = {elems, HB_PB};
    int e = make_tmap(&tmDY, dy, 2, ddims, dstr, dbox, !bf16, bf16);
    if (e) return e;
    long long xdims[4] = {C, IW, IH, N};
    long long xstr[4] = {1, C, (long long)IW * C, (long long)IH * IW * C};
    int xbox[4] = {elems, p.Wp, p.th, p.tn};
    e = make_tmap(&tmX, x, 4, xdims, xstr, xbox, !bf16, bf16);
    if (e) return e;
    // reduce-add epilogue: as many staging buffers per epilogue warp as the idle ring holds (all chunks, else 2, else 1)
    static const bool force_redg = [] { const char *e = getenv("SN_WGRAD_REDG"); return e && e[0] == '1'; }();
    const int chunks = CN / 32, buf_bytes = kw * 32 * 128, ring = p.stages * p.stage_bytes;
    p.red_bufs = force_redg ? 0 : (4 * chunks * buf_bytes <= ring ? chunks : (8 * buf_bytes <= ring ? 2 : (4 * buf_bytes <= ring ? 1 : 0)));
    CUtensorMap tmDW;
    long long wdims[3] = {C, p.ntaps, F};
    long long wstr[3] = {1, C, (long long)p.ntaps * C};
    int wbox[3] = {32, kw, 32};
    e = make_tmap_f32(&tmDW, dw, 3, wdims, wstr, wbox, false);
    if (e) return e;
    const int smem = p.stages * p.stage_bytes + 1024 + 256;
    const int grid = items * p.splits;
    if (bf16) return CN == 128 ? launch_halo<128, true>(tmDY, tmX, tmDW, p, grid, smem, st) : launch_halo<64, true>(tmDY, tmX, tmDW, p, grid, smem, st);
    return CN == 128 ? launch_halo<128, false>(tmDY, tmX, tmDW, p, grid, smem, st) : launch_halo<64, false>(tmDY, tmX, tmDW, p, grid, smem, st);
}

}  // namespace sn
