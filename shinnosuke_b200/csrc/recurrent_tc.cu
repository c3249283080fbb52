// LSTM recurrence on the tensor pipe (bf16 tier; BASELINE config 5 = the "recurrent GEMM path").
//
// layers/Recurrent.py:293-309 runs, per timestep, ot = [h_{t-1}, x_t] . W + b and the gate / cell update; on the
// device the input half is one GEMM over all timesteps (zx, recurrent.cu) and what is left per step is
//     z_t = zx_t + h_{t-1} . Wh^T          (N x H) . (H x 4H)        forward
//     dh_t = dh_ext_t + dz_{t+1} . Wh      (N x 4H) . (4H x H)       backward
// 33 MFLOP per step at batch 256: pure latency.  Here a CTA owns 2 batch rows for the WHOLE sequence and
// the weight matrix never leaves its shared memory: Wh (forward: 4 gate tiles of 128 units x 128 k; backward:
// Wh^T, 128 units x 512 k) is staged ONCE as the K-major, 128-byte-swizzled bf16 A operand (128 KB), and the
// recurrent operand — h_{t-1} or dz_{t+1} of the CTA's rows, 16 x K with rows 2..15 zero (N = 16 is the smallest
// UMMA N at M = 128) — is the B operand the cell threads rewrite every step.  So D = A . B^T lands TRANSPOSED
// in tensor memory: lane = hidden unit, column = batch row, and the thread that owns the cell (unit, row) reads
// the pre-activations of its four gates with four one-column tcgen05.ld — no shared-memory reduction, no
// cluster, no DSMEM exchange (the CUDA-core kernels in recurrent.cu need all three).  Per step: 32 tcgen05.mma
// (M128 x N16 x K16, four independent accumulators taken in turn so that consecutive MMAs never wait for each
// other) issued by one thread -> commit -> mbarrier -> tcgen05.ld -> gates -> bf16 h / dz back into the B tile
// -> fence.proxy.async -> __syncthreads.  The cell state (forward) / its gradient (backward) stay in registers.
// Two rows per CTA = one cell per thread and 128 CTAs at batch 256: the ~185 instructions of a cell's five
// transcendentals are on the critical path of every step (8 rows per CTA, four cells per thread, measured
// 2.28 us per step against 1.98 for the CUDA-core kernels — the MMAs were not the long pole).
// fp32 everywhere except the two MMA operands (bf16, fp32 accumulation) and SFU-approximated gates: the bf16 tier's 1e-2.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

using namespace sn;
using namespace sn::tc;

namespace {

constexpr int LT_H = 128;
constexpr int LT_ROWS = 2;                 // live batch rows per CTA
constexpr int LT_N = 16;                   // UMMA N (M = 128 needs N % 16 == 0): rows 2..15 of the B tile stay zero
constexpr int LT_CPT = LT_ROWS / 2;        // cells per thread
constexpr int LT_ISSUERS = 1;              // warps that issue a step's MMAs (1, 2 or 4; two issuers on different SM
                                           // sub-partitions measured exactly the same step time as one: 78.2 / 84.8 us)
constexpr int LT_THREADS = 256;            // thread = (unit, half): cells (rows LT_CPT*half .. +LT_CPT-1, unit)
constexpr int LT_ATOM_A = 128 * 128;       // one K-atom (64 bf16 = 128 B per row) of a 128-row operand tile
constexpr int LT_ATOM_B = LT_N * 128;
constexpr int LT_A_BYTES = 8 * LT_ATOM_A;  // forward: 4 gate tiles x 2 K-atoms; backward: 1 tile x 8 K-atoms

// Gate non-linearities (utils/Activator.py:58, :81) on the special-function unit: ex2.approx + rcp.approx, ~1e-6
// absolute — three orders of magnitude below the bf16 rounding of the recurrent operand this tier already accepts.
// expf / tanhf / IEEE division cost ~185 instructions per cell, all of them on the critical path of every step.
__device__ __forceinline__ float lt_sigmoid(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
__device__ __forceinline__ float lt_tanh(float z) { return fmaf(2.f, lt_sigmoid(2.f * z), -1.f); }

// LT_CPT consecutive columns of this warp's 32 lanes
template <int CPT>
__device__ __forceinline__ void lt_tmem_ld(uint32_t taddr, uint32_t (&r)[CPT]) {
    if constexpr (CPT == 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
    } else if constexpr (CPT == 2) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
    } else {
        static_assert(CPT == 4, "1, 2 or 4 cells per thread");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(taddr)
                     : "memory");
    }
}
__device__ __forceinline__ uint32_t lt_pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// byte offset of element (row r, k) in a K-major SW128 bf16 tile whose K-atoms (64 elements) are `atom` bytes apart
__device__ __forceinline__ uint32_t lt_off(int r, int k, int atom) {
    return (uint32_t)((k >> 6) * atom + r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + ((k & 7) << 1));
}
__device__ __forceinline__ void lt_store_bf16(uint8_t *tile, int r, int k, float v) {
    *reinterpret_cast<__nv_bfloat16 *>(tile + lt_off(r, k, LT_ATOM_B)) = __float2bfloat16_rn(v);
}

struct LtSmem {
    uint8_t *a, *b;
    uint64_t *done;
    uint32_t *tmem_ptr;
};
template <int B_ATOMS>
__device__ __forceinline__ LtSmem lt_carve(uint8_t *raw) {
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    LtSmem s;
    s.a = smem;
    s.b = smem + LT_A_BYTES;
    s.done = reinterpret_cast<uint64_t *>(s.b + B_ATOMS * LT_ATOM_B);
    s.tmem_ptr = reinterpret_cast<uint32_t *>(s.done + 1);
    return s;
}
template <int B_ATOMS> constexpr int lt_smem_bytes() { return LT_A_BYTES + B_ATOMS * LT_ATOM_B + 64 + 1024; }

// ------------------------------------------------------------------------------------------ forward
// zx: (T, N, 4H) input projections + bias; wh: [4H][H] (gate-major rows f, u, c~, o); acts (T, N, 4H), c_all / h_all
// (T + 1, N, H) with slot 0 = the initial state (read when h0 / c0 are given); h_seq (N, T, H) or NULL.
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_seq_fwd_tc_kernel(const float *__restrict__ zx, const float *__restrict__ wh, const float *__restrict__ h0,
                       const float *__restrict__ c0, float *__restrict__ acts, float *__restrict__ c_all,
                       float *__restrict__ h_all, float *__restrict__ h_seq, int T, int N) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int H = LT_H;
    const LtSmem s = lt_carve<2>(smem_raw);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int unit = q * 32 + lane;
    const int row0 = blockIdx.x * LT_ROWS;

    // A: row = gate column g*H + unit (tile g), 16 chunks of 8 k each
    for (int i = threadIdx.x; i < 4 * H * 16; i += LT_THREADS) {
        const int row = i >> 4, ch = i & 15;
        const float4 lo = ldg4(wh + (long long)row * H + ch * 8), hi = ldg4(wh + (long long)row * H + ch * 8 + 4);
        const int m = row >> 7, r = row & 127;
        *reinterpret_cast<uint4 *>(s.a + m * 2 * LT_ATOM_A + (ch >> 3) * LT_ATOM_A + r * 128 + (((ch & 7) ^ (r & 7)) << 4)) =
            make_uint4(lt_pack2(lo.x, lo.y), lt_pack2(lo.z, lo.w), lt_pack2(hi.x, hi.y), lt_pack2(hi.z, hi.w));
    }
    for (int i = threadIdx.x; i < 2 * LT_ATOM_B / 16; i += LT_THREADS) reinterpret_cast<uint4 *>(s.b)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) { mbar_init(s.done, LT_ISSUERS); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(s.tmem_ptr, 64); tmem_relinquish(); }
    __syncthreads();
    if (h0) {
        for (int i = threadIdx.x; i < LT_ROWS * H; i += LT_THREADS) {
            const int r = i >> 7, k = i & 127;
            if (row0 + r < N) lt_store_bf16(s.b, r, k, __ldg(h0 + (long long)(row0 + r) * H + k));
        }
    }
    float c[LT_CPT];
    bool live[LT_CPT];
#pragma unroll
    for (int i = 0; i < LT_CPT; ++i) {
        const int row = row0 + LT_CPT * half + i;
        live[i] = row < N;
        c[i] = (c0 && live[i]) ? __ldg(c0 + (long long)row * H + unit) : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s.tmem_ptr;
    constexpr uint32_t idesc = make_idesc(1, 128, LT_N, 0, 0);
    // descriptors = template | (start address >> 4): one add per operand in the issuing thread.  (ncu: the tensor pipe is
    // active ~270 cycles of a ~2250-cycle step while the cell warps wait ~1280 cycles for the commit: what a step pays
    // for its 32 MMAs is their ISSUE, ~35 cycles per tcgen05.mma whatever its N, not their execution.)
    const uint32_t a16 = smem_u32(s.a) >> 4, b16 = smem_u32(s.b) >> 4;
    const uint64_t tmpl = make_smem_desc(0, 0, 1024);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(LT_CPT * half);

    // The per-step operands are loaded ONE STEP AHEAD (registers): an HBM round trip is about as long as the MMAs of
    // a step, so loads issued at the top of the step they belong to were the tail of its critical path.
    auto load_zin = [&](int t, float (&z4)[LT_CPT][4]) {
#pragma unroll
        for (int i = 0; i < LT_CPT; ++i) {
            const float *z = zx + ((long long)t * N + row0 + LT_CPT * half + i) * 4 * H + unit;
#pragma unroll
            for (int g = 0; g < 4; ++g) z4[i][g] = live[i] ? __ldg(z + g * H) : 0.f;
        }
    };
    float zin[LT_CPT][4], znext[LT_CPT][4];
    load_zin(0, zin);
    for (int t = 0; t < T; ++t) {
        if (t + 1 < T) load_zin(t + 1, znext);
        if (warp < LT_ISSUERS) {
            // LT_ISSUERS warps (on different SM sub-partitions) issue the step's MMAs: warp w takes the gate tiles
            // m = w, w + LT_ISSUERS, ...; each commits to the same mbarrier (count LT_ISSUERS)
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < H / 16; ++ks)          // the gate tiles in turn: consecutive MMAs are independent
#pragma unroll
                    for (int mm = 0; mm < 4 / LT_ISSUERS; ++mm) {
                        const int m = warp + mm * LT_ISSUERS;
                        mma_bf16(tmem_base + m * LT_N, tmpl | (uint64_t)(a16 + ((m * 2 * LT_ATOM_A + (ks >> 2) * LT_ATOM_A + (ks & 3) * 32) >> 4)),
                                 tmpl | (uint64_t)(b16 + (((ks >> 2) * LT_ATOM_B + (ks & 3) * 32) >> 4)), idesc, ks ? 1u : 0u);
                    }
                mma_commit(s.done);
            }
            __syncwarp();
        }
        mbar_wait(s.done, (uint32_t)(t & 1));
        tc_fence_after();
        uint32_t zr[4][LT_CPT];                 // [gate][row]
#pragma unroll
        for (int g = 0; g < 4; ++g) lt_tmem_ld(tlane + g * LT_N, zr[g]);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int i = 0; i < LT_CPT; ++i) {
            const float f = lt_sigmoid(__uint_as_float(zr[0][i]) + zin[i][0]);
            const float ug = lt_sigmoid(__uint_as_float(zr[1][i]) + zin[i][1]);
            const float ct = lt_tanh(__uint_as_float(zr[2][i]) + zin[i][2]);
            const float o = lt_sigmoid(__uint_as_float(zr[3][i]) + zin[i][3]);
            c[i] = fmaf(f, c[i], ug * ct);                                           // layers/Recurrent.py:300-302
            const float h = o * lt_tanh(c[i]);
            if (live[i]) {
                lt_store_bf16(s.b, LT_CPT * half + i, unit, h);
                const int row = row0 + LT_CPT * half + i;
                float *a = acts + ((long long)t * N + row) * 4 * H + unit;
                a[0] = f; a[H] = ug; a[2 * H] = ct; a[3 * H] = o;
                const long long at = ((long long)(t + 1) * N + row) * H + unit;
                c_all[at] = c[i];
                h_all[at] = h;
                if (h_seq) h_seq[((long long)row * T + t) * H + unit] = h;
            }
        }
        fence_proxy_async();                   // the B tile was written through the generic proxy
        __syncthreads();                       // ... by every cell thread; the accumulators have been read
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < LT_CPT; ++i)
#pragma unroll
            for (int g = 0; g < 4; ++g) zin[i][g] = znext[i][g];
    }
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------ backward
// dz (T, N, 4H): gate pre-activation gradients in FORWARD column order (f, u, c~, o) — see recurrent.cu
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_seq_bwd_tc_kernel(const float *__restrict__ dh_last, const float *__restrict__ dh_seq, const float *__restrict__ wh,
                       const float *__restrict__ acts, const float *__restrict__ c_all, int c0_is_zero, float *__restrict__ dz,
                       int T, int N) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int H = LT_H, K = 4 * LT_H;
    const LtSmem s = lt_carve<8>(smem_raw);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int unit = q * 32 + lane;
    const int row0 = blockIdx.x * LT_ROWS;

    // A = Wh^T: row = unit j, k over the 4H gate columns: A[j][k] = wh[k][j]; a thread packs 8 consecutive k of one j
    for (int i = threadIdx.x; i < (K / 8) * H; i += LT_THREADS) {
        const int j = i & (H - 1), kc = i >> 7;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __ldg(wh + (long long)(kc * 8 + e) * H + j);
        *reinterpret_cast<uint4 *>(s.a + (kc >> 3) * LT_ATOM_A + j * 128 + (((kc & 7) ^ (j & 7)) << 4)) =
            make_uint4(lt_pack2(v[0], v[1]), lt_pack2(v[2], v[3]), lt_pack2(v[4], v[5]), lt_pack2(v[6], v[7]));
    }
    for (int i = threadIdx.x; i < 8 * LT_ATOM_B / 16; i += LT_THREADS) reinterpret_cast<uint4 *>(s.b)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) { mbar_init(s.done, LT_ISSUERS); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(s.tmem_ptr, 64); tmem_relinquish(); }
    bool live[LT_CPT];
    float dc[LT_CPT];
#pragma unroll
    for (int i = 0; i < LT_CPT; ++i) { live[i] = row0 + LT_CPT * half + i < N; dc[i] = 0.f; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s.tmem_ptr;
    constexpr uint32_t idesc = make_idesc(1, 128, LT_N, 0, 0);
    // descriptors = template | (start address >> 4): one add per operand in the issuing thread.  (ncu: the tensor pipe is
    // active ~270 cycles of a ~2250-cycle step while the cell warps wait ~1280 cycles for the commit: what a step pays
    // for its 32 MMAs is their ISSUE, ~35 cycles per tcgen05.mma whatever its N, not their execution.)
    const uint32_t a16 = smem_u32(s.a) >> 4, b16 = smem_u32(s.b) >> 4;
    const uint64_t tmpl = make_smem_desc(0, 0, 1024);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(LT_CPT * half);
    uint32_t phase = 0;

    struct Cell { float f, ug, ct, o, cv, cp, dh; };
    auto load_cells = [&](int t, Cell (&cl)[LT_CPT]) {                    // (one step ahead, see the forward kernel)
#pragma unroll
        for (int i = 0; i < LT_CPT; ++i) {
            cl[i] = Cell{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (live[i]) {
                const int row = row0 + LT_CPT * half + i;
                const float *a = acts + ((long long)t * N + row) * K + unit;
                cl[i].f = __ldg(a); cl[i].ug = __ldg(a + H); cl[i].ct = __ldg(a + 2 * H); cl[i].o = __ldg(a + 3 * H);
                cl[i].cv = __ldg(c_all + ((long long)(t + 1) * N + row) * H + unit);
                cl[i].cp = (t == 0 && c0_is_zero) ? 0.f : __ldg(c_all + ((long long)t * N + row) * H + unit);
                if (dh_seq) cl[i].dh = __ldg(dh_seq + ((long long)row * T + t) * H + unit);
                else if (t == T - 1 && dh_last) cl[i].dh = __ldg(dh_last + (long long)row * H + unit);
            }
        }
    };
    Cell cur[LT_CPT], nxt[LT_CPT];
    load_cells(T - 1, cur);
    for (int t = T - 1; t >= 0; --t) {
        if (t > 0) load_cells(t - 1, nxt);
        float f[LT_CPT], ug[LT_CPT], ct[LT_CPT], o[LT_CPT], cv[LT_CPT], cp[LT_CPT], dh[LT_CPT];
#pragma unroll
        for (int i = 0; i < LT_CPT; ++i) {
            f[i] = cur[i].f; ug[i] = cur[i].ug; ct[i] = cur[i].ct; o[i] = cur[i].o; cv[i] = cur[i].cv; cp[i] = cur[i].cp; dh[i] = cur[i].dh;
        }
        if (t != T - 1) {
            if (warp < LT_ISSUERS) {
                if (elect_one()) {
                    // the 32 K-steps go to FOUR accumulators in turn (one chain of 32 dependent MMAs paid the MMA
                    // latency, not its issue rate, 32 times); the cell thread adds the four partial sums.
                    // Issuer w takes the accumulators w, w + LT_ISSUERS, ...
#pragma unroll
                    for (int ks = 0; ks < K / 16; ++ks)
                        if ((ks & 3) % LT_ISSUERS == warp)
                            mma_bf16(tmem_base + (ks & 3) * LT_N, tmpl | (uint64_t)(a16 + (((ks >> 2) * LT_ATOM_A + (ks & 3) * 32) >> 4)),
                                     tmpl | (uint64_t)(b16 + (((ks >> 2) * LT_ATOM_B + (ks & 3) * 32) >> 4)), idesc, ks >= 4 ? 1u : 0u);
                    mma_commit(s.done);
                }
                __syncwarp();
            }
            mbar_wait(s.done, phase);
            phase ^= 1;
            tc_fence_after();
            uint32_t r[4][LT_CPT];
#pragma unroll
            for (int a = 0; a < 4; ++a) lt_tmem_ld(tlane + a * LT_N, r[a]);
            tmem_ld_wait();
            tc_fence_before();
#pragma unroll
            for (int i = 0; i < LT_CPT; ++i)
                dh[i] += (__uint_as_float(r[0][i]) + __uint_as_float(r[1][i])) + (__uint_as_float(r[2][i]) + __uint_as_float(r[3][i]));
        }
#pragma unroll
        for (int i = 0; i < LT_CPT; ++i) {
            const float tc = lt_tanh(cv[i]);
            const float dcv = dc[i] + dh[i] * o[i] * (1.f - tc * tc);                  // layers/Recurrent.py:337-338
            float d[4];
            d[0] = dcv * cp[i] * f[i] * (1.f - f[i]);                                  // forget gate   (:351-352)
            d[1] = dcv * ct[i] * ug[i] * (1.f - ug[i]);                                // update gate   (:345-346)
            d[2] = dcv * ug[i] * (1.f - ct[i] * ct[i]);                                // candidate     (:340-341)
            d[3] = dh[i] * tc * o[i] * (1.f - o[i]);                                   // output gate   (:329-330)
            dc[i] = dcv * f[i];                                                        // :366
            if (live[i]) {
                float *out = dz + ((long long)t * N + row0 + LT_CPT * half + i) * K + unit;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    out[g * H] = d[g];
                    if (t > 0) lt_store_bf16(s.b, LT_CPT * half + i, g * H + unit, d[g]);
                }
            }
        }
        fence_proxy_async();
        __syncthreads();
        tc_fence_after();
#pragma unroll
        for (int i = 0; i < LT_CPT; ++i) cur[i] = nxt[i];
    }
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

template <typename Kern>
int lt_configure(Kern kernel, int smem, int which) {
    static bool configured_dev[16][2] = {};
    bool &configured = configured_dev[current_device_slot()][which];
    if (!configured) {
        SN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    return SN_OK;
}

}  // namespace

extern "C" {

int sn_lstm_seq_tc_supported(int N, int H) {
    const char *e = getenv("SN_LSTM_TC");
    if (e && e[0] == '0') return 0;
    return (N > 0 && H == LT_H) ? 1 : 0;
}

int sn_lstm_seq_fwd_tc(const float *zx, const float *wh, const float *h0, const float *c0, float *acts, float *c_all, float *h_all,
                       float *h_seq, int T, int N, int H, void *stream) {
    SN_REQUIRE(zx && wh && acts && c_all && h_all, "null tensor");
    SN_REQUIRE(T > 0 && N > 0 && H == LT_H, "the tensor-pipe LSTM kernels take H = 128");
    SN_REQUIRE((((uintptr_t)wh) & 15) == 0, "wh must be 16-byte aligned");
    constexpr int smem = lt_smem_bytes<2>();
    int e = lt_configure(lstm_seq_fwd_tc_kernel, smem, 0);
    if (e) return e;
    lstm_seq_fwd_tc_kernel<<<(N + LT_ROWS - 1) / LT_ROWS, LT_THREADS, smem, as_stream(stream)>>>(zx, wh, h0, c0, acts, c_all, h_all, h_seq, T, N);
    return after_launch("lstm_seq_fwd_tc_kernel");
}

int sn_lstm_seq_bwd_tc(const float *dh_last, const float *dh_seq, const float *wh, const float *acts, const float *c_all, int c0_is_zero,
                       float *dz, int T, int N, int H, void *stream) {
    SN_REQUIRE(wh && acts && c_all && dz, "null tensor");
    SN_REQUIRE(dh_last || dh_seq, "no upstream gradient");
    SN_REQUIRE(T > 0 && N > 0 && H == LT_H, "the tensor-pipe LSTM kernels take H = 128");
    constexpr int smem = lt_smem_bytes<8>();
    int e = lt_configure(lstm_seq_bwd_tc_kernel, smem, 1);
    if (e) return e;
    lstm_seq_bwd_tc_kernel<<<(N + LT_ROWS - 1) / LT_ROWS, LT_THREADS, smem, as_stream(stream)>>>(dh_last, dh_seq, wh, acts, c_all, c0_is_zero, dz, T, N);
    return after_launch("lstm_seq_bwd_tc_kernel");
}

}  // extern "C"
