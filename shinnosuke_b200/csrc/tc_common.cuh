// Blackwell (sm_100a) building blocks: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld) and the UMMA descriptors, as thin inline-PTX wrappers.  Bit layouts follow the
// PTX ISA tables for the tcgen05 shared-memory descriptor and instruction descriptor.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}\n" ::"r"(addr), "r"(parity)
        : "memory");
}

// long waits (an epilogue waiting for a whole main loop): back off between polls
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(256);
    }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// bulk tensor STORE shared -> global (clipped at the tensor bounds), tracked by bulk async-groups
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the committed stores have finished READING shared memory (the buffer may be reused)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], fp32 operands read as TF32, single CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2)
// Two CTAs of a cluster (same TPC) run ONE tcgen05.mma of M = 256: each holds its own 128 rows of A,
// HALF of the B tile and the 128 x N accumulator of its rows.  The leader (cluster rank 0) issues the
// MMAs; both CTAs' TMA loads report to the leader's "full" barrier; tcgen05.commit multicasts its
// arrival to the barrier at the same shared-memory offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra.uni WAIT_DONE_C;\n\t"
        "bra.uni WAIT_LOOP_C;\n\t"
        "WAIT_DONE_C:\n\t}\n" ::"r"(addr), "r"(parity)
        : "memory");
}
// TMA loads of a CTA pair: data into THIS CTA's shared memory, completion bytes to `bar_cluster`
// (a shared::cluster address, normally the leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *m, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void *dst, const CUtensorMap *m, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void mma_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of the warp gets row (lane base + lane).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, SWIZZLE_128B.  Addresses / offsets are in 16-byte units.
//   K-major  (rows of 128 B along K): SBO = stride between 8-row groups (1024 B when dense), LBO unused.
//   MN-major (rows of 128 B along M/N, 8 K-rows per atom): LBO = stride between 128-byte-wide
//            atoms along M/N, SBO = stride between 8-row K groups.
//   fp32/tf32 MN-major operands must use the "128-byte swizzle with 32-byte atoms" (layout type 1,
//   TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): K groups are 4 rows of 128 B, SBO = stride between them.
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = LAYOUT_SW128) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}

// Instruction descriptor for kind::tf32 / kind::f16 with fp32 accumulation.
//   fmt: 0 = f16, 1 = bf16, 2 = tf32;  *_mn: 1 = operand is MN-major (transposed), 0 = K-major.
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace tc

// ---- host side: TMA tensor maps through the driver entry point (no link-time libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();
// fp32 tensor, rank <= 4, dims fastest-first, strides in ELEMENTS for dims 1..rank-1, 128-byte swizzle.
int make_tmap_f32(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
                  const int *box, bool atom32b = false);
// same for fp32 or bf16 elements (strides in elements)
int make_tmap_linear_f32(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
                         const int *box);
int make_tmap(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
              const int *box, bool atom32b, bool bf16);
// bf16 tensor with the 64-byte swizzle: the epilogue staging tiles of bf16 OUTPUT maps (32 rows x 32 bf16 = 64-byte rows)
int make_tmap_bf16_sw64(CUtensorMap *out, const void *base, int rank, const long long *dims, const long long *strides_elems,
                        const int *box);


// Column sums of one epilogue staging tile: 32 rows x 32 fp32 columns in the 128-byte-swizzled
// layout the TMA store reads (16-byte chunk j of row r at (j ^ (r & 7)) * 16).  Lane l takes the
// column pair 2*(l & 15) of rows 16*(l >> 4) .. +15 with 8-byte loads (each half warp reads one
// whole row: conflict-free) and packed f32x2 adds / fmas, one shuffle joins the two row halves,
// and lanes 0..15 add {sum, sum of squares} of their two columns to THIS WARP's tables in shared
// memory (plain read-modify-write: fp32 atomicAdd on shared memory is a CAS spin loop, ~20 us per
// launch when four warps share a table).  nvalid (< 32 only in the last pixel tile) masks rows past
// the end of the tensor.
__device__ __forceinline__ void staged_tile_colsums(const uint8_t *st, int lane, int nvalid, float *ssum, float *ssq, int ncols) {
    const int cp = lane & 15, r0 = (lane >> 4) * 16;
    const uint8_t *base = st + (cp & 1) * 8;
    const int chunk = cp >> 1;
    float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
    if (nvalid >= 32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int r = r0 + i;
            const float2 v = *reinterpret_cast<const float2 *>(base + r * 128 + ((chunk ^ (r & 7)) * 16));
            a = __fadd2_rn(a, v);
            b = __ffma2_rn(v, v, b);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int r = r0 + i;
            float2 v = *reinterpret_cast<const float2 *>(base + r * 128 + ((chunk ^ (r & 7)) * 16));
            if (r >= nvalid) v = make_float2(0.f, 0.f);
            a = __fadd2_rn(a, v);
            b = __ffma2_rn(v, v, b);
        }
    }
    a.x += __shfl_down_sync(0xffffffffu, a.x, 16);
    a.y += __shfl_down_sync(0xffffffffu, a.y, 16);
    b.x += __shfl_down_sync(0xffffffffu, b.x, 16);
    b.y += __shfl_down_sync(0xffffffffu, b.y, 16);
    if (lane < 16) {
        const int c = 2 * cp;
        if (c < ncols) { ssum[c] += a.x; ssq[c] += b.x; }
        if (c + 1 < ncols) { ssum[c + 1] += a.y; ssq[c + 1] += b.y; }
    }
}

// ---- bf16 output maps (bf16 tier: the convolution's own output is stored in bf16) -----------------------
// One epilogue warp's 32 rows x 32 columns as bf16: 64-byte rows in the TMA 64-byte swizzle (16-byte chunk j
// of row r at (j ^ ((r >> 1) & 3)) * 16; the tile is 512-byte aligned).  Lane = row writes its four chunks
// conflict-free (8 lanes of a 128-bit store phase cover 8 distinct 16-byte bank groups).
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void stage_row_bf16(uint8_t *st, int lane, const float (&v)[32]) {
#pragma unroll
    for (int jc = 0; jc < 4; ++jc) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * jc + 0], v[8 * jc + 1]); u.y = pack_bf16x2(v[8 * jc + 2], v[8 * jc + 3]);
        u.z = pack_bf16x2(v[8 * jc + 4], v[8 * jc + 5]); u.w = pack_bf16x2(v[8 * jc + 6], v[8 * jc + 7]);
        *reinterpret_cast<uint4 *>(st + lane * 64 + ((jc ^ ((lane >> 1) & 3)) * 16)) = u;
    }
}
// Column sums of such a tile (of the ROUNDED values: the statistics describe the map as stored).  Lane l
// takes the column pair (l & 15) of rows 2i + (l >> 4): each warp-wide 4-byte load covers two whole rows
// = 128 contiguous bytes, conflict-free.  Same per-warp tables as staged_tile_colsums.
__device__ __forceinline__ void staged_tile_colsums_bf16(const uint8_t *st, int lane, int nvalid, float *ssum, float *ssq, int ncols) {
    const int pr = lane & 15, rsel = lane >> 4;
    float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = 2 * i + rsel;
        const uint32_t w = *reinterpret_cast<const uint32_t *>(st + r * 64 + (((pr >> 2) ^ ((r >> 1) & 3)) * 16) + (pr & 3) * 4);
        float2 v = make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
        if (r >= nvalid) v = make_float2(0.f, 0.f);
        a = __fadd2_rn(a, v);
        b = __ffma2_rn(v, v, b);
    }
    a.x += __shfl_down_sync(0xffffffffu, a.x, 16);
    a.y += __shfl_down_sync(0xffffffffu, a.y, 16);
    b.x += __shfl_down_sync(0xffffffffu, b.x, 16);
    b.y += __shfl_down_sync(0xffffffffu, b.y, 16);
    if (lane < 16) {
        const int c = 2 * pr;
        if (c < ncols) { ssum[c] += a.x; ssq[c] += b.x; }
        if (c + 1 < ncols) { ssum[c + 1] += a.y; ssq[c + 1] += b.y; }
    }
}

}  // namespace sn
