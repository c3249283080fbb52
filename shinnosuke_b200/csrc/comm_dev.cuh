// Device-side pieces of the peer-memory exchange (csrc/comm.cu) that other kernels fuse into
// themselves: the window layout, the table of peer windows, the flag exchange.  See comm.cu for the
// protocol.
#pragma once
#include "common.cuh"

namespace sn {
namespace comm {

constexpr int MAX_RANKS = 8;
constexpr int MAX_CTAS = 128;                // CTAs of an exchange kernel (flag slots per phase and rank)
constexpr int CTA_THREADS = 512;
constexpr size_t STAGE_SLOT_BYTES = 256 * 1024;      // one rank's slot of the one-shot staging area (per parity)
constexpr size_t FLAG_WORDS = 3 * MAX_RANKS * MAX_CTAS;    // phases: 0 = two-shot entry, 1 = two-shot exit, 2 = one-shot
constexpr size_t ALIGN = 1024;
// "LL" region: small latency-critical exchanges (the BatchNorm sums) travel as 8-byte words {32 data bits,
// 32-bit sequence number}: the receiver polls the word itself, so there is no separate flag, no fence and
// no barrier — one one-way NVLink latency per exchange.  One slot per (parity, source rank).
constexpr size_t LL_SLOT_BYTES = 32 * 1024;              // 1024 channels x {sum_a, sum_b} x 2 halves x 8 bytes

struct Window {            // layout of one rank's window, identical on every rank
    size_t user_bytes, stage_off, flag_off, ll_off, total;
};
inline Window window_layout(size_t user_bytes) {
    Window w;
    w.user_bytes = (user_bytes + ALIGN - 1) / ALIGN * ALIGN;
    w.stage_off = w.user_bytes;
    w.flag_off = w.stage_off + 2 * MAX_RANKS * STAGE_SLOT_BYTES;
    w.ll_off = w.flag_off + (FLAG_WORDS * sizeof(unsigned) + ALIGN - 1) / ALIGN * ALIGN;
    w.total = w.ll_off + 2 * MAX_RANKS * LL_SLOT_BYTES;
    return w;
}

struct DevTable {          // kernel argument: every rank's window as seen from THIS rank
    char *win[MAX_RANKS];
};

struct Comm {
    int rank, world, device;
    Window lay;
    char *local;                       // this rank's window (cudaMalloc)
    char *peer[MAX_RANKS];             // peer[rank] == local
    bool ipc_opened[MAX_RANKS];
    unsigned *state;                   // device: [0] = sequence number of the last finished exchange, [1] = CTA arrival counter
    bool connected;
};

// A replica that never arrives (its host thread / process died) must not hang the GPUs of the others for ever:
// every wait gives up after SPIN_TIMEOUT_NS, counts the event (sn_comm_timeouts, checked by fit() at the end of
// every epoch) and lets the kernel finish with whatever it has.
constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct SpinGuard {
    unsigned *counter;             // state[2] of the communicator: give-ups since the last sn_comm_timeouts(reset)
    unsigned spins = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ explicit SpinGuard(unsigned *c) : counter(c) {}
    __device__ __forceinline__ bool expired() {
        if (++spins < 4096u) return false;
        spins = 0;
        const unsigned long long now = global_ns();
        if (!t0) { t0 = now; return false; }
        if (now - t0 < SPIN_TIMEOUT_NS) return false;
        if (counter) atomicAdd(counter, 1u);
        return true;
    }
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned *flag_ptr(char *win, size_t flag_off, int phase, int src_rank, int cta) {
    return reinterpret_cast<unsigned *>(win + flag_off) + ((size_t)phase * MAX_RANKS + src_rank) * MAX_CTAS + cta;
}

// all threads of the CTA: tell every rank's CTA `blockIdx.x` that this rank reached `phase` of exchange
// `seq`, then wait until every rank's CTA `blockIdx.x` has said the same.  Everything this CTA wrote
// before the call is visible to the peers after their wait (CTA barrier, then a system-scope release).
__device__ __forceinline__ void cta_exchange_flags(const DevTable &t, size_t flag_off, int phase, int rank, int world, unsigned seq, unsigned *state) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        // the release is cumulative: it covers the whole CTA's writes ordered before it by the barrier above
        // (a separate __threadfence_system() in front of it cost a second round trip to the farthest peer)
        st_release_sys(flag_ptr(t.win[peer], flag_off, phase, rank, blockIdx.x), seq);
        const unsigned *mine = flag_ptr(t.win[rank], flag_off, phase, peer, blockIdx.x);
        SpinGuard guard(state + 2);
        while ((int)(ld_acquire_sys(mine) - seq) < 0) { if (guard.expired()) break; }
    }
    __syncthreads();
}

// the last CTA of the kernel to get here publishes the new sequence number (every CTA has read the
// old one by then, because reading it is the first thing a CTA does)
__device__ __forceinline__ void finish_exchange(unsigned *state, unsigned seq) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&state[1], 1u) == gridDim.x - 1) {
            state[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned *>(&state[0]) = seq;
        }
    }
}


// ---- LL words -------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(void *dst, unsigned data, unsigned seq) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(data), "r"(seq) : "memory");
}
__device__ __forceinline__ unsigned ll_load(const void *src, unsigned seq, unsigned *state) {
    unsigned d, f;
    SpinGuard guard(state + 2);
    do {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(d), "=r"(f) : "l"(src) : "memory");
    } while (f != seq && !guard.expired());
    return d;
}
__device__ __forceinline__ void ll_store_f64x2(char *dst, double a, double b, unsigned seq) {
    const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
    ll_store(dst, (unsigned)ua, seq);
    ll_store(dst + 8, (unsigned)(ua >> 32), seq);
    ll_store(dst + 16, (unsigned)ub, seq);
    ll_store(dst + 24, (unsigned)(ub >> 32), seq);
}
__device__ __forceinline__ void ll_load_f64x2(const char *src, double &a, double &b, unsigned seq, unsigned *state) {
    const unsigned a0 = ll_load(src, seq, state), a1 = ll_load(src + 8, seq, state), b0 = ll_load(src + 16, seq, state), b1 = ll_load(src + 24, seq, state);
    a = __longlong_as_double((long long)(((unsigned long long)a1 << 32) | a0));
    b = __longlong_as_double((long long)(((unsigned long long)b1 << 32) | b0));
}

// kernel-side view of a communicator for kernels that embed an exchange (bn.cu)
struct DevComm {
    DevTable t;
    size_t stage_off, flag_off, ll_off;
    unsigned *state;
    int rank, world;
};
inline DevComm dev_comm(const Comm *c) {
    DevComm d;
    for (int p = 0; p < MAX_RANKS; ++p) d.t.win[p] = c->peer[p < c->world ? p : 0];
    d.stage_off = c->lay.stage_off; d.flag_off = c->lay.flag_off; d.ll_off = c->lay.ll_off; d.state = c->state; d.rank = c->rank; d.world = c->world;
    return d;
}

}  // namespace comm
}  // namespace sn
