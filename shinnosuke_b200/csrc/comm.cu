// Data-parallel exchange over NVLink 5 / NVSwitch peer memory, hand-written (no NCCL on the data path).
//
// The reference's training step (models.py:70-92) is forward sweep -> loss.backward -> backward sweep
// -> optimizer.update over ONE minibatch.  Sharded over G GPUs by rows, two things must be summed over
// the replicas to reproduce it exactly: the gradient arena (once per step, 3.9 MB for config 3) and
// the per-channel sums of every BatchNormalization (layers/Normalization.py:116-128 normalises over
// the WHOLE minibatch; 8 reductions of 8-32 KB of float64 per step for config 3).  Both are far
// below the size where link bandwidth matters: they are latency problems, so each is ONE kernel
// that moves the data with plain loads / stores to peer-mapped memory and synchronises with flags
// that live in the same peer-mapped memory — no host round trip, no multi-kernel protocol, safe to
// capture in a CUDA graph (the sequence number is device state).
//
// Memory: every rank owns one "window" (cudaMalloc) = [ user region | staging slots | flags ].
// Windows are mapped into every peer: cudaIpcOpenMemHandle across processes (one process per GPU,
// sn_comm_create + sn_comm_ipc_handle + sn_comm_connect) or cudaDeviceEnablePeerAccess inside one
// process that drives all GPUs (sn_comm_init_all).
//
//  * in-place all-reduce of a tensor that LIVES in the user region (the gradient arena is allocated
//    there): two-shot.  Rank r owns slice r: it pulls slice r from every peer's window (P2P loads),
//    sums in rank order 0..G-1, and pushes the result into slice r of every peer's window (P2P
//    stores).  Entry and exit barriers are per-CTA flag exchanges (release / acquire at system
//    scope).  Every element is reduced by exactly one rank in a fixed order, so all replicas hold
//    bit-identical results.  Traffic per rank: (G-1)/G of the tensor in, the same out.
//  * all-reduce of a small tensor anywhere in local memory (BatchNorm sums, float64): one-shot.
//    Every rank pushes its values into its own slot of every peer's staging area, flags it, waits
//    for the G slots of its own window and sums them in rank order (bit-identical again).  Slots
//    are double-buffered by the parity of the sequence number: a rank can be at most one exchange
//    ahead of the slowest peer, because finishing exchange k needs every peer's flag of exchange k.
//
// Flags are monotonically increasing sequence numbers (never reset), one word per (phase, source
// rank, CTA); the sequence counter is advanced by the last CTA of each exchange kernel to finish.
#include "common.cuh"
#include <new>
#include <string.h>

using namespace sn;

namespace {

constexpr int MAX_RANKS = 8;
constexpr int MAX_CTAS = 64;                 // CTAs of an exchange kernel (flag slots per phase and rank)
constexpr int CTA_THREADS = 512;
constexpr size_t STAGE_SLOT_BYTES = 256 * 1024;      // one rank's slot of the one-shot staging area (per parity)
constexpr size_t FLAG_WORDS = 3 * MAX_RANKS * MAX_CTAS;    // phases: 0 = two-shot entry, 1 = two-shot exit, 2 = one-shot
constexpr size_t ALIGN = 1024;

struct Window {            // layout of one rank's window, identical on every rank
    size_t user_bytes, stage_off, flag_off, total;
};
Window window_layout(size_t user_bytes) {
    Window w;
    w.user_bytes = (user_bytes + ALIGN - 1) / ALIGN * ALIGN;
    w.stage_off = w.user_bytes;
    w.flag_off = w.stage_off + 2 * MAX_RANKS * STAGE_SLOT_BYTES;
    w.total = w.flag_off + (FLAG_WORDS * sizeof(unsigned) + ALIGN - 1) / ALIGN * ALIGN;
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
__device__ __forceinline__ void cta_exchange_flags(const DevTable &t, size_t flag_off, int phase, int rank, int world, unsigned seq) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        st_release_sys(flag_ptr(t.win[peer], flag_off, phase, rank, blockIdx.x), seq);
        const unsigned *mine = flag_ptr(t.win[rank], flag_off, phase, peer, blockIdx.x);
        while ((int)(ld_acquire_sys(mine) - seq) < 0) { }
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

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };
__device__ __forceinline__ void vadd(float4 &a, const float4 &b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void vadd(double2 &a, const double2 &b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ float4 ld_cg(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ double2 ld_cg(const double2 *p) { return __ldcg(p); }

// ---- two-shot, in place, tensor at byte offset `off` of every window's user region --------------
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS) allreduce_inplace_kernel(DevTable t, size_t flag_off, unsigned *state, size_t off,
                                                                        long long n, int rank, int world) {
    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    const unsigned seq = *reinterpret_cast<volatile unsigned *>(&state[0]) + 1u;
    // entry: every rank's backward pass has finished writing its gradients (stream order on each rank)
    cta_exchange_flags(t, flag_off, 0, rank, world, seq);
    // slice of this rank, in vectors; the ragged tail (n % VN elements) belongs to the last rank
    const long long nv = n / VN;
    const long long per = (nv + world - 1) / world;
    const long long v0 = per * rank, v1 = min(nv, v0 + per);
    for (long long i = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += (long long)gridDim.x * blockDim.x) {
        V acc = ld_cg(reinterpret_cast<const V *>(t.win[0] + off) + i);
#pragma unroll
        for (int p = 1; p < MAX_RANKS; ++p)
            if (p < world) vadd(acc, ld_cg(reinterpret_cast<const V *>(t.win[p] + off) + i));
#pragma unroll
        for (int p = 0; p < MAX_RANKS; ++p)
            if (p < world) reinterpret_cast<V *>(t.win[p] + off)[i] = acc;
    }
    if (rank == world - 1 && blockIdx.x == 0) {
        for (long long i = nv * VN + threadIdx.x; i < n; i += blockDim.x) {
            T acc = __ldcg(reinterpret_cast<const T *>(t.win[0] + off) + i);
            for (int p = 1; p < world; ++p) acc += __ldcg(reinterpret_cast<const T *>(t.win[p] + off) + i);
            for (int p = 0; p < world; ++p) reinterpret_cast<T *>(t.win[p] + off)[i] = acc;
        }
    }
    // exit: every rank has pushed its slice into every window
    cta_exchange_flags(t, flag_off, 1, rank, world, seq);
    finish_exchange(state, seq);
}

// ---- one-shot, any local buffer of at most STAGE_SLOT_BYTES -------------------------------------
template <typename T>
__global__ void __launch_bounds__(CTA_THREADS) allreduce_oneshot_kernel(DevTable t, size_t stage_off, size_t flag_off, unsigned *state,
                                                                        T *buf, long long n, int rank, int world) {
    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    const unsigned seq = *reinterpret_cast<volatile unsigned *>(&state[0]) + 1u;
    const size_t slot0 = stage_off + (size_t)(seq & 1u) * MAX_RANKS * STAGE_SLOT_BYTES;
    const long long nv = n / VN;
    // this CTA's chunk of vectors (the same partition on every rank)
    const long long per = (nv + gridDim.x - 1) / gridDim.x;
    const long long v0 = per * blockIdx.x, v1 = min(nv, v0 + per);
    const bool vec_ok = (reinterpret_cast<uintptr_t>(buf) & (sizeof(V) - 1)) == 0;
    if (vec_ok) {
        for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) {
            const V v = reinterpret_cast<const V *>(buf)[i];
#pragma unroll
            for (int p = 0; p < MAX_RANKS; ++p)
                if (p < world) reinterpret_cast<V *>(t.win[p] + slot0 + (size_t)rank * STAGE_SLOT_BYTES)[i] = v;
        }
    } else {
        for (long long i = v0 * VN + threadIdx.x; i < v1 * VN; i += blockDim.x) {
            const T v = buf[i];
            for (int p = 0; p < world; ++p) reinterpret_cast<T *>(t.win[p] + slot0 + (size_t)rank * STAGE_SLOT_BYTES)[i] = v;
        }
    }
    if (blockIdx.x == 0)
        for (long long i = nv * VN + threadIdx.x; i < n; i += blockDim.x) {
            const T v = buf[i];
            for (int p = 0; p < world; ++p) reinterpret_cast<T *>(t.win[p] + slot0 + (size_t)rank * STAGE_SLOT_BYTES)[i] = v;
        }
    cta_exchange_flags(t, flag_off, 2, rank, world, seq);
    const char *mine = t.win[rank] + slot0;
    if (vec_ok) {
        for (long long i = v0 + threadIdx.x; i < v1; i += blockDim.x) {
            V acc = ld_cg(reinterpret_cast<const V *>(mine) + i);
#pragma unroll
            for (int p = 1; p < MAX_RANKS; ++p)
                if (p < world) vadd(acc, ld_cg(reinterpret_cast<const V *>(mine + (size_t)p * STAGE_SLOT_BYTES) + i));
            reinterpret_cast<V *>(buf)[i] = acc;
        }
    } else {
        for (long long i = v0 * VN + threadIdx.x; i < v1 * VN; i += blockDim.x) {
            T acc = __ldcg(reinterpret_cast<const T *>(mine) + i);
            for (int p = 1; p < world; ++p) acc += __ldcg(reinterpret_cast<const T *>(mine + (size_t)p * STAGE_SLOT_BYTES) + i);
            buf[i] = acc;
        }
    }
    if (blockIdx.x == 0)
        for (long long i = nv * VN + threadIdx.x; i < n; i += blockDim.x) {
            T acc = __ldcg(reinterpret_cast<const T *>(mine) + i);
            for (int p = 1; p < world; ++p) acc += __ldcg(reinterpret_cast<const T *>(mine + (size_t)p * STAGE_SLOT_BYTES) + i);
            buf[i] = acc;
        }
    finish_exchange(state, seq);
}

// local copies around a two-shot exchange of a large tensor that does not live in the window
template <typename T>
__global__ void copy_kernel(const T *__restrict__ src, T *__restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

int grid_two_shot(long long bytes, int world) {
    // latency-bound: enough CTAs to keep the NVLink ports busy (a CTA moves 512 x 16 B per trip), never
    // so many that the flag exchange (world remote stores + world polls per CTA) dominates
    long long per_rank = bytes / world;
    int g = (int)((per_rank + 32 * 1024 - 1) / (32 * 1024));
    return g < 1 ? 1 : (g > MAX_CTAS ? MAX_CTAS : g);
}

template <typename T>
int allreduce(Comm *c, T *buf, long long n, cudaStream_t st) {
    if (!c->connected) { set_error("sn_allreduce: the communicator is not connected (sn_comm_connect / sn_comm_init_all)"); return SN_ERR_INVALID_ARG; }
    if (n <= 0 || c->world == 1) return SN_OK;
    DevTable t;
    for (int p = 0; p < MAX_RANKS; ++p) t.win[p] = c->peer[p < c->world ? p : 0];
    const size_t bytes = (size_t)n * sizeof(T);
    char *b = reinterpret_cast<char *>(buf);
    const bool in_window = b >= c->local && b + bytes <= c->local + c->lay.user_bytes;
    if (in_window && bytes > 64 * 1024) {
        const size_t off = (size_t)(b - c->local);
        if (off % 16) { set_error("sn_allreduce: a tensor inside the window must be 16-byte aligned"); return SN_ERR_INVALID_ARG; }
        allreduce_inplace_kernel<T><<<grid_two_shot((long long)bytes, c->world), CTA_THREADS, 0, st>>>(t, c->lay.flag_off, c->state, off, n, c->rank, c->world);
        return after_launch("allreduce_inplace_kernel");
    }
    if (bytes <= STAGE_SLOT_BYTES) {
        int g = (int)((bytes + 16 * 1024 - 1) / (16 * 1024));
        g = g < 1 ? 1 : (g > MAX_CTAS ? MAX_CTAS : g);
        allreduce_oneshot_kernel<T><<<g, CTA_THREADS, 0, st>>>(t, c->lay.stage_off, c->lay.flag_off, c->state, buf, n, c->rank, c->world);
        return after_launch("allreduce_oneshot_kernel");
    }
    // large and outside the window: through the head of the user region, in pieces
    const size_t cap = c->lay.user_bytes / sizeof(T);
    if (cap == 0) { set_error("sn_allreduce: %zu bytes do not fit the staging slots and the communicator has no user region", bytes); return SN_ERR_INVALID_ARG; }
    for (long long done = 0; done < n; done += (long long)cap) {
        const long long m = (n - done) < (long long)cap ? (n - done) : (long long)cap;
        copy_kernel<T><<<grid_for(m, 256), 256, 0, st>>>(buf + done, reinterpret_cast<T *>(c->local), m);
        ++g_launch_count;
        allreduce_inplace_kernel<T><<<grid_two_shot(m * (long long)sizeof(T), c->world), CTA_THREADS, 0, st>>>(t, c->lay.flag_off, c->state, 0, m, c->rank, c->world);
        ++g_launch_count;
        copy_kernel<T><<<grid_for(m, 256), 256, 0, st>>>(reinterpret_cast<const T *>(c->local), buf + done, m);
        int e = after_launch("allreduce (staged)");
        if (e) return e;
    }
    return SN_OK;
}

int comm_alloc(Comm *c, size_t user_bytes) {
    c->lay = window_layout(user_bytes);
    SN_CUDA(cudaMalloc(&c->local, c->lay.total));
    SN_CUDA(cudaMemset(c->local, 0, c->lay.total));
    SN_CUDA(cudaMalloc(&c->state, 2 * sizeof(unsigned)));
    SN_CUDA(cudaMemset(c->state, 0, 2 * sizeof(unsigned)));
    SN_CUDA(cudaDeviceSynchronize());
    return SN_OK;
}

}  // namespace

extern "C" {

int sn_comm_create(int rank, int world, size_t user_bytes, void **comm) {
    SN_REQUIRE(comm && world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, "bad rank / world (at most 8 ranks)");
    Comm *c = new (std::nothrow) Comm();
    SN_REQUIRE(c, "out of host memory");
    memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world;
    SN_CUDA(cudaGetDevice(&c->device));
    int e = comm_alloc(c, user_bytes);
    if (e) { delete c; return e; }
    c->peer[rank] = c->local;
    c->connected = (world == 1);
    *comm = c;
    return SN_OK;
}

int sn_comm_window(void *comm, void **ptr, size_t *bytes) {
    SN_REQUIRE(comm && ptr, "bad arguments");
    Comm *c = static_cast<Comm *>(comm);
    *ptr = c->local;
    if (bytes) *bytes = c->lay.user_bytes;
    return SN_OK;
}

int sn_comm_ipc_handle(void *comm, void *handle64) {
    SN_REQUIRE(comm && handle64, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
    Comm *c = static_cast<Comm *>(comm);
    cudaIpcMemHandle_t h;
    SN_CUDA(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle64, &h, 64);
    return SN_OK;
}

int sn_comm_connect(void *comm, const void *handles) {
    SN_REQUIRE(comm && handles, "bad arguments");
    Comm *c = static_cast<Comm *>(comm);
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank || c->peer[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + 64 * p, 64);
        void *ptr = nullptr;
        SN_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[p] = static_cast<char *>(ptr);
        c->ipc_opened[p] = true;
    }
    c->connected = true;
    return SN_OK;
}

int sn_comm_init_all(int ndev, size_t user_bytes, void **comms) {
    SN_REQUIRE(comms && ndev >= 1 && ndev <= MAX_RANKS, "bad arguments (at most 8 devices)");
    int have = 0, prev = 0;
    SN_CUDA(cudaGetDeviceCount(&have));
    SN_REQUIRE(ndev <= have, "more ranks than visible devices");
    SN_CUDA(cudaGetDevice(&prev));
    Comm *cs[MAX_RANKS] = {};
    for (int r = 0; r < ndev; ++r) {
        SN_CUDA(cudaSetDevice(r));
        void *h = nullptr;
        int e = sn_comm_create(r, ndev, user_bytes, &h);
        if (e) { cudaSetDevice(prev); return e; }
        cs[r] = static_cast<Comm *>(h);
    }
    for (int r = 0; r < ndev; ++r) {
        SN_CUDA(cudaSetDevice(r));
        for (int p = 0; p < ndev; ++p) {
            if (p == r) continue;
            int can = 0;
            SN_CUDA(cudaDeviceCanAccessPeer(&can, r, p));
            if (!can) { cudaSetDevice(prev); set_error("sn_comm_init_all: device %d cannot access device %d", r, p); return SN_ERR_UNSUPPORTED; }
            cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else SN_CUDA(e);
            cs[r]->peer[p] = cs[p]->local;
        }
        cs[r]->connected = true;
        comms[r] = cs[r];
    }
    SN_CUDA(cudaSetDevice(prev));
    return SN_OK;
}

int sn_comm_destroy(void *comm) {
    if (!comm) return SN_OK;
    Comm *c = static_cast<Comm *>(comm);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
    for (int p = 0; p < c->world; ++p)
        if (c->ipc_opened[p]) cudaIpcCloseMemHandle(c->peer[p]);
    cudaFree(c->local);
    cudaFree(c->state);
    cudaSetDevice(prev);
    cudaGetLastError();
    delete c;
    return SN_OK;
}

int sn_allreduce_sum_f32(void *comm, float *buf, long long n, void *stream) {
    SN_REQUIRE(comm && (buf || n == 0), "bad arguments");
    return allreduce<float>(static_cast<Comm *>(comm), buf, n, as_stream(stream));
}
int sn_allreduce_sum_f64(void *comm, double *buf, long long n, void *stream) {
    SN_REQUIRE(comm && (buf || n == 0), "bad arguments");
    return allreduce<double>(static_cast<Comm *>(comm), buf, n, as_stream(stream));
}

// ---- CUDA graphs: the training step is captured once per batch shape and replayed ----------------
int sn_graph_begin(void *stream) {
    SN_CUDA(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeThreadLocal));
    return SN_OK;
}
int sn_graph_end(void *stream, void **graph_exec) {
    SN_REQUIRE(graph_exec, "null out pointer");
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(as_stream(stream), &g);
    if (e != cudaSuccess) { cudaGetLastError(); return check_cuda(e, "cudaStreamEndCapture"); }
    cudaGraphExec_t x = nullptr;
    e = cudaGraphInstantiate(&x, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { cudaGetLastError(); return check_cuda(e, "cudaGraphInstantiate"); }
    *graph_exec = x;
    return SN_OK;
}
int sn_graph_launch(void *graph_exec, void *stream) {
    SN_REQUIRE(graph_exec, "null graph");
    SN_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), as_stream(stream)));
    return SN_OK;
}
int sn_graph_destroy(void *graph_exec) {
    if (graph_exec) SN_CUDA(cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(graph_exec)));
    return SN_OK;
}
int sn_stream_wait_event(void *stream, void *event) {
    SN_CUDA(cudaStreamWaitEvent(as_stream(stream), (cudaEvent_t)event, 0));
    return SN_OK;
}
int sn_device_sync(void) { SN_CUDA(cudaDeviceSynchronize()); return SN_OK; }

}  // extern "C"
