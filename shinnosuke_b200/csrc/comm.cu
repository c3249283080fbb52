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
#include "comm_dev.cuh"
#include <new>
#include <string.h>

using namespace sn;
using namespace sn::comm;

namespace {

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };
__device__ __forceinline__ void vadd(float4 &a, const float4 &b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ void vadd(double2 &a, const double2 &b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ float4 ld_cg(const float4 *p) { return __ldcg(p); }
__device__ __forceinline__ double2 ld_cg(const double2 *p) { return __ldcg(p); }

// ---- two-shot, in place, tensor at byte offset `off` of every window's user region --------------
constexpr int INPLACE_THREADS = 256;        // 4 x world vectors in flight per thread need the registers of a half-size CTA
// SGD: the optimizer step rides on the exchange (float only).  After its exit barrier a CTA knows that the
// sums of exactly the elements its own index walks (in every rank's slice) have arrived in its window, so it
// applies w -= lr * g; g = 0 to those elements right there: the reference's optimizer.update
// (utils/Optimizers.py:29-32) without another launch and another pass over the arena.
template <typename T, bool SGD>
__global__ void __launch_bounds__(INPLACE_THREADS) allreduce_inplace_kernel(DevTable t, size_t flag_off, unsigned *state, size_t off,
                                                                        long long n, int rank, int world, float *w, long long n_update,
                                                                        float lr) {
    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::N;
    pdl_wait();
    const unsigned seq = *reinterpret_cast<volatile unsigned *>(&state[0]) + 1u;
    // entry: every rank's backward pass has finished writing its gradients (stream order on each rank)
    cta_exchange_flags(t, flag_off, 0, rank, world, seq, state);
    // slice of this rank, in vectors; the ragged tail (n % VN elements) belongs to the last rank
    const long long nv = n / VN;
    const long long per = (nv + world - 1) / world;
    const long long v0 = per * rank, v1 = min(nv, v0 + per);
    // four vectors per thread and trip, ALL loads (4 x world, most of them over NVLink, ~2-3 us each way)
    // issued before the first store: the compiler cannot hoist a load above a store through these
    // pointers, and a thread that walks its slice one round trip at a time is pure latency
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = v0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < v1; i0 += U * stride) {
        V acc[U], part[U][MAX_RANKS - 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < v1) {
                acc[u] = ld_cg(reinterpret_cast<const V *>(t.win[0] + off) + i);
#pragma unroll
                for (int p = 1; p < MAX_RANKS; ++p)
                    if (p < world) part[u][p - 1] = ld_cg(reinterpret_cast<const V *>(t.win[p] + off) + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < v1) {
#pragma unroll
                for (int p = 1; p < MAX_RANKS; ++p)
                    if (p < world) vadd(acc[u], part[u][p - 1]);
#pragma unroll
                for (int p = 0; p < MAX_RANKS; ++p)
                    if (p < world) reinterpret_cast<V *>(t.win[p] + off)[i] = acc[u];
            }
        }
    }
    if (rank == world - 1 && blockIdx.x == 0) {
        for (long long i = nv * VN + threadIdx.x; i < n; i += blockDim.x) {
            T acc = __ldcg(reinterpret_cast<const T *>(t.win[0] + off) + i);
            for (int p = 1; p < world; ++p) acc += __ldcg(reinterpret_cast<const T *>(t.win[p] + off) + i);
            for (int p = 0; p < world; ++p) reinterpret_cast<T *>(t.win[p] + off)[i] = acc;
        }
    }
    // exit: every rank has pushed its slice into every window
    cta_exchange_flags(t, flag_off, 1, rank, world, seq, state);
    if constexpr (SGD) {
        float4 *g4 = reinterpret_cast<float4 *>(t.win[rank] + off), *w4 = reinterpret_cast<float4 *>(w);
        const long long nu = n_update / 4;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < world; ++p) {
            const long long a = per * p, b = min(min(nv, a + per), nu);
            for (long long i = a + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += stride) {
                const float4 g = __ldcg(g4 + i);
                float4 wv = w4[i];
                wv.x -= lr * g.x; wv.y -= lr * g.y; wv.z -= lr * g.z; wv.w -= lr * g.w;
                w4[i] = wv;
                g4[i] = zero;
            }
        }
    }
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
    cta_exchange_flags(t, flag_off, 2, rank, world, seq, state);
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
    // latency-bound: enough CTAs to keep the NVLink ports busy (a CTA moves 256 x 4 x 16 B per trip), never
    // so many that the flag exchange (world remote stores + world polls per CTA) dominates
    long long per_rank = bytes / world;
    int g = (int)((per_rank + 16 * 1024 - 1) / (16 * 1024));
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
        SN_CUDA(launch_dependent(allreduce_inplace_kernel<T, false>, dim3(grid_two_shot((long long)bytes, c->world)), dim3(INPLACE_THREADS), 0, st,
                                 t, c->lay.flag_off, c->state, off, n, c->rank, c->world, (float *)nullptr, 0ll, 0.f));
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
        allreduce_inplace_kernel<T, false><<<grid_two_shot(m * (long long)sizeof(T), c->world), INPLACE_THREADS, 0, st>>>(t, c->lay.flag_off, c->state, 0, m, c->rank, c->world, nullptr, 0ll, 0.f);
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
    SN_CUDA(cudaMalloc(&c->state, 4 * sizeof(unsigned)));
    SN_CUDA(cudaMemset(c->state, 0, 4 * sizeof(unsigned)));
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

int sn_comm_timeouts(void *comm, long long *count, int reset) {
    SN_REQUIRE(comm && count, "null argument");
    Comm *c = static_cast<Comm *>(comm);
    unsigned v = 0;
    SN_CUDA(cudaMemcpy(&v, c->state + 2, sizeof(v), cudaMemcpyDeviceToHost));
    *count = (long long)v;
    if (reset && v) SN_CUDA(cudaMemset(c->state + 2, 0, sizeof(v)));
    return SN_OK;
}

int sn_allreduce_sum_f32(void *comm, float *buf, long long n, void *stream) {
    SN_REQUIRE(comm && (buf || n == 0), "bad arguments");
    return allreduce<float>(static_cast<Comm *>(comm), buf, n, as_stream(stream));
}
int sn_allreduce_sgd_update(void *comm, float *g, float *w, long long n, long long n_params, float lr, void *stream) {
    SN_REQUIRE(comm && g && w && n > 0 && n_params >= 0 && n_params <= n, "bad arguments");
    Comm *c = static_cast<Comm *>(comm);
    SN_REQUIRE(c->connected, "the communicator is not connected");
    char *b = reinterpret_cast<char *>(g);
    const size_t bytes = (size_t)n * 4;
    const bool in_window = b >= c->local && b + bytes <= c->local + c->lay.user_bytes;
    if (c->world == 1 || !in_window || (n & 3) || (n_params & 3) || ((uintptr_t)w & 15) || ((size_t)(b - c->local) & 15)) {
        // not expressible as the fused kernel: the same result in two steps
        int e = allreduce<float>(c, g, n, as_stream(stream));
        if (e) return e;
        return sn_sgd_update(w, g, n_params, lr, 1.0f, stream);
    }
    DevTable t;
    for (int p = 0; p < MAX_RANKS; ++p) t.win[p] = c->peer[p < c->world ? p : 0];
    SN_CUDA(launch_dependent(allreduce_inplace_kernel<float, true>, dim3(grid_two_shot((long long)bytes, c->world)), dim3(INPLACE_THREADS), 0,
                             as_stream(stream), t, c->lay.flag_off, c->state, (size_t)(b - c->local), n, c->rank, c->world, w, n_params, lr));
    return after_launch("allreduce_inplace_kernel (sgd)");
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
