"""-m gpu, needs >= 2 GPUs (skipped otherwise): the hand-written NVLink all-reduce (csrc/comm.cu) called
straight through the C ABI from ONE process driving every GPU (sn_comm_init_all): every replica must end
with the bit-exact rank-ordered sum, for the in-place two-shot path (tensor inside the window), the
one-shot staging path (small float64 / float32 tensors anywhere) and the staged path (large tensor outside
the window), back to back without host synchronisation, and from a replayed CUDA graph."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(window_floats):
    import torch
    G = torch.cuda.device_count()
    if G < 2:
        pytest.skip('needs two GPUs')
    G = min(G, 8)
    from shinnosuke_b200.comm import Parallel
    pars = Parallel.init_all(G, window_floats)
    streams = []
    for r in range(G):
        with torch.cuda.device(r):
            streams.append(torch.cuda.Stream())
    return torch, G, pars, streams


def _ordered_sum(parts):
    acc = parts[0].copy()
    for p in parts[1:]:
        acc = acc + p                      # same order and the same type as the kernels: bit-exact
    return acc


@pytest.mark.parametrize('n', [1, 7, 4096, 16385, 100003, 972554 + 64])
def test_inplace_window_allreduce_f32(n):
    torch, G, pars, streams = _setup(n + 64)
    rng = np.random.default_rng(n)
    host = [rng.standard_normal(n).astype(np.float32) for _ in range(G)]
    views = []
    for r in range(G):
        v = pars[r].grad.carve_f32(n)
        v.copy_(torch.from_numpy(host[r]))
        views.append(v)
    for r in range(G):
        torch.cuda.synchronize(r)
    for r in range(G):
        with torch.cuda.device(r):
            pars[r].allreduce_grads(views[r], streams[r].cuda_stream)
    want = _ordered_sum(host)
    for r in range(G):
        torch.cuda.synchronize(r)
        assert np.array_equal(views[r].cpu().numpy(), want), (n, r)


@pytest.mark.parametrize('dtype,n', [('f64', 16 * 64), ('f64', 16 * 256 + 3), ('f64', 16 * 1024), ('f32', 5), ('f32', 60000)])
def test_oneshot_allreduce_of_small_tensors_anywhere(dtype, n):
    torch, G, pars, streams = _setup(0)
    rng = np.random.default_rng(n)
    npt, tt = (np.float64, torch.float64) if dtype == 'f64' else (np.float32, torch.float32)
    bufs, rounds = [], 6
    host = [[rng.standard_normal(n).astype(npt) for _ in range(G)] for _ in range(rounds)]
    for r in range(G):
        bufs.append([torch.from_numpy(host[k][r]).to('cuda:%d' % r) for k in range(rounds)])
    for r in range(G):
        torch.cuda.synchronize(r)
    # `rounds` exchanges back to back: the double-buffered staging slots and the monotonic flags
    for k in range(rounds):
        for r in range(G):
            with torch.cuda.device(r):
                if dtype == 'f64':
                    pars[r].allreduce_stats(bufs[r][k], streams[r].cuda_stream)
                else:
                    pars[r].stats.allreduce_f32(bufs[r][k].data_ptr(), n, streams[r].cuda_stream)
    for r in range(G):
        torch.cuda.synchronize(r)
    for k in range(rounds):
        want = _ordered_sum(host[k])
        for r in range(G):
            assert np.array_equal(bufs[r][k].cpu().numpy(), want), (k, r)


def test_large_tensor_outside_the_window_is_staged():
    n = 300000
    torch, G, pars, streams = _setup(100000)            # window smaller than the tensor: three pieces
    rng = np.random.default_rng(3)
    host = [rng.standard_normal(n).astype(np.float32) for _ in range(G)]
    bufs = [torch.from_numpy(host[r]).to('cuda:%d' % r) for r in range(G)]
    for r in range(G):
        torch.cuda.synchronize(r)
    for r in range(G):
        with torch.cuda.device(r):
            pars[r].grad.allreduce_f32(bufs[r].data_ptr(), n, streams[r].cuda_stream)
    want = _ordered_sum(host)
    for r in range(G):
        torch.cuda.synchronize(r)
        assert np.array_equal(bufs[r].cpu().numpy(), want)


def test_allreduce_inside_a_replayed_graph():
    """What the training step does: exchange captured once, replayed; the sequence number is device state."""
    from shinnosuke_b200._lib import lib
    n = 200000
    torch, G, pars, streams = _setup(n)
    views = [pars[r].grad.carve_f32(n) for r in range(G)]
    small = [torch.zeros(1024, dtype=torch.float64, device='cuda:%d' % r) for r in range(G)]
    graphs = []
    for r in range(G):
        torch.cuda.synchronize(r)
    for r in range(G):
        with torch.cuda.device(r):
            h = ctypes.c_void_p()
            lib.sn_graph_begin(streams[r].cuda_stream)
            pars[r].allreduce_stats(small[r], streams[r].cuda_stream)
            pars[r].allreduce_grads(views[r], streams[r].cuda_stream)
            lib.sn_graph_end(streams[r].cuda_stream, ctypes.byref(h))
            graphs.append(h)
    rng = np.random.default_rng(11)
    for it in range(5):
        host = [rng.standard_normal(n).astype(np.float32) for _ in range(G)]
        hs = [rng.standard_normal(1024) for _ in range(G)]
        for r in range(G):
            with torch.cuda.device(r), torch.cuda.stream(streams[r]):
                views[r].copy_(torch.from_numpy(host[r]), non_blocking=False)
                small[r].copy_(torch.from_numpy(hs[r]), non_blocking=False)
        for r in range(G):
            with torch.cuda.device(r):
                lib.sn_graph_launch(graphs[r], streams[r].cuda_stream)
        for r in range(G):
            torch.cuda.synchronize(r)
            assert np.array_equal(views[r].cpu().numpy(), _ordered_sum(host)), (it, r)
            assert np.array_equal(small[r].cpu().numpy(), _ordered_sum(hs)), (it, r)
    for r in range(G):
        with torch.cuda.device(r):
            lib.sn_graph_destroy(graphs[r])


def test_fused_exchange_and_sgd_update():
    """sn_allreduce_sgd_update: the gradient exchange and StochasticGradientDescent.update
    (utils/Optimizers.py:29-32) in one kernel — w -= lr * sum_ranks(g), g = 0 over the parameters, the
    metric words behind them reduced and kept; replicas bit-identical."""
    from shinnosuke_b200._lib import lib
    n_params, n = 972576, 972576 + 32
    torch, G, pars, streams = _setup(n)
    rng = np.random.default_rng(5)
    w0 = rng.standard_normal(n).astype(np.float32)
    host = [rng.standard_normal(n).astype(np.float32) for _ in range(G)]
    gs, ws = [], []
    for r in range(G):
        g = pars[r].grad.carve_f32(n)
        g.copy_(torch.from_numpy(host[r]))
        gs.append(g)
        ws.append(torch.from_numpy(w0).to('cuda:%d' % r))
    for r in range(G):
        torch.cuda.synchronize(r)
    for r in range(G):
        with torch.cuda.device(r):
            lib.sn_allreduce_sgd_update(pars[r].grad.handle, gs[r].data_ptr(), ws[r].data_ptr(), n, n_params, 0.01, streams[r].cuda_stream)
    total = _ordered_sum(host)
    want = w0.astype(np.float64) - 0.01 * total.astype(np.float64)
    first = None
    for r in range(G):
        torch.cuda.synchronize(r)
        w = ws[r].cpu().numpy()
        g = gs[r].cpu().numpy()
        assert np.abs(w[:n_params] - want[:n_params]).max() < 1e-6 * np.abs(want).max()
        assert np.array_equal(w[n_params:], w0[n_params:])
        assert not g[:n_params].any()
        assert np.array_equal(g[n_params:], total[n_params:])
        if first is None:
            first = w
        assert np.array_equal(w, first)
