"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path — contiguous row
shards of each minibatch, loss gradients divided by the GLOBAL row count, summed with an
all-reduce — reproduces the full-batch gradients of the (oracle-pinned) reference algorithm.
The device kernels are not involved; this covers the sharding arithmetic fit() uses."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import shinnosuke_oracle as O
    from shinnosuke_b200.utils.MiniBatch import batch_permutation, batch_slices
    r = np.random.default_rng(5)
    X = r.standard_normal((50, 1, 12, 12))
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 40)]))
    spec = [('conv', 4, 3, 1, 'SAME', 'relu'), ('pool', 2), ('flatten',), ('dense', 10, 'softmax')]
    np.random.seed(0)
    net = O.OracleNet(spec, (None, 1, 12, 12))
    flat_after = None
    for epoch in range(2):
        perm = batch_permutation(50, epoch, True)
        for a, b in batch_slices(50, 16):                      # 16,16,16,2: the ragged tail gives rank shards 1/1
            rows = b - a
            lo = a + (rows * rank) // world
            hi = a + (rows * (rank + 1)) // world
            xs, ys = X[perm][lo:hi], Y[perm][lo:hi]
            y_hat, caches = net.forward(xs, True)
            net.backward((y_hat - ys) / rows, caches)          # GLOBAL rows, like sn_scce_bwd(denom_rows)
            flat = torch.from_numpy(np.concatenate([g.ravel() for g in net.grads]))
            dist.all_reduce(flat)                               # sum, no 1/world: shards already divide by the global N
            off = 0
            for i, g in enumerate(net.grads):
                net.grads[i] = flat[off:off + g.size].numpy().reshape(g.shape).copy()
                off += g.size
            O.sgd_update(net.params, net.grads, 0.01)
    flat_after = np.concatenate([p.ravel() for p in net.params])
    if rank == 0:
        q.put(flat_after)
    dist.destroy_process_group()


def test_two_rank_sharded_training_equals_single_process():
    from oracle import shinnosuke_oracle as O
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r = np.random.default_rng(5)
    X = r.standard_normal((50, 1, 12, 12))
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 40)]))
    spec = [('conv', 4, 3, 1, 'SAME', 'relu'), ('pool', 2), ('flatten',), ('dense', 10, 'softmax')]
    np.random.seed(0)
    net = O.OracleNet(spec, (None, 1, 12, 12))
    net.fit(X, Y, 16, 2, shuffle=True)
    ref = np.concatenate([p.ravel() for p in net.params])
    assert np.max(np.abs(got - ref)) < 1e-12
