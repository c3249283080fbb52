"""-m gpu, needs >= 2 GPUs (skipped otherwise): one process per GPU over NCCL.  Two data-parallel
ranks, each training on its contiguous half of every minibatch with the overlapped bucketed
all-reduce of models._backward_allreduce, must land on the float64 oracle's weights for the WHOLE
minibatch (fp32 tier, 1e-5) and on bit-identical replicas, for SGD and Momentum."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_training_matches_whole_batch_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(root, 'tests', 'multi_gpu', 'ddp_check.py')]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'DDP_CHECK_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_fit_n_gpus_from_one_process():
    """models.py fit() sharding each minibatch over the GPUs of the box by itself: fit(n_gpus=2) in ONE process
    (replica threads + peer-memory exchange kernels) against the whole-batch oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'tests', 'multi_gpu', 'fit_n_gpus_check.py'), '2'], cwd=root,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'FIT_N_GPUS_OK' in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
