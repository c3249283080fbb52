"""-m gpu: the C ABI alone trains.  tests/cabi_step.c — plain C99, compiled here with gcc against
include/shinnosuke_b200.h and linked to libshinnosuke_b200.so, no Python and no PyTorch in the process —
runs three SGD steps of BASELINE config 1 (eager, graph capture, graph replay) and must print the
live-reference trajectory of tests/golden/traj_cfg1.npz (KAT D)."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_err
from oracle import shinnosuke_oracle as O

pytestmark = pytest.mark.gpu


def test_c_program_reproduces_kat_d(tmp_path):
    from shinnosuke_b200 import Model, Input, Conv2D, MaxPooling2D, Flatten, Dense
    g = load_golden('traj_cfg1')
    r = np.random.default_rng(1234)
    X = r.random((768, 1, 28, 28)).astype(np.float32)
    Y = O.to_categorical(r.integers(0, 10, 768)).astype(np.float32)
    assert abs(float(X.astype(np.float64).sum()) - float(g['xsum'])) < 1e-6 * float(g['xsum'])
    # seed-0 initial weights in the reference's draw order (host side only: nothing touches the GPU here)
    np.random.seed(0)
    X_in = Input(shape=(None, 1, 28, 28))
    conv = Conv2D(8, (2, 2), padding='VALID', initializer='normal', activation='relu')(X_in)
    y = Dense(10, initializer='normal', activation='softmax')(Flatten()(MaxPooling2D((2, 2))(conv)))
    Model(inputs=X_in, outputs=y).compile(optimizer='sgd', loss='sparse_categorical_crossentropy')
    cw = np.asarray(conv.variables[0].output_tensor)             # (F, C, kh, kw), C = 1
    cb = np.asarray(conv.variables[1].output_tensor).reshape(-1)
    dw = np.asarray(y.variables[0].output_tensor)                 # (1352, 10), rows in the reference's (c, i, j) order
    d = str(tmp_path)
    X.tofile(os.path.join(d, 'x.bin'))                            # one channel: NCHW == NHWC
    Y.tofile(os.path.join(d, 'y.bin'))
    cw.transpose(0, 2, 3, 1).astype(np.float32).tofile(os.path.join(d, 'conv_w.bin'))
    cb.astype(np.float32).tofile(os.path.join(d, 'conv_b.bin'))
    dw.reshape(8, 13, 13, 10).transpose(3, 1, 2, 0).astype(np.float32).tofile(os.path.join(d, 'dense_w.bin'))   # [n_out][(i, j, c)]
    exe = os.path.join(d, 'cabi_step')
    lib_dir = os.path.join(ROOT, 'shinnosuke_b200', 'lib')
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-O1', os.path.join(ROOT, 'tests', 'cabi_step.c'), '-o', exe,
                           '-L' + lib_dir, '-lshinnosuke_b200', '-Wl,-rpath,' + lib_dir])
    # the process must not have PyTorch (or Python) in it: only libc, the CUDA driver and our library
    ldd = subprocess.run(['ldd', exe], capture_output=True, text=True).stdout
    assert 'torch' not in ldd and 'python' not in ldd
    out = subprocess.run([exe, d], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    losses = [float(m) for m in re.findall(r'loss ([-0-9.e]+)', out.stdout)]
    accs = [float(m) for m in re.findall(r'acc ([-0-9.e]+)', out.stdout)]
    assert np.allclose(losses, g['train_loss'], rtol=1e-5), (losses, g['train_loss'])
    assert np.allclose(accs, g['train_acc'], atol=1e-6)
    cw3 = np.array([float(v) for v in re.search(r'conv_w (.*)', out.stdout).group(1).split()]).reshape(8, 2, 2, 1).transpose(0, 3, 1, 2)
    cb3 = np.array([float(v) for v in re.search(r'conv_b (.*)', out.stdout).group(1).split()]).reshape(1, 8)
    assert rel_err(cw3, g['convW']) < 1e-5 and rel_err(cb3, g['convb']) < 1e-5
    assert int(re.search(r'launches (\d+)', out.stdout).group(1)) > 0
