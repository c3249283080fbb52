#!/usr/bin/env python
"""Headline benchmark: training samples/s of the data-parallel hot path.

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Workload (BASELINE.json configs[2]): VGG-style CNN on synthetic CIFAR-shaped 3x32x32 data,
4x[Conv2D 3x3 SAME + ReLU -> BatchNormalization -> MaxPooling 2x2] with 64-128-256-256 filters,
Flatten, Dense(10, softmax), SGD; batch 512 per GPU (weak scaling: global batch 512*N).
A "step" = forward sweep, softmax cross-entropy, backward sweep, gradient all-reduce (N>1),
fused SGD update over the flat parameter arena.

JSON line (rank 0): see the bench contract in DESIGN.md.  `value` is measured with the step's
inputs already resident in HBM (CUDA-graph replays); `e2e` goes through Sequential.fit() with
HOST buffers — every step copies its inputs host->device from pinned memory and copies the
(loss, acc) pair back.  `roofline` is the dominant kernel timed with CUDA events on its own
stream during an instrumented pass of the same steps; `cpu_baseline` is the CPU oracle (a NumPy
port of the reference's algorithm, pinned to the reference) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    'cfg3': dict(name='VGG-style CNN 3x32x32: 4x[Conv2D(64-128-256-256,3x3,SAME,relu)+BatchNorm+MaxPool2x2]+Dense10, SGD',
                 shape=(3, 32, 32), per_gpu_batch=512, flops_per_sample=290.3e6),
    'cfg1': dict(name='README CNN 1x28x28: Conv2D(8,2x2,VALID,relu)+MaxPool2x2+Flatten+Dense10, SGD',
                 shape=(1, 28, 28), per_gpu_batch=256, flops_per_sample=0.174e6),
    'cfg2': dict(name='MLP 784-500-10, SGD', shape=(784,), per_gpu_batch=128, flops_per_sample=1.60e6),
    'cfg4': dict(name='ResNet-style 3x32x32 (functional API): Conv64 + residual block@64 + pool + Conv128 + residual block@128 '
                      '+ pool + Dense10, BatchNorm in the blocks, SGD', shape=(3, 32, 32), per_gpu_batch=128,
                 flops_per_sample=1026.8e6),
    'cfg5': dict(name='Embedding(1000->64) + LSTM(128) over 64 timesteps + Dense2, SGD (recurrent GEMM path)', shape=(64,),
                 per_gpu_batch=256, flops_per_sample=3 * 2 * 64 * 192 * 512, classes=2),
}


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], bf16=p['bf16_tflops'], bf16_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured')
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source='fallback')


def synth(cfg, n, seed=1234):
    r = np.random.default_rng(seed)
    shape = CONFIGS[cfg]['shape']
    if cfg in ('cfg3', 'cfg4'):
        X = r.standard_normal((n,) + shape, dtype=np.float32)
    elif cfg == 'cfg5':
        X = r.integers(0, 1000, (n,) + shape).astype(np.float32)       # token ids travel as float32 values
    else:
        X = r.random((n,) + shape, dtype=np.float32)
    classes = CONFIGS[cfg].get('classes', 10)
    labels = r.integers(0, classes, n)
    labels[:classes] = np.arange(classes)
    Y = np.eye(classes, dtype=np.float32)[labels]
    return X, Y


def build_model(cfg, precision, sync_bn=None):
    from shinnosuke_b200 import (Sequential, Model, Input, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense,
                                 Activation, Add)
    np.random.seed(0)
    if cfg == 'cfg4':
        def block(h, width):
            r = Conv2D(width, (3, 3), padding='SAME', activation='linear', initializer='normal')(h)
            r = Activation('relu')(BatchNormalization()(r))
            r = Conv2D(width, (3, 3), padding='SAME', activation='linear', initializer='normal')(r)
            r = BatchNormalization()(r)
            return MaxPooling2D((2, 2))(Activation('relu')(Add()([h, r])))
        x_in = Input(shape=(None, 3, 32, 32))
        h = Conv2D(64, (3, 3), padding='SAME', activation='relu', initializer='normal')(x_in)
        h = block(h, 64)
        h = Conv2D(128, (3, 3), padding='SAME', activation='relu', initializer='normal')(h)
        h = block(h, 128)
        y = Dense(10, activation='softmax')(Flatten()(h))
        m = Model(inputs=x_in, outputs=y)
        m.compile(optimizer='sgd', loss='sparse_categorical_crossentropy', precision=precision, sync_bn=sync_bn)
        return m
    m = Sequential()
    if cfg == 'cfg5':
        from shinnosuke_b200 import Embedding, LSTM
        m.add(Embedding(1000, 64, input_length=64))
        m.add(LSTM(128))
        m.add(Dense(2, activation='softmax'))
    elif cfg == 'cfg3':
        first = True
        for f in (64, 128, 256, 256):
            kw = dict(input_shape=(None, 3, 32, 32)) if first else {}
            first = False
            m.add(Conv2D(f, (3, 3), padding='SAME', activation='relu', initializer='normal', **kw))
            m.add(BatchNormalization())
            m.add(MaxPooling2D((2, 2)))
        m.add(Flatten())
        m.add(Dense(10, activation='softmax'))
    elif cfg == 'cfg1':
        m.add(Conv2D(8, (2, 2), input_shape=(None, 1, 28, 28), padding='VALID', activation='relu', initializer='normal'))
        m.add(MaxPooling2D((2, 2)))
        m.add(Flatten())
        m.add(Dense(10, activation='softmax', initializer='normal'))
    else:
        m.add(Dense(500, activation='relu', n_in=784))
        m.add(Dense(10, activation='softmax'))
    m.compile(optimizer='sgd', loss='sparse_categorical_crossentropy', precision=precision, sync_bn=sync_bn)
    return m


def oracle_spec(cfg):
    from oracle import shinnosuke_oracle as O
    return {'cfg1': O.spec_cfg1, 'cfg2': O.spec_cfg2, 'cfg3': O.spec_cfg3, 'cfg4': O.spec_cfg4, 'cfg5': O.spec_cfg5}[cfg]()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self, wait_s=8.0):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                       '-lms', '10'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
            return
        # nvidia-smi needs a while to come up (seconds on an 8-GPU box): the timed region is only tens
        # of milliseconds long, so do not enter it before the first sample has been written
        t0 = time.time()
        while time.time() - t0 < wait_s and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------- roofline table
def kernel_work(name, a, precision):
    """Algorithmic work of one launch from its integer arguments (SURVEY.md 8(d) formulas).
    Returns (flops, bytes): the bytes are the tensors the operation must read and write once."""
    conv = ('sn_conv2d_fprop', 'sn_conv2d_dgrad', 'sn_conv2d_wgrad', 'sn_conv2d_fprop_bf16', 'sn_conv2d_dgrad_bf16',
            'sn_conv2d_wgrad_bf16', 'sn_conv2d_fprop_stats', 'sn_conv2d_fprop_bf16_stats', 'sn_conv2d_fprop_x3',
            'sn_conv2d_dgrad_x3', 'sn_conv2d_wgrad_x3', 'sn_conv2d_fprop_bf16_stats_ybf16', 'sn_conv2d_fprop_stats_ybf16',
            'sn_conv2d_stem_fprop', 'sn_conv2d_stem_wgrad_dybf16')
    if name in conv:
        N, C, H, W, F, kh, kw, stride, ph, pw = a[:10]
        oh = (H + 2 * ph - kh) // stride + 1
        ow = (W + 2 * pw - kw) // stride + 1
        x, y, w = N * C * H * W, N * F * oh * ow, F * C * kh * kw
        flops = 2.0 * y * C * kh * kw
        if name == 'sn_conv2d_stem_fprop':
            byt = {'fprop': 4 * x + 4 * w + (2 if (len(a) > 11 and a[11]) else 4) * y}       # fp32 image and filters in, fp32 or bf16 map out
        elif name == 'sn_conv2d_stem_wgrad_dybf16':
            byt = {'wgrad': 4 * x + 2 * y + 4 * w}              # fp32 image and the bf16 gradient map in, fp32 dw out
        elif name == 'sn_conv2d_fprop_stats_ybf16':
            byt = {'fprop': 4 * x + 4 * w + 2 * y}              # the stem: fp32 image and filters in, bf16 map out
        elif name == 'sn_conv2d_fprop_bf16_stats_ybf16':
            byt = {'fprop': 2 * x + 2 * w + 2 * y}              # bf16 operand copies in, bf16 map out
        elif 'bf16' in name:
            # bf16 operand copies in, fp32 result out (fprop: y, dgrad: dx = x-sized, wgrad: dw)
            byt = {'fprop': 2 * x + 2 * w + 4 * y, 'dgrad': 2 * y + 2 * w + 4 * x, 'wgrad': 2 * x + 2 * y + 4 * w}
        else:
            byt = {'fprop': 4 * (x + w + y), 'dgrad': 4 * (x + w + y), 'wgrad': 4 * (x + w + y)}
        return flops, float(byt['fprop' if 'fprop' in name else ('dgrad' if 'dgrad' in name else 'wgrad')])
    if name in ('sn_dense_fprop', 'sn_dense_dgrad', 'sn_dense_wgrad'):
        M, K, Nout = a[:3]
        return 2.0 * M * K * Nout, 4.0 * (M * K + K * Nout + M * Nout)
    if name in ('sn_maxpool2d_fwd', 'sn_maxpool2d_bwd'):
        N, H, W, C, pool, stride = a[:6]
        oh, ow = (H - pool) // stride + 1, (W - pool) // stride + 1
        return 0.0, 4.0 * N * C * (H * W + oh * ow) + N * C * oh * ow
    if name == 'sn_batchnorm_fwd_train':
        rows, C = a[0], a[1]
        return 0.0, 4.0 * rows * C * 3          # read x (stats), read x + write y (apply)
    if name == 'sn_batchnorm_bwd':
        rows, C = a[0], a[1]
        return 0.0, 4.0 * rows * C * 5          # read dy,x (reduce), read dy,x + write dx (apply)
    if name == 'sn_bn_pool_fwd_train':
        N, H, W, C = a[:4]
        return 0.0, 8.0 * N * H * W * C + 5.0 * N * (H // 2) * (W // 2) * C     # x read twice; pooled y + code written
    if name in ('sn_bn_pool_fwd_apply', 'sn_bn_pool_fwd_apply_xbf16'):
        N, H, W, C = a[:4]
        xbytes = 2.0 if name.endswith('xbf16') else 4.0
        bf16_out = a[-1] if len(a) >= 6 else 0          # + the bf16 copy of the pooled map the next convolution reads
        return 0.0, xbytes * N * H * W * C + (5.0 + 2.0 * bf16_out) * N * (H // 2) * (W // 2) * C     # x read once; pooled y + code written
    if name == 'sn_batchnorm_fwd_apply':
        return 0.0, 8.0 * a[0] * a[1]
    if name in ('sn_bn_pool_bwd', 'sn_bn_pool_bwd_xbf16', 'sn_bn_pool_bwd_apply', 'sn_bn_pool_bwd_apply_xbf16'):
        N, H, W, C = a[:4]
        xbytes = 2.0 if name.endswith('xbf16') else 4.0
        pooled = 9.0 if 'apply' not in name else 5.0           # the apply half alone reads pooled dy + code only
        # algorithmic: x read once, dx written once (fp32 and / or the bf16 copy the producing convolution's
        # wgrad / dgrad read: the last two key entries say which), pooled dy, pooled y and the one-byte code
        # read once (the kernels read dy_pool twice, once per pass: that re-read is NOT algorithmic traffic)
        f32_out, bf16_out = (a[-2], a[-1]) if len(a) >= 8 else (1, 0)
        return 0.0, (xbytes + 4.0 * f32_out + 2.0 * bf16_out) * N * H * W * C + pooled * N * (H // 2) * (W // 2) * C
    if name == 'sn_sgd_update':
        return 0.0, 16.0 * a[0]
    if name == 'sn_momentum_update':
        return 0.0, 20.0 * a[0]
    if name == 'sn_adaptive_update':          # (kind, P): w r+w, g r, one (RMSprop, AdaGrad) or two (AdaDelta, Adam) state words r+w
        return 0.0, (28.0 if a[0] >= 2 else 20.0) * a[1]
    if name == 'sn_conv2d_stem_prepare':
        N, C, H, W = a[:4]
        return 0.0, 4.0 * N * H * W * (C + 4)
    if name == 'sn_cast_f32_to_bf16':
        return 0.0, 6.0 * a[0]
    if name in ('sn_relu_bwd',):
        return 0.0, 12.0 * a[0]
    if name in ('sn_nchw_to_nhwc', 'sn_nhwc_to_nchw'):
        return 0.0, 8.0 * a[0] * a[1] * a[2] * a[3]
    if name in ('sn_lstm_step_fwd', 'sn_lstm_step_bwd'):
        # one timestep: the recurrent product h . Wh (N x H x 4H) + the gate tensors once
        N, H = (a[2], a[3]) if name == 'sn_lstm_step_fwd' else (a[1], a[2])
        return 2.0 * N * H * 4 * H, 4.0 * (4 * H * H + 2 * N * 4 * H + 4 * N * H)
    if name in ('sn_lstm_seq_fwd', 'sn_lstm_seq_bwd', 'sn_lstm_seq_fwd_tc', 'sn_lstm_seq_bwd_tc'):
        # the whole recurrence: T dependent steps (latency-bound by construction); algorithmic work = the
        # recurrent products + the per-step tensors read / written once (zx or dh, acts, c, h / dz)
        T, N, H = a[-3:]
        return 2.0 * T * N * H * 4 * H, 4.0 * T * N * (2 * 4 * H + 2 * H) + 4.0 * 4 * H * H
    if name in ('sn_embedding_fwd', 'sn_embedding_bwd'):
        n_idx, vocab, dim = a[:3]
        return 0.0, 4.0 * n_idx * (1 + 2 * dim)
    if name == 'sn_swap_leading_axes':
        return 0.0, 8.0 * a[0] * a[1] * a[2]
    return 0.0, 0.0


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    `ncu --set full` capture, keyed like the kernel table; see profiles/README.md."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            return json.load(f)
    except Exception:
        return {}


def roofline_from_profile(summary, precision, pk):
    """Pick the kernel class with the largest share of the step and report it against its roof: the
    tensor pipe if its flops at the tensor peak take longer than its bytes at the HBM peak, else HBM."""
    rows = []
    total = 0.0
    for (name, a), ms_list in summary.items():
        ms = float(np.mean(ms_list))
        flops, byt = kernel_work(name, a, precision)
        # the roof of the kernel that actually ran: bf16 entry points against the bf16 peak, the
        # fp32-pointer tensor kernels (tf32) against half of it (also the reference point of the fp32 tier)
        tensor_peak = pk['bf16'] if ('_bf16' in name and name != 'sn_conv2d_fprop_stats_ybf16') else pk['bf16'] / 2
        t_tensor, t_hbm = flops / (tensor_peak * 1e12), byt / (pk['hbm'] * 1e9)
        kind = 'flops' if t_tensor > t_hbm else 'bytes'
        rows.append(dict(name=name, args=list(a), ms=ms, launches=len(ms_list), kind=kind, work=flops if kind == 'flops' else byt,
                         flops=flops, bytes=byt, tensor_peak=tensor_peak))
        total += float(np.sum(ms_list))
    rows.sort(key=lambda r: -r['ms'] * r['launches'])
    for r in rows:
        r['share'] = r['ms'] * r['launches'] / total if total else 0.0
    top = rows[0]
    if top['kind'] == 'flops':
        achieved = top['flops'] / (top['ms'] * 1e-3) / 1e12
        roof = dict(bound='tensor', achieved=achieved, peak=top['tensor_peak'], unit='TFLOP/s', frac=achieved / top['tensor_peak'])
    else:
        achieved = top['bytes'] / (top['ms'] * 1e-3) / 1e9
        roof = dict(bound='hbm', achieved=achieved, peak=pk['hbm'], unit='GB/s', frac=achieved / pk['hbm'])
    key = '%s|%s|%s' % (precision, top['name'], ','.join(str(v) for v in top['args'][:4]))
    roof.update(kernel=top['name'], kernel_args=top['args'], kernel_ms=top['ms'], share_of_step=top['share'],
                traffic=ncu_traffic().get(key), peak_source=pk['source'])
    # the tensor-bound side of the step (BASELINE.json's second metric: "conv fwd+bwd TFLOP/s vs peak"): the top
    # tensor-bound entry, and all large-channel contractions of one step together
    tens = [r for r in rows if r['kind'] == 'flops']
    if tens:
        tt = tens[0]
        ach = tt['flops'] / (tt['ms'] * 1e-3) / 1e12
        roof_t = dict(bound='tensor', achieved=ach, peak=tt['tensor_peak'], unit='TFLOP/s', frac=ach / tt['tensor_peak'], kernel=tt['name'],
                      kernel_args=tt['args'], kernel_ms=tt['ms'], share_of_step=tt['share'], peak_source=pk['source'])
        n_steps = max(r['launches'] for r in rows)
        fl = sum(r['flops'] * r['launches'] for r in tens) / n_steps
        ms = sum(r['ms'] * r['launches'] for r in tens) / n_steps
        peak = pk['bf16'] if precision == 'bf16' else pk['bf16'] / 2
        roof_t['all_contractions'] = dict(gflop_per_step=fl / 1e9, ms_per_step=ms, tflops=fl / (ms * 1e-3) / 1e12 if ms else 0.0,
                                          frac_of_peak=(fl / (ms * 1e-3) / 1e12 / peak) if ms else 0.0, peak=peak,
                                          entries=len(tens))
    else:
        roof_t = None
    return roof, rows, roof_t


# --------------------------------------------------------------- CPU legs
def cpu_oracle_throughput(cfg, batch, steps, warmup=1):
    """The reference's algorithm (NumPy float64, all host cores through OpenBLAS) on `steps`
    minibatches of `batch` samples of the workload."""
    from oracle import shinnosuke_oracle as O
    spec, ishape = oracle_spec(cfg)
    np.random.seed(0)
    net = O.OracleNet(spec, ishape)
    X, Y = synth(cfg, batch * (steps + warmup))
    X = X.astype(np.float64)
    Y = Y.astype(np.float64)
    with all_host_threads():
        for i in range(warmup):
            net.step(X[i * batch:(i + 1) * batch], Y[i * batch:(i + 1) * batch])
        t0 = time.perf_counter()
        for i in range(warmup, warmup + steps):
            net.step(X[i * batch:(i + 1) * batch], Y[i * batch:(i + 1) * batch])
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def all_host_threads():
    """Context manager: BLAS thread pools at the number of host cores.  torchrun exports OMP_NUM_THREADS=1 to every
    rank, which would time the reference arm on ONE core at N > 1; the contract is "all the host threads it can use"."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=os.cpu_count())
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [d.get('num_threads', 0) for d in threadpool_info() if d.get('user_api') == 'blas']
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def cpu_reference_throughput(cfg, batch, steps, warmup=1, warmup_batch=None):
    """The UNMODIFIED reference (oracle/_ref, copied from /root/reference by oracle/make_ref.sh) through its
    own Sequential.fit(): float64 NumPy, all host cores through OpenBLAS, same synthetic data and the same
    seed-0 initial weights as the GPU arm.  Returns (samples/s, s/step, losses) or None when the reference
    cannot build this config as shipped (configs 4, 5: SURVEY.md 8(c)) or oracle/_ref is missing."""
    from oracle import ref_driver
    if not ref_driver.can_build(cfg):
        return None
    wb = warmup_batch or batch
    X, Y = synth(cfg, batch * steps + wb * warmup)
    with all_host_threads():
        if warmup and wb != batch:
            # a short untimed pass on a small batch first (imports, OpenBLAS thread pool), then the timed steps
            ref_driver.fit_throughput(cfg, X[:wb * warmup], Y[:wb * warmup], wb, 0, warmup)
            return ref_driver.fit_throughput(cfg, X[wb * warmup:], Y[wb * warmup:], batch, 0, steps)
        return ref_driver.fit_throughput(cfg, X, Y, batch, warmup, steps)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.
    oracle/_ref holds the unmodified reference package (pure Python, so "building" it is a copy:
    oracle/make_ref.sh); its own Sequential.fit() runs the SAME config as the GPU arm — same layers, same
    minibatch size, same synthetic data, same initial weights (`kind: "reference"`, `same_config: true`).
    A step of config 3 at batch 512 takes ~10 s on 16 cores, so the number of timed steps is bounded
    (stated in `sample`).  Configs the reference cannot construct as shipped (4: Add crashes at graph build;
    5: wrong LSTM backward) and a box without oracle/_ref fall back to the NumPy port (`kind: "port"`)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = args.config
    B = args.per_gpu_batch or CONFIGS[cfg]['per_gpu_batch']
    cores = os.cpu_count()
    budget_steps = {'cfg3': 2, 'cfg1': 30, 'cfg2': 100}.get(cfg, 3)
    steps = max(1, min(args.steps, budget_steps))
    warm = max(1, min(args.warmup, 1))
    res = cpu_reference_throughput(cfg, B, steps, warm, warmup_batch=64 if cfg == 'cfg3' else None)
    if res is not None:
        value, s_per_step, losses = res
        kind, sample_batch, same = 'reference', B, True
        sample = ('%d timed minibatches of %d through the unmodified reference Sequential.fit() (float64 NumPy, OpenBLAS on all %d '
                  'cores), after %d untimed warm-up minibatch(es)' % (steps, B, cores, warm))
    else:
        sample_batch = {'cfg3': 64, 'cfg1': 256, 'cfg2': 128, 'cfg4': 32, 'cfg5': 256}[cfg]
        value, s_per_step = cpu_oracle_throughput(cfg, sample_batch, steps, warm)
        kind, same, losses = 'port', sample_batch == B, None
        sample = '%d steps of batch %d (float64 NumPy port of the reference, OpenBLAS threads = all %d cores)' % (steps, sample_batch, cores)
    line = dict(impl='reference', metric='train_samples_per_sec', value=value, unit='samples/s', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, steps_timed=steps, ms_per_step=s_per_step * 1e3, higher_is_better=True,
                scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload=CONFIGS[cfg]['name'], per_gpu_batch=B, global_batch=B, sample_batch=sample_batch,
                            same_config=same),
                cpu_baseline=dict(value=value, unit='samples/s', cores=cores, kind=kind, sample=sample),
                e2e=dict(value=value, unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0,
                losses=losses)
    print(json.dumps(line))


# --------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from shinnosuke_b200._lib import lib
    from shinnosuke_b200.device import current_stream

    cfg = args.config
    B = args.per_gpu_batch or CONFIGS[cfg]['per_gpu_batch']
    G = B * world                                   # global minibatch (weak scaling)
    K, Wm = args.steps, max(3, args.warmup)
    precision = args.precision
    pk = peaks()
    m = build_model(cfg, precision, False if args.no_sync_bn else None)

    # every rank holds the whole synthetic set, like every rank running the same script would
    X, Y = synth(cfg, G * max(K, Wm, 4))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: through the public API with HOST buffers (e2e) ----------------
    Xt, Yt = X[:G * K], Y[:G * K]
    # warm-up: same call, same arrays (captures the step graphs, page-locks the host arrays once)
    m.fit(Xt, Yt, batch_size=G, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    m.fit(Xt, Yt, batch_size=G, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    losses_e2e = list(m.train_loss[-K:])

    # ---- leg 2: inputs resident in HBM, graph replays (value) -------------------
    inp = m._input_layer()
    n_loc = B
    xs = inp._buf('fit_x0', (n_loc,) + X.shape[1:], 'nhwc' if X.ndim == 4 else 'flat')
    ys = inp._buf('fit_y0', (n_loc, Y.shape[1]), 'flat')
    for _ in range(Wm):
        m._run_step(xs, ys, G)
    barrier()
    ev0.record()
    for _ in range(K):
        m._run_step(xs, ys, G)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()

    # ---- the BASELINE.json variant of config 3 / 4: global batch 512 (1024) over 8 GPUs = 64 (128) per GPU.  The
    # headline `value` keeps 512 per GPU (the configuration the kernels' roofline numbers are quoted on); this leg
    # times the same step at BASELINE's per-GPU shard, where the step is launch-latency bound (36 kernels in ~0.2 ms).
    small = None
    Bs = {'cfg3': 64, 'cfg4': 128}.get(cfg)
    if Bs and Bs != B and not args.per_gpu_batch:
        xs2 = inp._buf('fit_x0', (Bs,) + X.shape[1:], 'nhwc' if X.ndim == 4 else 'flat')
        ys2 = inp._buf('fit_y0', (Bs, Y.shape[1]), 'flat')
        lib.sn_memcpy_d2d(xs2.ptr, xs.ptr, xs2.size * 4, current_stream())
        lib.sn_memcpy_d2d(ys2.ptr, ys.ptr, ys2.size * 4, current_stream())
        for _ in range(Wm):
            m._run_step(xs2, ys2, Bs * world)
        barrier()
        ev0.record()
        for _ in range(K):
            m._run_step(xs2, ys2, Bs * world)
        ev1.record()
        barrier()
        small = ev0.elapsed_time(ev1)

    # ---- instrumented pass: per-kernel CUDA-event durations (eager, same buffers) ---
    m.use_cuda_graph = False
    lib.sn_launch_count_reset()
    m._run_step(xs, ys, G)
    torch.cuda.synchronize()
    launches_per_step = int(lib.sn_launch_count())
    lib.start_profiling()
    for _ in range(min(K, 10)):
        # A few ms of device-side spinning first, so that the host has the whole eager step enqueued
        # before the GPU reaches it: otherwise entries that launch several kernels get host launch
        # gaps inside their event brackets whenever the GPU catches up with the Python loop.
        torch.cuda._sleep(int(1.2e7))
        m._run_step(xs, ys, G)
    summary = lib.stop_profiling()
    m.use_cuda_graph = True
    roof, rows, roof_t = roofline_from_profile(summary, precision, pk)

    t = torch.tensor([dev_ms, e2e_ms, small or 0.0], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, small = float(t[0]), float(t[1]), (float(t[2]) if small else None)

    if rank == 0:
        if args.kernel_table:
            os.makedirs(os.path.dirname(os.path.abspath(args.kernel_table)), exist_ok=True)
            with open(args.kernel_table, 'w') as f:
                json.dump(dict(config=cfg, per_gpu_batch=B, precision=precision, step_ms_graph=dev_ms / K, kernels=rows), f, indent=1)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            # the reference's own fit() at the benchmarked minibatch size when it can build the config,
            # else the NumPy port on a smaller batch; ~10-30 s of host time either way
            res = cpu_reference_throughput(cfg, B, {'cfg3': 2, 'cfg1': 30, 'cfg2': 100}.get(cfg, 2), 1,
                                           warmup_batch=64 if cfg == 'cfg3' else None) if B <= 512 else None
            if res is not None:
                cpu = dict(value=res[0], unit='samples/s', cores=os.cpu_count(), kind='reference',
                           sample='%d minibatches of %d through the unmodified reference Sequential.fit() (oracle/_ref), float64 NumPy'
                                  % ({'cfg3': 2, 'cfg1': 30, 'cfg2': 100}.get(cfg, 2), B))
            else:
                cpu_batch = {'cfg3': 64, 'cfg1': 256, 'cfg2': 128, 'cfg4': 32, 'cfg5': 256}[cfg]
                cpu_steps = {'cfg3': 6, 'cfg1': 30, 'cfg2': 200, 'cfg4': 3, 'cfg5': 10}[cfg]
                cpu_val, _ = cpu_oracle_throughput(cfg, cpu_batch, cpu_steps)
                cpu = dict(value=cpu_val, unit='samples/s', cores=os.cpu_count(), kind='port',
                           sample='%d steps of batch %d of the same workload, float64 NumPy oracle' % (cpu_steps, cpu_batch))
        # live parity check (N = 1, skipped with --no-cpu-baseline): the first, pre-update loss of a 64-sample minibatch
        # of this workload on a FRESH model of the benchmarked tier against the float64 oracle — the full-size
        # parity evidence is tests/test_gpu_fullsize.py (every layer shape and the trajectory at batch 512)
        parity = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import shinnosuke_oracle as O
                Xp, Yp = synth(cfg, 64, seed=4321)
                np.random.seed(0)
                spec, ishape = oracle_spec(cfg)
                net = O.OracleNet(spec, ishape)
                net.step(Xp.astype(np.float64), Yp.astype(np.float64))
                mp = build_model(cfg, precision)
                mp.fit(Xp, Yp, batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
                tol = {'fp32': 1e-5, 'tf32x3': 5e-5, 'tf32': 2e-3, 'bf16': 1e-2}[precision]
                if not np.isfinite(net.train_loss[0]):
                    raise ValueError('the oracle loss is not finite at this initialisation (softmax underflow -> log 0 in float64)')
                err = abs(mp.train_loss[0] - net.train_loss[0]) / abs(net.train_loss[0])
                parity = dict(check='first loss of a 64-sample minibatch vs the float64 oracle', loss=mp.train_loss[0],
                              oracle_loss=net.train_loss[0], rel_err=err, tolerance=tol, ok=bool(err < tol),
                              full_size_tests='tests/test_gpu_fullsize.py')
            except Exception as e:          # the bench line must not die on its own checker
                parity = dict(error=str(e)[:200])
        h2d = int(n_loc * (np.prod(X.shape[1:]) + Y.shape[1]) * 4)
        line = dict(metric='train_samples_per_sec', value=G * K / (dev_ms * 1e-3), unit='samples/s', n_gpus=world, steps=K,
                    warmup=Wm, ms_per_step=dev_ms / K, higher_is_better=True, scaling='weak', vs_baseline=None,
                    dtype=precision, data='synthetic',
                    config=dict(workload=CONFIGS[cfg]['name'], per_gpu_batch=B, global_batch=G, parallelism='dp%d' % world,
                                precision=precision, sync_bn=bool(world > 1 and not args.no_sync_bn), comm=('nccl' if os.environ.get('SN_COMM') == 'nccl' else 'own NVLink kernels') if world > 1 else None, l2='working set per step exceeds the 126 MB L2 (no flush)' if cfg in ('cfg3', 'cfg4') and B >= 128
                                else 'small working set, no flush: the step is launch-latency bound',
                                step_tflops=CONFIGS[cfg]['flops_per_sample'] * G / world / (dev_ms / K * 1e-3) / 1e12),
                    clocks=clocks,
                    e2e=dict(value=G * K / (e2e_ms * 1e-3), unit='samples/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=8,
                             ms_per_step=e2e_ms / K, final_loss=losses_e2e[-1] if losses_e2e else None),
                    gpu_launches=launches_per_step * K, roofline=roof, roofline_tensor=roof_t, cpu_baseline=cpu, parity=parity,
                    baseline_variant=dict(per_gpu_batch=Bs, global_batch=Bs * world, ms_per_step=small / K, value=Bs * world * K / (small * 1e-3),
                                          note='BASELINE.json batch (512 over 8 GPUs for config 3): same step at the per-GPU shard size') if small else None)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Captured graphs hold NCCL kernels; tearing the communicator down under them blocks in
        # destroy_process_group() (seen on torch 2.11 / NCCL 2.28).  Drain, agree, and leave.
        m._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)



# --------------------------------------------------------------- micro roofline sweep
def run_micro(args):
    """`bench.py --micro`: the HBM-bound and Dense kernels north_star sets bars for, stand-alone, at sizes
    far beyond the 126 MB L2 (so the number is HBM, not cache): max-pool forward / backward, the fused
    optimizer updates, Dense 784 -> 500 at large M.  Each kernel: 3 warm-up launches, then 10 timed ones
    bracketed by CUDA events on the launching stream; the buffers of consecutive launches rotate through
    a pool > 2 x L2 where the working set itself is smaller.  Algorithmic bytes / flops = kernel_work().
    Prints one JSON object; profiles/ holds the committed copy."""
    import torch
    torch.cuda.set_device(0)
    from shinnosuke_b200._lib import lib
    from shinnosuke_b200.device import current_stream
    pk = peaks()
    st = current_stream()
    rows = []

    def timed(name, a, fn, precision='fp32', reps=10, note=None, tensor_peak=None):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))
        flops, byt = kernel_work(name, a, precision)
        tp = tensor_peak or pk['bf16'] / 2
        row = dict(kernel=name, args=list(a), ms=ms, precision=precision)
        if flops / (tp * 1e12) > byt / (pk['hbm'] * 1e9):
            row.update(bound='tensor', achieved=flops / ms / 1e9, peak=tp, unit='TFLOP/s', frac=flops / ms / 1e9 / tp)
        else:
            row.update(bound='hbm', achieved=byt / ms / 1e6, peak=pk['hbm'], unit='GB/s', frac=byt / ms / 1e6 / pk['hbm'],
                       algorithmic_bytes=byt)
        if note:
            row['note'] = note
        rows.append(row)
        print('%-22s %-34s %8.3f ms  %9.1f %s  = %.2f of %s' % (name, a, ms, row['achieved'], row['unit'], row['frac'], row['bound']),
              file=sys.stderr, flush=True)

    f32 = dict(dtype=torch.float32, device='cuda')
    # ---- MaxPooling2D 2x2/2, tie-split codes ('reshape' mode, every BASELINE config) ----
    for N in (512, 2048):
        H = W = 32; C = 64
        x = torch.randn(N * H * W * C, **f32)
        y = torch.empty(N * (H // 2) * (W // 2) * C, **f32)
        code = torch.empty(y.numel(), dtype=torch.uint8, device='cuda')
        dx = torch.empty_like(x)
        a = (N, H, W, C, 2, 2)
        timed('sn_maxpool2d_fwd', a, lambda: lib.sn_maxpool2d_fwd(x.data_ptr(), y.data_ptr(), code.data_ptr(), N, H, W, C, 2, 2, 0, st))
        timed('sn_maxpool2d_bwd', a, lambda: lib.sn_maxpool2d_bwd(y.data_ptr(), code.data_ptr(), dx.data_ptr(), N, H, W, C, 2, 2, 0, 0, st))
        del x, y, code, dx
    # ---- optimizers over a 64 M-parameter arena (256 MB per tensor) ----
    P = 64 * 1024 * 1024
    w = torch.randn(P, **f32); g = torch.randn(P, **f32) * 1e-3
    timed('sn_sgd_update', (P,), lambda: lib.sn_sgd_update(w.data_ptr(), g.data_ptr(), P, 0.01, 1.0, st))
    v = torch.zeros(P, **f32)
    timed('sn_momentum_update', (P,), lambda: lib.sn_momentum_update(w.data_ptr(), g.data_ptr(), v.data_ptr(), None, P, 0.01, 0.9, 1.0, st))
    s2 = torch.zeros(P, **f32); tick = torch.zeros(4, **f32)
    timed('sn_adaptive_update', (0, P), lambda: lib.sn_adaptive_update(0, w.data_ptr(), g.data_ptr(), v.data_ptr(), None, None, None, P, 1e-3, 0.9, 0.0, 1e-7, 1.0, st),
          note='RMSprop')
    timed('sn_adaptive_update', (3, P), lambda: lib.sn_adaptive_update(3, w.data_ptr(), g.data_ptr(), v.data_ptr(), s2.data_ptr(), None, tick.data_ptr(), P, 1e-3, 0.9, 0.999, 1e-7, 1.0, st),
          note='Adam')
    del w, g, v, s2
    # ---- Dense 784 -> 500 (config 2's layer) at large M: fprop, dX, dW ----
    K, Nout = 784, 500
    for M in (8192, 65536):
        x = torch.randn(M * K, **f32); wt = torch.randn(Nout * K, **f32) * 0.05; b = torch.zeros(Nout, **f32)
        y = torch.empty(M * Nout, **f32); dy = torch.randn(M * Nout, **f32); dx = torch.empty(M * K, **f32)
        dw = torch.zeros(Nout * K, **f32); db = torch.zeros(Nout, **f32)
        ws_bytes = lib.load().sn_conv2d_dgrad_workspace(K, Nout, 1, 1, 1)
        ws = torch.empty(max(1, ws_bytes // 4), **f32)
        a = (M, K, Nout)
        tp = pk['bf16'] / 2
        timed('sn_dense_fprop', a, lambda: lib.sn_dense_fprop(x.data_ptr(), wt.data_ptr(), b.data_ptr(), y.data_ptr(), M, K, Nout, 1, 1, st),
              precision='tf32', tensor_peak=tp)
        timed('sn_dense_dgrad', a, lambda: lib.sn_dense_dgrad(dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), M, K, Nout, 0, 1, ws.data_ptr(), st),
              precision='tf32', tensor_peak=tp)
        timed('sn_dense_wgrad', a, lambda: lib.sn_dense_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), M, K, Nout, 1, 1, st),
              precision='tf32', tensor_peak=tp, note='K = 784 is 24.5 TMA boxes: ragged tail zero-filled, tcgen05 kind::tf32')
        # bf16 tier: what layers/FC.py runs — bf16 operand copies through the 1x1 case of the implicit-GEMM kernels
        xb = x.to(torch.bfloat16); wb = wt.to(torch.bfloat16); dyb = dy.to(torch.bfloat16)
        wtb = torch.empty(Nout * K, dtype=torch.bfloat16, device='cuda')
        lib.sn_flip_transpose_filters_bf16(wt.data_ptr(), wtb.data_ptr(), Nout, K, 1, 1, st)
        g = (M, K, 1, 1, Nout, 1, 1, 1, 0, 0)
        timed('sn_conv2d_fprop_bf16', g, lambda: lib.sn_conv2d_fprop_bf16(xb.data_ptr(), wb.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, st),
              precision='bf16', tensor_peak=pk['bf16'], note='Dense fprop, bf16 tier')
        if lib.load().sn_conv2d_bf16_supported(1, *g):
            timed('sn_conv2d_dgrad_bf16', g, lambda: lib.sn_conv2d_dgrad_bf16(dyb.data_ptr(), wtb.data_ptr(), dx.data_ptr(), *g, 0, st),
                  precision='bf16', tensor_peak=pk['bf16'], note='Dense dX, bf16 tier')
        else:
            rows.append(dict(kernel='sn_conv2d_dgrad_bf16', args=list(g), precision='bf16',
                             note='not taken: n_out = 500 is not a multiple of 8 bf16 elements (16-byte TMA rows); layers/FC.py uses the tf32 kernel for dX'))
        del xb, wb, dyb, wtb
        del x, wt, y, dy, dx, dw
    out = dict(micro=True, peaks=pk, rows=rows)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg3', choices=sorted(CONFIGS))
    ap.add_argument('--per-gpu-batch', type=int, default=0)
    ap.add_argument('--precision', default='bf16', choices=['fp32', 'tf32x3', 'tf32', 'bf16'])
    ap.add_argument('--kernel-table', default='', help='write the per-kernel CUDA-event table (JSON) here')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--micro', action='store_true', help='stand-alone roofline sweep of the pooling / optimizer / Dense kernels at > L2 sizes')
    ap.add_argument('--sync-bn', action='store_true', help='(default at N > 1) cross-replica BatchNorm statistics: exact data parallelism')
    ap.add_argument('--no-sync-bn', action='store_true', help='per-replica BatchNorm statistics (a deviation from the reference; A/B only)')
    args = ap.parse_args()
    wd = int(os.environ.get('SN_BENCH_WATCHDOG', '0'))
    if wd > 0:      # debugging aid: dump every thread's stack and exit if the run takes longer than this
        import faulthandler
        faulthandler.dump_traceback_later(wd, exit=True)
    if args.micro:
        run_micro(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
