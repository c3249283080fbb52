"""TEST / MEASUREMENT INFRASTRUCTURE — drives the UNMODIFIED reference (oracle/_ref, built by
oracle/make_ref.sh) through its own public API: Sequential / Model, compile(), fit().

Only bench.py (`--impl reference` and the `cpu_baseline` leg) and tests/ may import this; the
product (shinnosuke_b200/) never does.  Nothing here re-implements the reference's arithmetic: the
functions below only BUILD the BASELINE.json configurations out of the reference's own layers, with
the constructions SURVEY.md section 8(c) found necessary for the reference to run them at all:

  * matplotlib is stubbed (models.py:5 imports it unconditionally; it is not installed);
  * padding='SAME' is expressed as ZeroPadding2D((1,1)) + Conv2D(padding='VALID') — the reference's
    literal SAME forward mis-indexes the padded buffer (layers/Convolution.py:137-146), and the
    ZeroPadding2D + VALID pair is its own exact, correct SAME;
  * the first ZeroPadding2D of a Sequential gets an Input to connect to (ZeroPadding2D.connect
    dereferences prev_layer, layers/Convolution.py:219);
  * config 1's MaxPooling2D runs in mode='im2col' (the shipped 'reshape' mode raises on the odd
    27x27 map, :319; its own declared output_shape is what 'im2col' computes).
"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')


def available():
    return os.path.isfile(os.path.join(REF_DIR, 'shinnosuke_ref', 'models.py'))


def load():
    """Import the reference package from oracle/_ref (never from /root/reference at run time)."""
    if not available():
        raise RuntimeError('oracle/_ref is missing: run `sh oracle/make_ref.sh` where /root/reference exists')
    if 'shinnosuke_ref' not in sys.modules:
        mpl, plt = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
        mpl.pyplot = plt
        sys.modules.setdefault('matplotlib', mpl)
        sys.modules.setdefault('matplotlib.pyplot', plt)
        sys.dont_write_bytecode = True
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import shinnosuke_ref  # noqa: F401
    return sys.modules['shinnosuke_ref']


def _first_pad(ref, zp, input_shape):
    inp = ref.layers.Base.Input(input_shape)
    orig = zp.connect

    def connect(prev):
        orig(inp)
        zp.inbounds = []
    zp.connect = connect
    return zp


def build(cfg):
    """BASELINE.json configs 1-3 out of the reference's own layers (np.random.seed(0) first, like
    bench.build_model, so both arms start from the same weights)."""
    ref = load()
    L = ref.layers
    Conv2D, MaxPooling2D, ZeroPadding2D = L.Convolution.Conv2D, L.Convolution.MaxPooling2D, L.Convolution.ZeroPadding2D
    Dense, Flatten = L.FC.Dense, L.FC.Flatten
    BatchNormalization = L.Normalization.BatchNormalization
    np.random.seed(0)
    m = ref.models.Sequential()
    if cfg == 'cfg3':
        first = True
        for f in (64, 128, 256, 256):
            zp = ZeroPadding2D((1, 1))
            if first:
                zp.input_shape = (None, 3, 32, 32)
                zp = _first_pad(ref, zp, zp.input_shape)
                first = False
            m.add(zp)
            m.add(Conv2D(f, (3, 3), padding='VALID', activation='relu', initializer='normal'))
            m.add(BatchNormalization())
            m.add(MaxPooling2D((2, 2)))
        m.add(Flatten())
        m.add(Dense(10, activation='softmax'))
    elif cfg == 'cfg1':
        m.add(Conv2D(8, (2, 2), input_shape=(None, 1, 28, 28), padding='VALID', activation='relu', initializer='normal'))
        pool = MaxPooling2D((2, 2))
        m.add(pool)
        m.add(Flatten())
        m.add(Dense(10, activation='softmax', initializer='normal'))
    elif cfg == 'cfg2':
        m.add(Dense(500, activation='relu', n_in=784))
        m.add(Dense(10, activation='softmax'))
    else:
        raise NotImplementedError('the reference cannot build %s as shipped (SURVEY.md 8(c)): use the oracle port' % cfg)
    m.compile(loss='sparse_categorical_crossentropy', optimizer='sgd')
    if cfg == 'cfg1':
        pool.mode = 'im2col'
    return m


def can_build(cfg):
    return available() and cfg in ('cfg1', 'cfg2', 'cfg3')


def fit_throughput(cfg, X, Y, batch, warmup_steps, steps):
    """samples/s of the reference's own fit() over `steps` minibatches of `batch` (after
    `warmup_steps` untimed ones, a separate fit() call on the same model); returns
    (samples_per_s, seconds_per_step, losses of the timed steps)."""
    m = build(cfg)
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):            # the reference prints a progress bar per minibatch
        if warmup_steps:
            m.fit(X[:batch * warmup_steps], Y[:batch * warmup_steps], batch_size=batch, epochs=1, shuffle=False,
                  validation_ratio=0.)
        a, b = batch * warmup_steps, batch * (warmup_steps + steps)
        n0 = len(m.train_loss)
        t0 = time.perf_counter()
        m.fit(X[a:b], Y[a:b], batch_size=batch, epochs=1, shuffle=False, validation_ratio=0.)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, [float(v) for v in m.train_loss[n0:]]
