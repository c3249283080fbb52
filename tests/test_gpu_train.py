"""-m gpu: whole training trajectories through Sequential/Model.compile/fit on the device against
the live-reference fixtures (SURVEY.md section 9 KATs C and D and the VGG-style config) and the
CPU oracle; plus size-independent properties at BASELINE.json's full sizes."""
import os
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import shinnosuke_oracle as O

pytestmark = pytest.mark.gpu
TOL32 = 1e-5


def _data(shape, kind='random'):
    r = np.random.default_rng(1234)
    X = (r.random(shape) if kind == 'random' else r.standard_normal(shape)).astype(np.float32).astype(np.float64)
    return r, X


def _mlp(hidden=500, optimizer='sgd', precision='fp32'):
    from shinnosuke_b200 import Sequential, Dense
    np.random.seed(0)
    m = Sequential()
    m.add(Dense(hidden, activation='relu', n_in=784))
    m.add(Dense(10, activation='softmax'))
    m.compile(loss='sparse_categorical_crossentropy', optimizer=optimizer, precision=precision)
    return m


@pytest.mark.parametrize('graph', [False, True])
def test_trajectory_cfg2_kat_c(graph):
    g = load_golden('traj_cfg2')
    r, X = _data((384, 784))
    Y = O.to_categorical(r.integers(0, 10, 384))
    m = _mlp()
    m.use_cuda_graph = graph
    m.fit(X, Y, batch_size=128, epochs=2 if graph else 1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.allclose(m.train_loss[:3], g['train_loss'], rtol=TOL32, atol=0)
    assert np.allclose(m.train_loss[:3], [3.797970846920, 3.140630456196, 2.993179112976], rtol=TOL32)
    assert list(m.train_acc[:3]) == [0.1484375, 0.109375, 0.09375]
    if not graph:
        assert abs(np.asarray(m.layers[0].variables[0].output_tensor).sum() - 85.568083440698) < 1e-3
        assert rel_err(np.asarray(m.layers[1].variables[0].output_tensor), g['W2']) < TOL32
    else:
        # epoch 2 ran from captured graphs; compare with the oracle continuing the same trajectory
        np.random.seed(0)
        spec, ishape = O.spec_cfg2()
        net = O.OracleNet(spec, ishape)
        net.fit(X, Y, 128, 2, shuffle=False)
        assert np.allclose(m.train_loss, net.train_loss, rtol=2e-5)
        assert rel_err(np.asarray(m.layers[1].variables[0].output_tensor), net.params[2]) < 2e-5


def test_trajectory_momentum_shuffled():
    """get_batches order (reseeded per epoch) + Momentum's never-zeroed gradients."""
    from shinnosuke_b200.utils.Optimizers import Momentum
    g = load_golden('traj_mlp_momentum')
    r, X = _data((384, 784))
    Y = O.to_categorical(r.integers(0, 10, 384))
    m = _mlp(hidden=64, optimizer=Momentum(lr=0.01))
    m.fit(X[:200], Y[:200], batch_size=64, epochs=2, shuffle=True, validation_ratio=0., verbose=0)
    assert len(m.train_loss) == 8                                   # 3 full + 1 ragged batch per epoch
    assert np.allclose(m.train_loss, g['train_loss'], rtol=2e-5)
    assert rel_err(np.asarray(m.layers[1].variables[0].output_tensor), g['W2']) < 2e-5


def test_trajectory_cfg1_kat_d():
    from shinnosuke_b200 import Model, Input, Conv2D, MaxPooling2D, Flatten, Dense
    g = load_golden('traj_cfg1')
    r, X = _data((768, 1, 28, 28))
    Y = O.to_categorical(r.integers(0, 10, 768))
    np.random.seed(0)
    X_in = Input(shape=(None, 1, 28, 28))
    h = Conv2D(8, (2, 2), padding='VALID', initializer='normal', activation='relu')(X_in)
    pool = MaxPooling2D((2, 2))(h)
    pool.mode = 'im2col'            # what the fixture forced on the reference (its 'reshape' raises on 27x27)
    h = Flatten()(pool)
    y = Dense(10, initializer='normal', activation='softmax')(h)
    model = Model(inputs=X_in, outputs=y)
    model.compile(optimizer='sgd', loss='sparse_categorical_crossentropy')
    model.fit(X, Y, batch_size=256, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.allclose(model.train_loss, g['train_loss'], rtol=TOL32)
    assert np.allclose(model.train_loss, [4.071828044837, 2.796486836493, 2.469934709778], rtol=TOL32)
    assert np.allclose(model.train_acc, [0.093750, 0.164062, 0.078125], atol=1e-6)
    conv = model.forward_graph[1]
    assert rel_err(np.asarray(conv.variables[0].output_tensor), g['convW']) < TOL32
    assert rel_err(np.asarray(conv.variables[1].output_tensor), g['convb']) < TOL32
    # default mode ('reshape' = tie split, floor on the odd map) gives the same trajectory: the
    # only ties are all-blocked ReLU windows
    np.random.seed(0)
    X_in = Input(shape=(None, 1, 28, 28))
    h = Conv2D(8, (2, 2), padding='VALID', initializer='normal', activation='relu')(X_in)
    h = Flatten()(MaxPooling2D((2, 2))(h))
    y = Dense(10, initializer='normal', activation='softmax')(h)
    m2 = Model(inputs=X_in, outputs=y)
    m2.compile(optimizer='sgd', loss='sparse_categorical_crossentropy')
    m2.fit(X, Y, batch_size=256, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.allclose(m2.train_loss, g['train_loss'], rtol=TOL32)


def _vgg(precision='fp32'):
    from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
    np.random.seed(0)
    m = Sequential()
    first = True
    for f in (64, 128, 256, 256):
        kw = dict(input_shape=(None, 3, 32, 32)) if first else {}
        first = False
        m.add(Conv2D(f, (3, 3), padding='SAME', activation='relu', initializer='normal', **kw))
        m.add(BatchNormalization())
        m.add(MaxPooling2D((2, 2)))
    m.add(Flatten())
    m.add(Dense(10, activation='softmax'))
    m.compile(loss='sparse_categorical_crossentropy', optimizer='sgd', precision=precision)
    return m


def test_trajectory_cfg3_vgg_small_batch():
    """BASELINE config 3 at full widths, batch 8, 2 SGD steps, against the reference run with
    ZeroPadding2D+VALID as its (only correct) SAME."""
    g = load_golden('traj_cfg3_small')
    r, X = _data((16, 3, 32, 32), kind='normal')
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 6)]))
    m = _vgg()
    m.fit(X, Y, batch_size=8, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.allclose(m.train_loss, g['train_loss'], rtol=5e-5)
    assert list(m.train_acc) == list(g['train_acc'])
    convs = [l for l in m.layers if l.__class__.__name__ == 'Conv2D']
    bns = [l for l in m.layers if l.__class__.__name__ == 'BatchNormalization']
    assert rel_err(np.asarray(convs[0].variables[0].output_tensor), g['conv1W']) < TOL32
    assert rel_err(np.asarray(bns[0].variables[0].output_tensor), g['bn1_gamma']) < TOL32
    assert rel_err(bns[0].moving_mean, g['bn1_mm']) < TOL32
    assert rel_err(bns[3].moving_variance, g['bn4_mv']) < TOL32
    assert rel_err(np.asarray(m.layers[-1].variables[0].output_tensor), g['denseW']) < TOL32


def test_evaluate_predict_and_pickle_roundtrip(tmp_path):
    r, X = _data((256, 784))
    Y = O.to_categorical(r.integers(0, 10, 256))
    m = _mlp(hidden=32)
    m.fit(X, Y, batch_size=64, epochs=1, shuffle=True, validation_ratio=0., verbose=0)
    acc, loss = m.evaluate(X, Y, batch_size=100)                    # ragged last chunk
    p = np.asarray(m.predict(X[:10], is_training=False))
    assert p.shape == (10, 10) and np.allclose(p.sum(1), 1, atol=1e-5)
    # oracle with the trained weights
    W1, b1, W2, b2 = [np.asarray(v.output_tensor) for v in m.trainable_variables]
    h, _ = O.dense_forward(X, W1, b1, 'relu')
    ph, _ = O.dense_forward(h, W2, b2, 'softmax')
    chunks = [(0, 100), (100, 200), (200, 256)]
    ref_loss = np.mean([O.scce_loss(ph[a:b], Y[a:b]) for a, b in chunks])
    ref_acc = np.mean([O.scce_acc(ph[a:b], Y[a:b]) for a, b in chunks])
    assert abs(loss - ref_loss) < 1e-5 * ref_loss and abs(acc - ref_acc) < 1e-9
    path = str(tmp_path / 'mlp')
    m.save(path)
    from shinnosuke_b200 import Sequential
    m2 = Sequential()
    m2.load(path)
    acc2, loss2 = m2.evaluate(X, Y, batch_size=100)
    assert acc2 == acc and abs(loss2 - loss) < 1e-6
    # training continues after a load
    m2.fit(X, Y, batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert len(m2.train_loss) == 4 and np.isfinite(m2.train_loss).all()


def test_validation_split_default_like_reference():
    r, X = _data((100, 784))
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 90)]))
    m = _mlp(hidden=16)
    m.fit(X, Y, batch_size=45, epochs=1, shuffle=False, verbose=0)   # validation_ratio defaults to 0.1
    assert len(m.train_loss) == 2 and len(m.valid_loss) == 2         # validated after every minibatch


def test_full_size_cfg3_properties():
    """BASELINE config 3 at the bench's per-GPU batch: properties that need no oracle run.
    (1) shard linearity: the gradient of the full batch equals the sum of the gradients of two
    half shards when each divides by the global row count (what data parallelism relies on);
    (2) SGD idempotence: g == 0 after an update, and a second update with g == 0 is a no-op."""
    from shinnosuke_b200.device import as_device
    N = 128
    r, X = _data((N, 3, 32, 32), kind='normal')
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, N - 10)]))
    # BN couples the rows of a batch, so linearity is checked on the BN-free sub-network
    from shinnosuke_b200 import Sequential, Conv2D, MaxPooling2D, Flatten, Dense
    def build():
        np.random.seed(0)
        m = Sequential()
        m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
        m.add(MaxPooling2D((2, 2)))
        m.add(Conv2D(128, (3, 3), padding='SAME', activation='relu'))
        m.add(MaxPooling2D((2, 2)))
        m.add(Flatten()); m.add(Dense(10, activation='softmax'))
        m.compile('sgd', 'sparse_categorical_crossentropy')
        m._materialize()
        return m
    def grads(m, xs, ys, denom):
        m._arena.g.zero_()
        y_hat = m.predict(xs)
        m.calc_gradients(y_hat, ys, denom_rows=denom)
        return m._arena.g[:m._arena.param_floats].double().cpu().numpy()
    m = build()
    full = grads(m, X, Y, N)
    halves = grads(m, X[:N // 2], Y[:N // 2], N) + grads(m, X[N // 2:], Y[N // 2:], N)
    assert rel_err(halves, full) < TOL32
    m.optimizer.update(m.trainable_variables)
    assert not m._arena.g[:m._arena.param_floats].any().item()
    w = m._arena.w.clone()
    m.optimizer.update(m.trainable_variables)
    assert (m._arena.w == w).all().item()


def test_fused_and_unfused_models_agree():
    """compile(fuse=True) (BatchNormalization+MaxPooling2D as one pass, conv-epilogue handoffs, Flatten
    as a view: the default) and fuse=False (one kernel per layer, every layer's output_tensor
    materialised) train identically.  The model is kept well conditioned (initial loss ~3, not the
    ~25 a Dense(std 0.1) over 8192 features starts from): at loss 25 a single ReLU / max-pool
    decision flipped by a 1-ulp difference — fp32 atomics accumulate in arrival order — moves the
    weights by 1e-2 within four steps, in BOTH variants and against the float64 oracle alike."""
    r, X = _data((32, 3, 16, 16), kind='normal')
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 22)]))
    from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
    runs = []
    for fuse in (True, False):
        np.random.seed(0)
        m = Sequential()
        m.add(Conv2D(16, (3, 3), input_shape=(None, 3, 16, 16), padding='SAME', activation='relu'))
        m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
        m.add(Conv2D(32, (3, 3), padding='SAME', activation='relu'))
        m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
        m.add(Flatten()); m.add(Dense(10, activation='softmax'))
        m.compile('sgd', 'sparse_categorical_crossentropy', fuse=fuse)
        m.fit(X, Y, batch_size=16, epochs=2, shuffle=True, validation_ratio=0., verbose=0)
        runs.append((m.train_loss, [np.asarray(v.output_tensor) for v in m.trainable_variables]))
        assert (m.layers[1].output_tensor is None) == fuse
    # and both against the oracle (whole trajectory, every variable)
    np.random.seed(0)
    spec = [('conv', 16, 3, 1, 'SAME', 'relu'), ('bn',), ('pool', 2), ('conv', 32, 3, 1, 'SAME', 'relu'), ('bn',), ('pool', 2),
            ('flatten',), ('dense', 10, 'softmax')]
    net = O.OracleNet(spec, (None, 3, 16, 16), optimizer='sgd')
    net.fit(X, Y, 16, epochs=2, shuffle=True)
    assert net.train_loss[0] < 8
    for losses, ws in runs:
        assert np.allclose(losses, net.train_loss, rtol=2e-5)
        for w, p in zip(ws, net.params):
            assert rel_err(w, p) < 2e-5
    assert np.allclose(runs[0][0], runs[1][0], rtol=1e-5)


def _residual_model(width, k, padding, in_shape, precision='fp32'):
    from shinnosuke_b200 import Model, Input, Conv2D, BatchNormalization, Activation, Add, MaxPooling2D, Flatten, Dense
    np.random.seed(0)
    X_in = Input(shape=(None,) + in_shape)
    stem = Conv2D(width, (k, k), padding=padding, activation='relu', initializer='normal')(X_in)
    r = Conv2D(width, (k, k), padding=padding, activation='linear', initializer='normal')(stem)
    r = BatchNormalization()(r)
    r = Activation('relu')(r)
    r = Conv2D(width, (k, k), padding=padding, activation='linear', initializer='normal')(r)
    r = BatchNormalization()(r)
    h = Add()([stem, r])
    h = Activation('relu')(h)
    h = MaxPooling2D((2, 2))(h)
    h = Flatten()(h)
    y = Dense(10, initializer='normal', activation='softmax')(h)
    model = Model(inputs=X_in, outputs=y)
    model.compile(optimizer='sgd', loss='sparse_categorical_crossentropy', precision=precision)
    return model, stem, y


def test_trajectory_residual_add_vs_reference():
    """BASELINE config 4's building block through the functional API (Add), against the fixture the
    live reference produced (1x1 VALID convolutions, see tests/golden/make_golden.py)."""
    g = load_golden('traj_residual')
    r, X = _data((48, 3, 8, 8), kind='normal')
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 38)]))
    model, stem, y = _residual_model(8, 1, 'VALID', (3, 8, 8))
    assert [l.__class__.__name__ for l in model.forward_graph].count('Add') == 1
    assert sum(v.size for v in model.trainable_variables) == 3 * 8 + 8 + 2 * (64 + 8) + 4 * 8 + 128 * 10 + 10
    model.fit(X, Y, batch_size=16, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.allclose(model.train_loss, g['train_loss'], rtol=2e-5)
    assert rel_err(np.asarray(stem.variables[0].output_tensor), g['stemW']) < TOL32
    assert rel_err(np.asarray(y.variables[0].output_tensor), g['denseW']) < TOL32


def test_trajectory_residual_3x3_same_vs_oracle():
    r, X = _data((32, 3, 16, 16), kind='normal')
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 22)]))
    np.random.seed(0)
    net = O.OracleNet(O.spec_residual(32) + [('flatten',), ('dense', 10, 'softmax')], (None, 3, 16, 16))
    net.fit(X, Y, 16, 2, shuffle=True)
    model, stem, y = _residual_model(32, 3, 'SAME', (3, 16, 16))
    model.fit(X, Y, batch_size=16, epochs=2, shuffle=True, validation_ratio=0., verbose=0)
    assert np.allclose(model.train_loss, net.train_loss, rtol=5e-5)
    assert rel_err(np.asarray(stem.variables[0].output_tensor), net.params[0]) < 2e-5


def test_no_uninitialised_reads_and_no_allocation_inside_graph_capture():
    """SN_DEBUG_POISON=1 fills every fresh device buffer with NaN and refuses allocations while a CUDA
    graph is being captured (a captured step holds raw addresses).  A run with a ragged last
    minibatch — two shapes alternating, both captured — must stay finite, allocate nothing during
    capture, and still land on the oracle."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import shinnosuke_oracle as O
from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
r = np.random.default_rng(9)
X = r.standard_normal((20, 3, 16, 16)).astype(np.float32).astype(np.float64)
Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 10)]))
for opt in ('sgd', 'momentum', 'adam'):
    np.random.seed(0)
    m = Sequential()
    m.add(Conv2D(16, (3, 3), input_shape=(None, 3, 16, 16), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Flatten()); m.add(Dense(10, activation='softmax'))
    m.compile(opt, 'sparse_categorical_crossentropy')
    m.fit(X, Y, batch_size=8, epochs=4, shuffle=True, validation_ratio=0., verbose=0)     # batches 8, 8, 4 x 4 epochs
    assert np.all(np.isfinite(m.train_loss)), m.train_loss
    if opt == 'sgd':
        np.random.seed(0)
        net = O.OracleNet([('conv', 16, 3, 1, 'SAME', 'relu'), ('bn',), ('pool', 2), ('flatten',), ('dense', 10, 'softmax')],
                          (None, 3, 16, 16), optimizer='sgd')
        net.fit(X, Y, 8, epochs=4, shuffle=True)
        assert np.allclose(m.train_loss, net.train_loss, rtol=5e-5), (m.train_loss, net.train_loss)
print('POISON_OK')
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SN_DEBUG_POISON='1')
    out = subprocess.run([sys.executable, '-c', code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'POISON_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
