"""Pins oracle/shinnosuke_oracle.py to outputs of the live reference
(tests/golden/*.npz, made by tests/golden/make_golden.py) and to the
known-answer values of SURVEY.md §9.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import shinnosuke_oracle as O

TOL = 1e-12   # float64 restatement vs float64 reference (summation order differs)


@pytest.mark.parametrize('name', ['conv_katA', 'conv_valid_s2', 'conv_cfg1_shape', 'conv_c64_same', 'conv_5x5_same'])
def test_conv_matches_reference(name):
    g = load_golden(name)
    N, C, H, W, F, k, stride, pad = [int(v) for v in g['meta']]
    act = str(g['act'])
    y, cache = O.conv2d_forward(g['X'], g['W'], g['b'], stride, (pad, pad), act)
    assert rel_err(y, g['Y']) < TOL
    dx, dw, db = O.conv2d_backward(g['G'], cache)
    assert rel_err(dw, g['dW']) < TOL
    assert rel_err(db.reshape(1, -1), g['db']) < TOL
    assert rel_err(dx, g['dX']) < TOL
    # independent loop-nest check of the forward
    z = O.conv2d_naive(g['X'], g['W'], g['b'], stride, (pad, pad))
    assert rel_err(O.apply_activation(z, act), g['Y']) < TOL


def test_conv_katA_survey_values():
    """SURVEY.md §9 KAT A (weights come from np.random.seed(0))."""
    g = load_golden('conv_katA')
    np.random.seed(0)
    W = np.random.normal(0, 0.1, (4, 3, 3, 3))
    b = np.random.randn(1, 4)
    assert np.allclose(W[0, 0, 0, :], [0.176405234597, 0.040015720837, 0.097873798411], atol=1e-12)
    assert np.allclose(b[0, :2], [1.92294202648, 1.480514791434], atol=1e-11)
    r = np.random.default_rng(1234)
    X = r.standard_normal((2, 3, 8, 8)).astype(np.float32).astype(np.float64)
    G = r.standard_normal((2, 4, 8, 8)).astype(np.float32).astype(np.float64)
    y, cache = O.conv2d_forward(X, W, b, 1, (1, 1), 'relu')
    dx, dw, db = O.conv2d_backward(G, cache)
    assert abs(y.sum() - 788.215565747495) < 1e-8
    assert abs(y[1, 2, 3, 4] - 1.935920418301) < 1e-10
    assert abs(np.abs(dw).sum() - 886.558538922629) < 1e-8
    assert abs(db.sum() - 15.266000591917) < 1e-9
    # the survey read dX from the ZeroPadding2D layer's .grads, i.e. INCLUDING the halo
    Xp = np.pad(X, ((0, 0), (0, 0), (1, 1), (1, 1)))
    _, cache_p = O.conv2d_forward(Xp, W, b, 1, (0, 0), 'relu')
    dxp, _, _ = O.conv2d_backward(G, cache_p)
    assert abs(np.abs(dxp).sum() - 222.887959045916) < 1e-8
    assert np.allclose(dxp[:, :, 1:-1, 1:-1], dx, atol=1e-14)


@pytest.mark.parametrize('name', ['pool_katB_reshape', 'pool_katB_im2col', 'pool_odd27_im2col', 'pool_32_reshape',
                                  'pool_3x3_ties'])
def test_pool_matches_reference(name):
    g = load_golden(name)
    y, cache = O.maxpool_forward(g['X'], int(g['pool']), None, str(g['mode']))
    assert np.array_equal(y, g['Y'])
    if 'argmax' in g:
        assert cache['argmax'].dtype == np.int64
        assert np.array_equal(cache['argmax'], g['argmax'])
    dx = O.maxpool_backward(g['G'], cache)
    assert rel_err(dx, g['dX']) < 1e-15


def test_pool_katB_survey_values():
    g = load_golden('pool_katB_im2col')
    assert g['Y'].sum() == 138.0
    assert list(g['argmax'][:24]) == [0, 0, 0, 2, 0, 0, 0, 0, 0, 1, 2, 1, 1, 1, 0, 0, 1, 1, 0, 0, 0, 3, 1, 2]
    assert g['argmax'].sum() == 45
    _, c1 = O.maxpool_forward(g['X'], 2, None, 'im2col')
    _, c2 = O.maxpool_forward(g['X'], 2, None, 'reshape')
    G32 = g['G'].astype(np.float32).astype(np.float64)      # the survey cast every input through float32
    d1, d2 = O.maxpool_backward(G32, c1), O.maxpool_backward(G32, c2)
    assert abs(np.abs(d1).sum() - 39.616064611357) < 1e-9 and np.count_nonzero(d1) == 54
    assert abs(np.abs(d2).sum() - 39.616064611357) < 1e-9 and np.count_nonzero(d2) == 100


@pytest.mark.parametrize('name', ['dense_relu', 'dense_softmax', 'dense_linear'])
def test_dense_matches_reference(name):
    g = load_golden(name)
    y, cache = O.dense_forward(g['X'], g['W'], g['b'], str(g['act']))
    assert rel_err(y, g['Y']) < TOL
    dx, dw, db = O.dense_backward(g['G'], cache)
    assert rel_err(dw, g['dW']) < TOL and rel_err(db, g['db']) < TOL and rel_err(dx, g['dX']) < TOL


@pytest.mark.parametrize('name', ['bn_4d', 'bn_2d'])
def test_batchnorm_matches_reference(name):
    g = load_golden(name)
    C = g['gamma'].shape[0]
    y, cache, mm, mv = O.batchnorm_forward(g['X'], g['gamma'], g['beta'], np.zeros(C), np.ones(C))
    assert rel_err(y, g['Y']) < TOL
    assert rel_err(mm, g['moving_mean']) < TOL and rel_err(mv, g['moving_var']) < TOL
    dx, dg, db = O.batchnorm_backward(g['G'], cache)
    assert rel_err(dx, g['dX']) < 1e-10
    assert rel_err(dg, g['dgamma']) < TOL and rel_err(db, g['dbeta']) < TOL
    y2, _, _, _ = O.batchnorm_forward(g['X2'], g['gamma'], g['beta'], mm, mv, is_training=False)
    assert rel_err(y2, g['Y2']) < TOL


def test_loss_matches_reference():
    g = load_golden('loss_scce')
    p = O.softmax_forward(g['Z'])
    assert rel_err(p, g['P']) < TOL
    assert abs(O.scce_loss(p, g['Y']) - float(g['loss'])) < 1e-12
    assert O.scce_acc(p, g['Y']) == float(g['acc'])
    assert rel_err(O.scce_backward(p, g['Y']), g['dZ']) < TOL


def test_optimizers_match_reference():
    g = load_golden('optimizers')
    for name in ('sgd', 'momentum'):
        params = [g['w0'].copy(), g['w1'].copy()]
        grads = [np.zeros_like(p) for p in params]
        vel = [np.zeros_like(p) for p in params]
        for t in range(3):
            for i in range(2):
                grads[i] = grads[i] + g['g%d_t%d' % (i, t)]
            if name == 'sgd':
                O.sgd_update(params, grads, lr=0.05)
            else:
                O.momentum_update(params, grads, vel, lr=0.05, rho=0.9)
            for i in range(2):
                assert rel_err(params[i], g['%s_w%d_t%d' % (name, i, t)]) < TOL
                assert rel_err(grads[i], g['%s_g%d_t%d' % (name, i, t)]) < TOL or not grads[i].any()


def test_adaptive_optimizers_match_reference():
    g = load_golden('optimizers_adaptive')
    for name in ('rmsprop', 'adagrad', 'adadelta', 'adam'):
        params = [g['w0'].copy(), g['w1'].copy()]
        grads = [np.zeros_like(p) for p in params]
        s1 = [np.zeros_like(p) for p in params]
        s2 = [np.zeros_like(p) for p in params]
        it = 0
        for t in range(4):
            for i in range(2):
                grads[i] = grads[i] + g['g%d_t%d' % (i, t)]          # never zeroed by these optimizers
            if name == 'rmsprop':
                O.rmsprop_update(params, grads, s1, lr=0.01, rho=0.9)
            elif name == 'adagrad':
                O.adagrad_update(params, grads, s1, lr=0.05)
            elif name == 'adadelta':
                O.adadelta_update(params, grads, s1, s2, rho=0.95)
            else:
                it += 1
                O.adam_update(params, grads, s1, s2, it, lr=0.01, beta1=0.9, beta2=0.999)
                it += 1                                              # the base class's increment
            for i in range(2):
                assert rel_err(params[i], g['%s_w%d_t%d' % (name, i, t)]) < TOL, (name, t)
    assert int(g['adam_iterations']) == 8 and int(g['rmsprop_iterations']) == 4


def test_get_batches_matches_reference():
    g = load_golden('get_batches')
    X = np.arange(46, dtype=np.float64).reshape(23, 2)
    Y = np.arange(23, dtype=np.float64).reshape(23, 1)
    for epoch in (0, 1):
        mb = O.get_batches(X, Y, 5, epoch, True)
        assert len(mb) == int(g['e%d_n' % epoch])
        assert np.array_equal(np.concatenate([b[1] for b in mb]), g['e%d_y' % epoch])
        assert [b[0].shape[0] for b in mb] == list(g['e%d_sizes' % epoch])


def test_convcol_matches_reference():
    g = load_golden('convcol')
    assert np.array_equal(O.im2col(g['x'], (3, 2), (3, 3), 2), g['col'])
    assert rel_err(O.col2im((2, 3, 6, 6), (1, 1), (3, 3), 1, g['dcol']), g['back']) < 1e-15


def _data(shape, n_cls=10, kind='random'):
    r = np.random.default_rng(1234)
    X = (r.random(shape) if kind == 'random' else r.standard_normal(shape)).astype(np.float32).astype(np.float64)
    return r, X


def test_trajectory_cfg2():
    """SURVEY.md §9 KAT C + live fixture."""
    g = load_golden('traj_cfg2')
    r, X = _data((384, 784))
    Y = O.to_categorical(r.integers(0, 10, 384))
    assert abs(X.sum() - float(g['xsum'])) < 1e-6, 'default_rng stream differs from the fixture'
    np.random.seed(0)
    spec, ishape = O.spec_cfg2()
    net = O.OracleNet(spec, ishape)
    net.fit(X, Y, 128, 1, shuffle=False)
    assert np.allclose(net.train_loss, g['train_loss'], rtol=0, atol=1e-10)
    assert np.allclose(net.train_loss, [3.797970846920, 3.140630456196, 2.993179112976], atol=1e-10)
    assert list(net.train_acc) == [0.1484375, 0.109375, 0.09375]
    assert abs(net.params[0].sum() - 85.568083440698) < 1e-8
    assert rel_err(net.params[2], g['W2']) < 1e-10


def test_trajectory_momentum_shuffled():
    g = load_golden('traj_mlp_momentum')
    r, X = _data((384, 784))
    Y = O.to_categorical(r.integers(0, 10, 384))
    np.random.seed(0)
    net = O.OracleNet([('dense', 64, 'relu'), ('dense', 10, 'softmax')], (None, 784), optimizer='momentum')
    net.fit(X[:200], Y[:200], 64, 2, shuffle=True)
    assert np.allclose(net.train_loss, g['train_loss'], rtol=0, atol=1e-10)
    assert rel_err(net.params[2], g['W2']) < 1e-10


def test_trajectory_cfg1():
    """SURVEY.md §9 KAT D."""
    g = load_golden('traj_cfg1')
    r, X = _data((768, 1, 28, 28))
    Y = O.to_categorical(r.integers(0, 10, 768))
    np.random.seed(0)
    spec, ishape = O.spec_cfg1()
    net = O.OracleNet(spec, ishape)
    net.fit(X, Y, 256, 1, shuffle=False)
    assert np.allclose(net.train_loss, g['train_loss'], rtol=0, atol=1e-10)
    assert np.allclose(net.train_loss, [4.071828044837, 2.796486836493, 2.469934709778], atol=1e-10)
    assert np.allclose(net.train_acc, [0.093750, 0.164062, 0.078125], atol=1e-6)
    assert rel_err(net.params[0], g['convW']) < 1e-10 and rel_err(net.params[1], g['convb']) < 1e-10


def test_trajectory_cfg3_small():
    g = load_golden('traj_cfg3_small')
    r, X = _data((16, 3, 32, 32), kind='normal')
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 6)]))
    assert abs(X.sum() - float(g['xsum'])) < 1e-6
    np.random.seed(0)
    spec, ishape = O.spec_cfg3()
    net = O.OracleNet(spec, ishape)
    net.fit(X, Y, 8, 1, shuffle=False)
    assert np.allclose(net.train_loss, g['train_loss'], rtol=1e-9, atol=0)
    assert list(net.train_acc) == list(g['train_acc'])
    assert rel_err(net.params[0], g['conv1W']) < 1e-9
    assert rel_err(net.params[2], g['bn1_gamma']) < 1e-9
    assert rel_err(net.state[1]['mm'], g['bn1_mm']) < 1e-9
    assert rel_err(net.state[10]['mv'], g['bn4_mv']) < 1e-9
    assert rel_err(net.params[-2], g['denseW']) < 1e-9


def test_trajectory_residual():
    """Functional residual graph (Add) against the live reference (with its two workarounds)."""
    g = load_golden('traj_residual')
    r, X = _data((48, 3, 8, 8), kind='normal')
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 38)]))
    assert abs(X.sum() - float(g['xsum'])) < 1e-6
    np.random.seed(0)
    spec = O.spec_residual(8, k=1, padding='VALID') + [('flatten',), ('dense', 10, 'softmax')]
    net = O.OracleNet(spec, (None, 3, 8, 8))
    net.fit(X, Y, 16, 1, shuffle=False)
    assert np.allclose(net.train_loss, g['train_loss'], rtol=1e-9, atol=0)
    assert rel_err(net.params[0], g['stemW']) < 1e-9
    assert rel_err(net.params[-2], g['denseW']) < 1e-9


def test_embedding_matches_reference():
    g = load_golden('embedding')
    assert np.array_equal(O.embedding_forward(g['W'], g['ids']), g['Y'])
    assert rel_err(O.embedding_backward(g['W'].shape, g['ids'], g['G']), g['dW']) < 1e-15


@pytest.mark.parametrize('kind', ['seq', 'last'])
def test_lstm_matches_reference(kind):
    """Forward: bit for bit.  Backward: the reference's own numbers come out of the oracle's
    reference_gate_order=True branch (gate gradients concatenated as (o, c~, u, f), Recurrent.py:358)."""
    g = load_golden('lstm')
    X, W, b, G = g['X_' + kind], g['W_' + kind], g['b_' + kind], g['G_' + kind]
    hs, cache = O.lstm_forward(X, W, b)
    y = hs if kind == 'seq' else hs[:, -1]
    assert rel_err(y, g['Y_' + kind]) < 1e-15
    dx, dw, db = O.lstm_backward(G, cache, kind == 'seq', reference_gate_order=True)
    assert rel_err(dw, g['dW_ref_' + kind]) < 1e-13
    assert rel_err(db, g['db_ref_' + kind]) < 1e-13
    assert rel_err(dx, g['dX_ref_' + kind]) < 1e-13
    # ... and they are NOT the gradient of the forward pass
    dx2, dw2, db2 = O.lstm_backward(G, cache, kind == 'seq')
    assert rel_err(dw2, g['dW_ref_' + kind]) > 1e-2


@pytest.mark.parametrize('return_sequences', [True, False])
def test_lstm_backward_is_the_numeric_gradient(return_sequences):
    """The oracle's default LSTM backward (forward gate order) against central differences."""
    r = np.random.default_rng(0)
    N, T, D, H = 3, 5, 4, 6
    x = r.standard_normal((N, T, D)); w = 0.5 * r.standard_normal((H + D, 4 * H)); b = r.standard_normal((1, 4 * H))
    g = r.standard_normal((N, T, H)) if return_sequences else r.standard_normal((N, H))

    def loss():
        hs, _ = O.lstm_forward(x, w, b)
        return float(np.sum((hs if return_sequences else hs[:, -1]) * g))

    _, cache = O.lstm_forward(x, w, b)
    dx, dw, db = O.lstm_backward(g, cache, return_sequences)
    eps = 1e-6
    for arr, grad in ((x, dx), (w, dw), (b, db)):
        num = np.zeros_like(arr)
        it = np.nditer(arr, flags=['multi_index'])
        for _ in it:
            i = it.multi_index
            old = arr[i]
            arr[i] = old + eps; lp = loss()
            arr[i] = old - eps; lm = loss()
            arr[i] = old
            num[i] = (lp - lm) / (2 * eps)
        assert rel_err(grad, num) < 1e-7


def test_unmodified_reference_fit_matches_the_oracle_trajectory():
    """oracle/_ref (the reference package copied by oracle/make_ref.sh; what `bench.py --impl reference`
    times) trained through its own Sequential.fit() on bench.py's synthetic config-3 data gives the
    oracle's losses to float64 rounding: the port and the real thing are interchangeable as the
    checker.  Skipped where oracle/_ref has not been built."""
    import pytest
    from oracle import ref_driver
    if not ref_driver.available():
        pytest.skip('oracle/_ref not built (sh oracle/make_ref.sh)')
    import bench
    from oracle import shinnosuke_oracle as O
    X, Y = bench.synth('cfg3', 24)
    _, _, losses = ref_driver.fit_throughput('cfg3', X, Y, 8, 0, 3)
    np.random.seed(0)
    spec, ishape = O.spec_cfg3()
    net = O.OracleNet(spec, ishape)
    net.fit(X.astype(np.float64), Y.astype(np.float64), 8, 1, shuffle=False)
    assert np.allclose(losses, net.train_loss, rtol=1e-12, atol=0)


def test_config3_gradient_is_ill_conditioned():
    """Evidence for the bounds in tests/test_gpu_fullsize.py::test_trajectory_bench_config_at_batch_512:
    the float64 oracle's own config-3 gradient is discontinuous in its inputs at the scale of a
    reduced-precision rounding.  Truncating only the input and the convolution filters to TF32's 10
    mantissa bits (everything else stays float64) moves the conv / BatchNorm gradients by more than
    2 % (max|a - r| / max|r| per tensor), while the loss moves by less than 1e-3 and the classifier's
    gradient by less than 2e-2: ReLU masks and max-pool winners flip, and each BatchNormalization
    backward amplifies what is left.  No reduced-precision implementation can match those gradients
    tighter than this; their parity is checked per operator with the implementation's own masks."""
    import bench

    def trunc(a, bits):
        u = a.astype(np.float32).view(np.uint32) & np.uint32((0xFFFFFFFF << (23 - bits)) & 0xFFFFFFFF)
        return u.view(np.float32).astype(np.float64)
    X, Y = bench.synth('cfg3', 32)

    def grads(bits):
        np.random.seed(0)
        spec, ishape = O.spec_cfg3()
        net = O.OracleNet(spec, ishape)
        net.params[-2] *= 0.1
        x = X.astype(np.float64)
        if bits:
            net.params = [trunc(p, bits) if p.ndim == 4 else p for p in net.params]
            x = trunc(x, bits)
        y_hat, caches = net.forward(x, True)
        net.backward(O.scce_backward(y_hat, Y.astype(np.float64)), caches)
        return [g.copy() for g in net.grads], O.scce_loss(y_hat, Y)
    g0, l0 = grads(0)
    g1, l1 = grads(10)
    assert abs(l1 - l0) < 1e-3 * l0
    errs = [rel_err(a, b) for a, b in zip(g1, g0)]
    conv_w = [errs[i] for i in (0, 4, 8, 12)]
    assert min(conv_w) > 2e-2, conv_w
    assert errs[-2] < 2e-2, errs[-2]          # the classifier, downstream of every decision, stays put
