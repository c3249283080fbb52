"""-m gpu: op-level parity of the CUDA path against (a) the golden fixtures made from the live
reference and (b) the CPU oracle on seeded inputs.  Everything goes through the Layer API, i.e.
through the ctypes C ABI.

Tolerances (north_star / SURVEY.md 8(c)), relative error = max|a-r| / max|r| per tensor:
  fp32 tier 1e-5 against float64; pooling values, argmax indices and tie masks bit-exact.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from gpu_helpers import drive, set_params
from oracle import shinnosuke_oracle as O

pytestmark = pytest.mark.gpu
TOL32 = 1e-5


def _conv_layer(meta, act):
    from shinnosuke_b200 import Conv2D
    N, C, H, W, F, k, stride, pad = meta
    if pad:
        assert 2 * pad == k - 1 and stride == 1
        return Conv2D(F, (k, k), stride=stride, padding='SAME', activation=act)
    return Conv2D(F, (k, k), stride=stride, padding='VALID', activation=act)


@pytest.mark.parametrize('name', ['conv_katA', 'conv_valid_s2', 'conv_cfg1_shape', 'conv_c64_same', 'conv_5x5_same'])
def test_conv2d_golden(name):
    g = load_golden(name)
    meta = [int(v) for v in g['meta']]
    layer = _conv_layer(meta, str(g['act']))
    from shinnosuke_b200 import Input, Activation
    # build first so the variables exist, then overwrite them with the fixture's
    inp = Input((None,) + g['X'].shape[1:]); src = Activation('linear')(inp); layer(src)
    set_params(layer, g['W'], g['b'])
    inp.input_tensor = g['X']
    for l in (inp, src, layer):
        l.forward(True)
    y = np.asarray(layer.output_tensor)
    assert rel_err(y, g['Y']) < TOL32
    from shinnosuke_b200.device import as_device
    layer._seed_grads(as_device(g['G']))
    layer.backward()
    W, b = layer.variables
    assert rel_err(np.asarray(W.grads), g['dW']) < TOL32
    assert rel_err(np.asarray(b.grads), g['db']) < TOL32
    assert rel_err(np.asarray(src.grads), g['dX']) < TOL32


@pytest.mark.parametrize('shape', [
    # N, C, H, W, F, k, stride, padding, act
    (3, 3, 32, 32, 64, 3, 1, 'SAME', 'relu'),       # cfg3 conv1 shape (small batch)
    (2, 64, 16, 16, 128, 3, 1, 'SAME', 'relu'),     # cfg3 conv2
    (4, 128, 8, 8, 256, 3, 1, 'SAME', 'relu'),      # cfg3 conv3
    (8, 256, 4, 4, 256, 3, 1, 'SAME', 'linear'),    # cfg3 conv4
    (5, 1, 28, 28, 8, 2, 1, 'VALID', 'relu'),       # cfg1
    (2, 7, 13, 9, 5, 3, 2, 'VALID', 'linear'),      # ragged everything, stride 2
    (1, 2, 5, 5, 3, 5, 1, 'VALID', 'relu'),         # filter == image
])
def test_conv2d_vs_oracle(shape):
    from shinnosuke_b200 import Conv2D
    N, C, H, W, F, k, stride, padding, act = shape
    r = np.random.default_rng(sum((i + 1) * v for i, v in enumerate(shape[:7])))      # stable across processes
    x = r.standard_normal((N, C, H, W)).astype(np.float32).astype(np.float64)
    w = (0.1 * r.standard_normal((F, C, k, k))).astype(np.float32).astype(np.float64)
    b = r.standard_normal((1, F)).astype(np.float32).astype(np.float64)
    (_, _, oh, ow), pad = O.conv_output_shape((N, C, H, W), F, (k, k), stride, padding)
    gup = r.standard_normal((N, F, oh, ow)).astype(np.float32).astype(np.float64)
    y_ref, cache = O.conv2d_forward(x, w, b, stride, pad, act)
    layer = Conv2D(F, (k, k), stride=stride, padding=padding, activation=act)
    from shinnosuke_b200 import Input, Activation
    from shinnosuke_b200.device import as_device
    inp = Input((None, C, H, W)); src = Activation('linear')(inp); layer(src)
    set_params(layer, w, b)
    inp.input_tensor = x
    for l in (inp, src, layer):
        l.forward(True)
    y = np.asarray(layer.output_tensor)
    assert rel_err(y, y_ref) < TOL32
    if act == 'relu':
        # a pre-activation within fp32 rounding (~1e-7) of zero may sit on the other side of the ReLU
        # kink than in float64; one such element moves dW by a whole |g*x| term.  The backward kernels
        # are therefore checked against the oracle run with the mask the device forward produced, and
        # the number of such elements is bounded (expected: 0, rarely 1).
        blocked = np.signbit(y)
        assert np.count_nonzero(blocked != (cache['z'] < 0)) <= 2
        cache['z'] = np.where(blocked, -1.0, 1.0)
    dx_ref, dw_ref, db_ref = O.conv2d_backward(gup, cache)
    layer._seed_grads(as_device(gup)); layer.backward()
    assert rel_err(np.asarray(layer.variables[0].grads), dw_ref) < TOL32
    assert rel_err(np.asarray(layer.variables[1].grads).reshape(-1), db_ref) < TOL32
    assert rel_err(np.asarray(src.grads), dx_ref) < TOL32
    # wgrad accumulates (+=) like the reference: a second backward doubles dW
    layer._seed_grads(as_device(gup)); layer._grads_premasked = False; layer.backward()
    assert rel_err(np.asarray(layer.variables[0].grads), 2 * dw_ref) < TOL32


def test_relu_zero_preactivation_passes_gradient():
    """utils/Activator.py:25 uses a strict '<': where the pre-activation is exactly 0 the
    gradient passes.  The device keeps that through the -0.0f convention."""
    from shinnosuke_b200 import Conv2D, Input, Activation
    from shinnosuke_b200.device import as_device
    x = np.zeros((1, 1, 4, 4)); x[0, 0, 1, 1] = -1.0; x[0, 0, 2, 2] = 2.0
    layer = Conv2D(1, (1, 1), padding='VALID', activation='relu')
    inp = Input((None, 1, 4, 4)); src = Activation('linear')(inp); layer(src)
    set_params(layer, np.ones((1, 1, 1, 1)), np.zeros((1, 1)))
    inp.input_tensor = x
    for l in (inp, src, layer):
        l.forward(True)
    y = np.asarray(layer.output_tensor)
    assert y.min() == 0.0 and y[0, 0, 2, 2] == 2.0
    layer._seed_grads(as_device(np.ones((1, 1, 4, 4)))); layer.backward()
    dx = np.asarray(src.grads)
    expect = np.ones((1, 1, 4, 4)); expect[0, 0, 1, 1] = 0.0       # only the strictly negative one is blocked
    assert np.array_equal(dx, expect)


@pytest.mark.parametrize('name', ['pool_katB_reshape', 'pool_katB_im2col', 'pool_odd27_im2col', 'pool_32_reshape',
                                  'pool_3x3_ties'])
def test_maxpool_golden_bit_exact(name):
    from shinnosuke_b200 import MaxPooling2D, Input, Activation
    from shinnosuke_b200.device import as_device
    g = load_golden(name)
    p = int(g['pool'])
    layer = MaxPooling2D((p, p))
    inp = Input((None,) + g['X'].shape[1:]); src = Activation('linear')(inp); layer(src)
    layer.mode = str(g['mode'])
    inp.input_tensor = g['X']
    for l in (inp, src, layer):
        l.forward(True)
    assert np.array_equal(np.asarray(layer.output_tensor), g['Y'])                 # values bit-exact
    if 'argmax' in g:
        am = layer.argmax_index
        assert am.dtype == np.int64 and np.array_equal(am, g['argmax'])            # indices bit-exact, (n,i,j,c) order
    G32 = g['G'].astype(np.float32).astype(np.float64)
    layer._seed_grads(as_device(G32)); layer.backward()
    dx = np.asarray(src.grads)
    _, cache = O.maxpool_forward(g['X'], p, None, str(g['mode']))
    dx_ref = O.maxpool_backward(G32, cache)
    assert np.array_equal(dx != 0, dx_ref != 0)                                    # tie mask / routing bit-exact
    assert rel_err(dx, dx_ref) < 1e-7
    assert rel_err(dx, g['dX']) < 1e-6


@pytest.mark.parametrize('cfg', [(2, 64, 32, 32, 2, 2), (3, 5, 11, 13, 3, 2), (1, 8, 27, 27, 2, 2), (2, 4, 9, 9, 3, 3),
                                 (1, 3, 8, 8, 2, 1)])
@pytest.mark.parametrize('mode', ['reshape', 'im2col'])
def test_maxpool_vs_oracle(cfg, mode):
    from shinnosuke_b200 import MaxPooling2D, Input, Activation
    from shinnosuke_b200.device import as_device
    N, C, H, W, p, s = cfg
    r = np.random.default_rng(N * 1000 + C)
    x = r.integers(-3, 4, (N, C, H, W)).astype(np.float64)         # many ties on purpose
    layer = MaxPooling2D((p, p), stride=s)
    inp = Input((None, C, H, W)); src = Activation('linear')(inp); layer(src)
    layer.mode = mode
    inp.input_tensor = x
    for l in (inp, src, layer):
        l.forward(True)
    y_ref, cache = O.maxpool_forward(x, p, s, mode)
    assert np.array_equal(np.asarray(layer.output_tensor), y_ref)
    if mode == 'im2col':
        assert np.array_equal(layer.argmax_index, cache['argmax'])
    gup = r.standard_normal(y_ref.shape).astype(np.float32).astype(np.float64)
    layer._seed_grads(as_device(gup)); layer.backward()
    assert rel_err(np.asarray(src.grads), O.maxpool_backward(gup, cache)) < 1e-6


@pytest.mark.parametrize('name', ['dense_relu', 'dense_softmax', 'dense_linear'])
def test_dense_golden(name):
    from shinnosuke_b200 import Dense, Input, Activation
    from shinnosuke_b200.device import as_device
    g = load_golden(name)
    layer = Dense(g['W'].shape[1], activation=str(g['act']))
    inp = Input((None, g['X'].shape[1])); src = Activation('linear')(inp); layer(src)
    set_params(layer, g['W'], g['b'])
    inp.input_tensor = g['X']
    for l in (inp, src, layer):
        l.forward(True)
    assert rel_err(np.asarray(layer.output_tensor), g['Y']) < TOL32
    layer._seed_grads(as_device(g['G'])); layer.backward()
    assert rel_err(np.asarray(layer.variables[0].grads), g['dW']) < TOL32
    assert rel_err(np.asarray(layer.variables[1].grads), g['db']) < TOL32
    assert rel_err(np.asarray(src.grads), g['dX']) < TOL32


@pytest.mark.parametrize('mkn', [(128, 784, 500), (256, 1352, 10), (64, 1024, 10), (1, 3, 2), (130, 70, 66)])
def test_dense_vs_oracle(mkn):
    from shinnosuke_b200 import Dense, Input, Activation
    from shinnosuke_b200.device import as_device
    M, K, Nout = mkn
    r = np.random.default_rng(M + K)
    x = r.standard_normal((M, K)).astype(np.float32).astype(np.float64)
    w = (0.1 * r.standard_normal((K, Nout))).astype(np.float32).astype(np.float64)
    b = r.standard_normal((1, Nout)).astype(np.float32).astype(np.float64)
    gup = r.standard_normal((M, Nout)).astype(np.float32).astype(np.float64)
    y_ref, cache = O.dense_forward(x, w, b, 'relu')
    dx_ref, dw_ref, db_ref = O.dense_backward(gup, cache)
    layer = Dense(Nout, activation='relu')
    inp = Input((None, K)); src = Activation('linear')(inp); layer(src)
    set_params(layer, w, b)
    inp.input_tensor = x
    for l in (inp, src, layer):
        l.forward(True)
    assert rel_err(np.asarray(layer.output_tensor), y_ref) < TOL32
    layer._seed_grads(as_device(gup)); layer.backward()
    assert rel_err(np.asarray(layer.variables[0].grads), dw_ref) < TOL32
    assert rel_err(np.asarray(layer.variables[1].grads), db_ref) < TOL32
    assert rel_err(np.asarray(src.grads), dx_ref) < TOL32


@pytest.mark.parametrize('name', ['bn_4d', 'bn_2d'])
def test_batchnorm_golden(name):
    from shinnosuke_b200 import BatchNormalization, Input, Activation
    from shinnosuke_b200.device import as_device
    g = load_golden(name)
    layer = BatchNormalization()
    inp = Input((None,) + g['X'].shape[1:]); src = Activation('linear')(inp); layer(src)
    set_params(layer, g['gamma'], g['beta'])
    inp.input_tensor = g['X']
    for l in (inp, src, layer):
        l.forward(True)
    assert rel_err(np.asarray(layer.output_tensor), g['Y']) < TOL32
    assert rel_err(layer.moving_mean, g['moving_mean']) < TOL32 and rel_err(layer.moving_variance, g['moving_var']) < TOL32
    layer._seed_grads(as_device(g['G'])); layer.backward()
    assert rel_err(np.asarray(src.grads), g['dX']) < TOL32
    assert rel_err(np.asarray(layer.variables[0].grads), g['dgamma']) < TOL32
    assert rel_err(np.asarray(layer.variables[1].grads), g['dbeta']) < TOL32
    inp.input_tensor = g['X2']
    for l in (inp, src, layer):
        l.forward(False)
    assert rel_err(np.asarray(layer.output_tensor), g['Y2']) < TOL32


def test_batchnorm_large_offset_and_relu_mask_fusion():
    """|mean| >> std (float64 accumulation keeps the 1e-5 tier) and the fused ReLU mask of the
    producing layer: Conv(relu) -> BN must equal the oracle's unfused sequence."""
    from shinnosuke_b200 import BatchNormalization, Conv2D, Input, Activation
    from shinnosuke_b200.device import as_device
    r = np.random.default_rng(42)
    x = (50.0 + r.standard_normal((8, 16, 6, 6))).astype(np.float32).astype(np.float64)
    layer = BatchNormalization()
    inp = Input((None, 16, 6, 6)); src = Activation('linear')(inp); layer(src)
    inp.input_tensor = x
    for l in (inp, src, layer):
        l.forward(True)
    y_ref, cache, _, _ = O.batchnorm_forward(x, np.ones(16), np.zeros(16), np.zeros(16), np.ones(16))
    assert rel_err(np.asarray(layer.output_tensor), y_ref) < 2e-5     # x itself carries 50 * 2^-24 of fp32 rounding
    # conv(relu) -> bn
    x = r.standard_normal((4, 3, 8, 8)).astype(np.float32).astype(np.float64)
    w = (0.3 * r.standard_normal((8, 3, 3, 3))).astype(np.float32).astype(np.float64)
    b = (0.1 * r.standard_normal((1, 8))).astype(np.float32).astype(np.float64)
    gup = r.standard_normal((4, 8, 8, 8)).astype(np.float32).astype(np.float64)
    conv = Conv2D(8, (3, 3), padding='SAME', activation='relu')
    bn = BatchNormalization()
    inp = Input((None, 3, 8, 8)); src = Activation('linear')(inp); conv(src); bn(conv)
    set_params(conv, w, b)
    inp.input_tensor = x
    for l in (inp, src, conv, bn):
        l.forward(True)
    bn._seed_grads(as_device(gup)); bn.backward()
    assert conv._grads_premasked
    conv.backward()
    yc, cc = O.conv2d_forward(x, w, b, 1, (1, 1), 'relu')
    yb, cb, _, _ = O.batchnorm_forward(yc, np.ones(8), np.zeros(8), np.zeros(8), np.ones(8))
    dxb, dg, dbt = O.batchnorm_backward(gup, cb)
    dxc, dw, db = O.conv2d_backward(dxb, cc)
    assert rel_err(np.asarray(bn.output_tensor), yb) < TOL32
    assert rel_err(np.asarray(conv.variables[0].grads), dw) < TOL32
    assert rel_err(np.asarray(src.grads), dxc) < TOL32


def test_loss_golden():
    from shinnosuke_b200.utils.Objectives import SparseCategoricalCrossEntropy
    from shinnosuke_b200.device import as_device
    from shinnosuke_b200 import Activation, Input
    g = load_golden('loss_scce')
    inp = Input((None, 10)); sm = Activation('softmax')(inp)
    inp.input_tensor = g['Z']
    inp.forward(False); sm.forward(False)
    p = np.asarray(sm.output_tensor)
    assert rel_err(p, g['P']) < TOL32
    L = SparseCategoricalCrossEntropy()
    dz = np.asarray(L.backward(as_device(g['P']), as_device(g['Y'])))
    assert rel_err(dz, g['dZ']) < TOL32
    loss, acc = L.calc_loss_acc(as_device(g['P']), as_device(g['Y']))
    assert abs(loss - float(g['loss'])) < 1e-5 * abs(float(g['loss']))
    assert acc == float(g['acc'])


def test_optimizers_golden():
    from shinnosuke_b200 import Variable
    from shinnosuke_b200.utils.Optimizers import StochasticGradientDescent, Momentum
    g = load_golden('optimizers')
    for name, opt in (('sgd', StochasticGradientDescent(lr=0.05)), ('momentum', Momentum(lr=0.05, rho=0.9))):
        vs = [Variable(g['w0']), Variable(g['w1'])]
        for v in vs:
            v.ensure_device()
        for t in range(3):
            for i, v in enumerate(vs):
                from shinnosuke_b200.device import as_device
                v.grads.add_(as_device(g['g%d_t%d' % (i, t)]))      # layers accumulate with +=
            opt.update(vs)
            for i, v in enumerate(vs):
                assert rel_err(np.asarray(v.output_tensor), g['%s_w%d_t%d' % (name, i, t)]) < TOL32
                gref = g['%s_g%d_t%d' % (name, i, t)]
                got = np.asarray(v.grads)
                assert (not gref.any() and not got.any()) or rel_err(got, gref) < TOL32


def test_flatten_and_padding_roundtrip():
    from shinnosuke_b200 import Flatten, ZeroPadding2D, Input, Activation
    from shinnosuke_b200.device import as_device
    r = np.random.default_rng(8)
    x = r.standard_normal((3, 5, 4, 6)).astype(np.float32).astype(np.float64)
    fl = Flatten()
    y, dx = drive(fl, x, x.reshape(3, -1) * 2)
    assert np.array_equal(y, x.reshape(3, -1))                     # NCHW (c,i,j) order, like the reference
    assert np.array_equal(dx, 2 * x)
    zp = ZeroPadding2D((1, 2))
    gup = r.standard_normal((3, 5, 6, 10)).astype(np.float32).astype(np.float64)
    y, dx = drive(zp, x, gup)
    assert np.array_equal(y, np.pad(x, ((0, 0), (0, 0), (1, 1), (2, 2))))
    assert np.array_equal(dx, gup[:, :, 1:-1, 2:-2])


def test_empty_batch_is_a_noop():
    from shinnosuke_b200._lib import lib
    lib.sn_conv2d_fprop(8, 8, None, 8, 0, 3, 8, 8, 4, 3, 3, 1, 1, 1, 0, 0, None)
    lib.sn_maxpool2d_fwd(8, 8, None, 0, 8, 8, 4, 2, 2, 0, None)
    lib.sn_sgd_update(8, 8, 0, 0.1, 1.0, None)


@pytest.mark.parametrize('mode', ['reshape', 'im2col'])
@pytest.mark.parametrize('shape', [(4, 64, 8, 8), (2, 16, 6, 10), (3, 256, 4, 4)])
def test_fused_batchnorm_maxpool_pair(shape, mode):
    """csrc/bn_pool.cu: BatchNormalization -> MaxPooling2D((2,2)) as ONE pass must equal the two
    oracle layers back to back (forward values, pool codes, dX with the producer's ReLU mask,
    dgamma/dbeta, moving statistics)."""
    from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Input, Activation
    from shinnosuke_b200.device import as_device
    N, C, H, W = shape
    r = np.random.default_rng(C + H)
    x = r.standard_normal((N, 3, H, W)).astype(np.float32).astype(np.float64)
    w = (0.3 * r.standard_normal((C, 3, 3, 3))).astype(np.float32).astype(np.float64)
    b = (0.1 * r.standard_normal((1, C))).astype(np.float32).astype(np.float64)
    gam = (1 + 0.3 * r.standard_normal(C)).astype(np.float32).astype(np.float64)
    gam[::5] *= -1                                            # negative gamma: max must be taken AFTER normalisation
    bet = (0.2 * r.standard_normal(C)).astype(np.float32).astype(np.float64)
    gup = r.standard_normal((N, C, H // 2, W // 2)).astype(np.float32).astype(np.float64)
    conv = Conv2D(C, (3, 3), padding='SAME', activation='relu')
    bn = BatchNormalization()
    pool = MaxPooling2D((2, 2))
    inp = Input((None, 3, H, W)); src = Activation('linear')(inp); conv(src); bn(conv); pool(bn)
    pool.mode = mode
    bn._fuse_with_pool = pool
    set_params(conv, w, b); set_params(bn, gam, bet)
    inp.input_tensor = x
    for l in (inp, src, conv, bn, pool):
        l.forward(True)
    assert bn._fused_active and bn.output_tensor is None
    yc, cc = O.conv2d_forward(x, w, b, 1, (1, 1), 'relu')
    yb, cb, mm, mv = O.batchnorm_forward(yc, gam, bet, np.zeros(C), np.ones(C))
    yp, cp = O.maxpool_forward(yb, 2, None, mode)
    assert rel_err(np.asarray(pool.output_tensor), yp) < TOL32
    assert rel_err(bn.moving_mean, mm) < TOL32 and rel_err(bn.moving_variance, mv) < TOL32
    pool._seed_grads(as_device(gup))
    pool.backward(); bn.backward()
    assert conv._grads_premasked
    conv.backward()
    dyb = O.maxpool_backward(gup, cp)
    dxb, dg, dbt = O.batchnorm_backward(dyb, cb)
    dxc, dw, db = O.conv2d_backward(dxb, cc)
    assert rel_err(np.asarray(bn.variables[0].grads), dg) < TOL32
    assert rel_err(np.asarray(bn.variables[1].grads), dbt) < TOL32
    assert rel_err(np.asarray(conv.variables[0].grads), dw) < TOL32
    assert rel_err(np.asarray(src.grads), dxc) < TOL32


# ------------------------------------------------------------------ RMSprop / AdaGrad / AdaDelta / Adam
@pytest.mark.parametrize('name', ['rmsprop', 'adagrad', 'adadelta', 'adam'])
def test_adaptive_optimizers_match_reference_golden(name):
    """utils/Optimizers.py:61-157 on the device against the values the live reference produced
    (tests/golden/optimizers_adaptive.npz): four updates on a never-zeroed gradient stream."""
    from shinnosuke_b200.layers.Base import Variable
    from shinnosuke_b200.utils.Optimizers import RMSprop, AdaGrad, AdaDelta, Adam
    g = load_golden('optimizers_adaptive')
    opt = dict(rmsprop=RMSprop(lr=0.01, rho=0.9), adagrad=AdaGrad(lr=0.05), adadelta=AdaDelta(rho=0.95),
               adam=Adam(lr=0.01, beta1=0.9, beta2=0.999))[name]
    vs = [Variable(g['w0'].copy()), Variable(g['w1'].copy())]
    for v in vs:
        v.ensure_device()
    acc = [np.zeros_like(g['w0']), np.zeros_like(g['w1'])]
    for t in range(4):
        for i, v in enumerate(vs):
            acc[i] = acc[i] + g['g%d_t%d' % (i, t)]           # layers accumulate with +=; never reset
            v.grads = acc[i]
        opt.update(vs)
        for i, v in enumerate(vs):
            assert rel_err(np.asarray(v.output_tensor), g['%s_w%d_t%d' % (name, i, t)]) < 1e-5, (name, t, i)
    assert opt.iterations == int(g['%s_iterations' % name])


@pytest.mark.parametrize('name', ['adam', 'rmsprop'])
def test_adaptive_optimizer_training_matches_oracle_under_graph_replay(name):
    """A whole model trained with Adam / RMSprop: eager first steps, then captured-graph replays
    (Adam's bias correction must advance on the device), against the float64 oracle."""
    from shinnosuke_b200 import Sequential, Dense
    r = np.random.default_rng(3)
    X = r.random((96, 20)).astype(np.float32).astype(np.float64)
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 86)]))
    np.random.seed(0)
    m = Sequential()
    m.add(Dense(32, n_in=20, activation='relu'))
    m.add(Dense(10, activation='softmax'))
    m.compile(name, 'sparse_categorical_crossentropy')
    m.fit(X, Y, batch_size=16, epochs=2, shuffle=False, validation_ratio=0., verbose=0)        # 12 steps, 8+ of them replays
    np.random.seed(0)
    w1 = 0.1 * np.random.randn(20, 32); w2 = 0.1 * np.random.randn(32, 10)      # initializer 'Normal' (std 0.1), layer order
    params = [w1, np.zeros((1, 32)), w2, np.zeros((1, 10))]
    grads = [np.zeros_like(p) for p in params]
    s1 = [np.zeros_like(p) for p in params]; s2 = [np.zeros_like(p) for p in params]
    it = 0
    losses = []
    for ep in range(2):
        for a in range(0, 96, 16):
            xs, ys = X[a:a + 16], Y[a:a + 16]
            z1 = xs @ params[0] + params[1]; h = np.maximum(z1, 0)
            p = O.softmax_forward(h @ params[2] + params[3])
            losses.append(O.scce_loss(p, ys))
            dz2 = O.scce_backward(p, ys)
            grads[2] += h.T @ dz2; grads[3] += dz2.sum(0, keepdims=True)
            dh = dz2 @ params[2].T; dh[z1 < 0] = 0
            grads[0] += xs.T @ dh; grads[1] += dh.sum(0, keepdims=True)
            if name == 'adam':
                it += 1
                O.adam_update(params, grads, s1, s2, it)
                it += 1
            else:
                O.rmsprop_update(params, grads, s1)
    assert np.allclose(m.train_loss, losses, rtol=2e-5)
    for v, pr in zip(m.trainable_variables, params):
        assert rel_err(np.asarray(v.output_tensor), pr) < 2e-5
