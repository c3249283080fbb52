"""-m gpu: parity AT THE BENCH'S GRID SIZES.

The tensor-core kernels are persistent (grid = min(tiles, 148)): with fewer tiles than SMs every
CTA handles exactly one tile and the TMEM double buffer (`acc = iter & 1`, phase `(iter >> 1) & 1`),
the carry-over of the shared-memory ring from tile to tile and the stem kernel's second epilogue
warp set never run.  Everything here has at least three tiles per CTA — the layer shapes of
BASELINE config 3 at batch 512 (the configuration bench.py times) among them — and is compared
with the float64 oracle on the WHOLE batch: a convolution is independent per image, so the oracle
runs over chunks of images (its float64 column matrix for conv2 at batch 512 would be 604 MB in one
piece) and dW / db are summed over the chunks.

Tolerances: tf32 2e-3, bf16 1e-2 (max|a - r| / max|r| per tensor); pooling
codes bit-exact; the BatchNorm + pool pair on the tensor tiers 1e-5 against the oracle fed with the
device's own convolution output (that isolates the fused pass from the reduced-precision products).
"""
import numpy as np
import pytest

from conftest import rel_err
from gpu_helpers import set_params
from oracle import shinnosuke_oracle as O

pytestmark = pytest.mark.gpu
TOL = {'tf32': 2e-3, 'bf16': 1e-2}


def _conv_full(shape, precision, act='relu', chunk=32, pair_mode=None, need=('y', 'dw', 'db', 'dx')):
    """Drive ONE Conv2D layer on the device at `shape` and return {name: (device, oracle)}; the oracle
    runs over chunks of `chunk` images.  With ReLU the backward reference uses the device's own mask
    (see test_gpu_tf32._run_conv) and the flipped fraction is bounded."""
    from shinnosuke_b200 import Conv2D, Input, Activation
    from shinnosuke_b200.device import as_device
    from shinnosuke_b200._lib import lib
    N, C, H, W, F, k = shape
    pad = (k - 1) // 2
    r = np.random.default_rng(N + 7 * C + 13 * H + 17 * F)
    x = r.standard_normal((N, C, H, W), dtype=np.float32).astype(np.float64)
    w = (0.1 * r.standard_normal((F, C, k, k))).astype(np.float32).astype(np.float64)
    b = r.standard_normal((1, F)).astype(np.float32).astype(np.float64)
    gup = r.standard_normal((N, F, H, W), dtype=np.float32).astype(np.float64)
    if pair_mode is not None:
        lib.sn_igemm_set_pair_mode(pair_mode)
    try:
        layer = Conv2D(F, (k, k), padding='SAME', activation=act)
        inp = Input((None, C, H, W)); src = Activation('linear')(inp); layer(src)
        layer.precision = precision
        set_params(layer, w, b)
        inp.input_tensor = x
        pairs0 = lib.sn_igemm_pair_launches()
        for l in (inp, src, layer):
            l.forward(True)
        y = np.asarray(layer.output_tensor)
        layer._seed_grads(as_device(gup)); layer.backward()
        pairs = lib.sn_igemm_pair_launches() - pairs0
        dw = np.asarray(layer.variables[0].grads)
        db = np.asarray(layer.variables[1].grads).reshape(-1)
        dx = np.asarray(src.grads)
    finally:
        if pair_mode is not None:
            lib.sn_igemm_set_pair_mode(-1)
    y_ref = np.empty_like(y); dx_ref = np.empty_like(dx)
    dw_ref = np.zeros_like(dw); db_ref = np.zeros_like(db)
    flips = 0
    for a in range(0, N, chunk):
        sl = slice(a, min(N, a + chunk))
        yc, cache = O.conv2d_forward(x[sl], w, b, 1, (pad, pad), act)
        y_ref[sl] = yc
        if act == 'relu':
            blocked = np.signbit(y[sl])
            flips += np.count_nonzero(blocked != (cache['z'] < 0))
            cache['z'] = np.where(blocked, -1.0, 1.0)
        dxc, dwc, dbc = O.conv2d_backward(gup[sl], cache)
        dx_ref[sl] = dxc; dw_ref += dwc; db_ref += np.asarray(dbc).reshape(-1)
    assert flips <= 2e-3 * y.size, flips
    out = dict(y=(y, y_ref), dw=(dw, dw_ref), db=(db, db_ref), dx=(dx, dx_ref))
    return {k_: v for k_, v in out.items() if k_ in need}, pairs


# BASELINE config 3 at batch 512 = what bench.py runs: (N, C, H, W, F, k) and the tile counts of the
# three contraction kernels of each layer (128-pixel tiles; fprop / dgrad are persistent over them)
BENCH_LAYERS = [
    (512, 3, 32, 32, 64, 3),       # stem: 4096 pixel tiles on 148 CTAs, both epilogue warp sets
    (512, 64, 16, 16, 128, 3),     # conv2: 1024 tiles (6.9 per CTA)
    (512, 128, 8, 8, 256, 3),      # conv3: 256 tiles x 256 filters
    (512, 256, 4, 4, 256, 3),      # conv4: 64 pixel tiles x 2 filter tiles
]


@pytest.mark.parametrize('precision', ['bf16', 'tf32'])
@pytest.mark.parametrize('shape', BENCH_LAYERS)
def test_bench_layer_shapes_at_batch_512(shape, precision):
    out, _ = _conv_full(shape, precision)
    for name, (got, ref) in out.items():
        assert np.isfinite(got).all(), name
        assert rel_err(got, ref) < TOL[precision], (name, rel_err(got, ref))


@pytest.mark.parametrize('precision', ['bf16', 'tf32'])
@pytest.mark.parametrize('shape', [
    (160, 32, 16, 16, 64, 3),      # 320 tiles of BN = 64 (two k-slabs per stage), 2.2 per CTA
    (300, 64, 8, 8, 96, 3),        # 150 tiles: two CTAs get a second tile, the others stop after one
    (37, 64, 16, 16, 128, 3),      # 74 tiles (one per CTA) — the single-tile path next to the ones above
])
def test_persistent_tiles_more_than_sms(shape, precision):
    out, _ = _conv_full(shape, precision, act='linear')
    for name, (got, ref) in out.items():
        assert rel_err(got, ref) < TOL[precision], (name, rel_err(got, ref))


@pytest.mark.parametrize('precision', ['bf16', 'tf32'])
@pytest.mark.parametrize('shape', [(512, 64, 16, 16, 128, 3), (512, 128, 8, 8, 256, 3)])
def test_cta_pairs_at_batch_512(shape, precision):
    """cta_group::2 tile pairs forced on at the bench shapes: 512 / 128 tile pairs on 74 clusters, so
    every cluster walks several pairs (accumulator double buffer across the pair, multicast commits)."""
    out, pairs = _conv_full(shape, precision, act='linear', pair_mode=2, need=('y', 'dx'))
    assert pairs >= 1, 'the pair variant was not launched'
    for name, (got, ref) in out.items():
        assert rel_err(got, ref) < TOL[precision], (name, rel_err(got, ref))


def test_dyadic_data_is_exact_at_batch_512():
    """Small dyadic values are exact in bf16 and every partial sum is exact in fp32, so the bf16
    kernels must reproduce the oracle to fp32 rounding at the full grid size: any tile / phase /
    ring mix-up between consecutive tiles of a CTA shows up as an O(1) error here."""
    from shinnosuke_b200 import Conv2D, Input, Activation
    from shinnosuke_b200.device import as_device
    r = np.random.default_rng(5)
    for (N, C, H, F) in [(512, 64, 16, 128), (512, 128, 8, 256)]:
        x = r.integers(-4, 5, (N, C, H, H)).astype(np.float64) / 4.0
        w = r.integers(-4, 5, (F, C, 3, 3)).astype(np.float64) / 8.0
        b = r.integers(-4, 5, (1, F)).astype(np.float64)
        gup = r.integers(-2, 3, (N, F, H, H)).astype(np.float64) / 2.0
        layer = Conv2D(F, (3, 3), padding='SAME', activation='linear')
        inp = Input((None, C, H, H)); src = Activation('linear')(inp); layer(src)
        layer.precision = 'bf16'
        set_params(layer, w, b)
        inp.input_tensor = x
        for l in (inp, src, layer):
            l.forward(True)
        y = np.asarray(layer.output_tensor)
        layer._seed_grads(as_device(gup)); layer.backward()
        dx = np.asarray(src.grads)
        dw = np.asarray(layer.variables[0].grads)
        dw_ref = np.zeros_like(dw)
        for a in range(0, N, 64):
            yc, cache = O.conv2d_forward(x[a:a + 64], w, b, 1, (1, 1), 'linear')
            dxc, dwc, _ = O.conv2d_backward(gup[a:a + 64], cache)
            assert np.array_equal(y[a:a + 64], yc), (N, C, a)             # exact
            assert np.array_equal(dx[a:a + 64], dxc), (N, C, a)
            dw_ref += dwc
        assert rel_err(dw, dw_ref) < 1e-6        # 131072-term sums accumulate across CTAs with fp32 atomics


BLOCKS = [((512, 3, 32, 32, 64), 'tf32'), ((512, 3, 32, 32, 64), 'bf16'),
          ((512, 64, 16, 16, 128), 'bf16'), ((512, 128, 8, 8, 256), 'bf16'), ((512, 256, 4, 4, 256), 'bf16')]


@pytest.mark.parametrize('shape,precision', BLOCKS)
def test_conv_bn_pool_block_at_batch_512(shape, precision):
    """Conv2D -> BatchNormalization -> MaxPooling2D, the block of config 3, at every layer shape of the
    bench: the convolution's epilogue statistics over up to 4096 tiles, the fused BN + pool forward and
    backward passes over maps of up to 134 MB.  In the bf16 tier the convolution stores its output map in
    bf16 (csrc/conv_tc.cu OBF16, conv_stem_tc.cu) and the fused passes read it as such.  The oracle's
    BatchNorm / pool run on the device's own convolution output, so the bound on everything behind the
    convolution is the fp32 tier's 1e-5 in every tier; pool values bit-exact given the same map."""
    from shinnosuke_b200 import Conv2D, Input, Activation, BatchNormalization, MaxPooling2D
    from shinnosuke_b200.device import as_device
    N, C, H, W, F = shape
    r = np.random.default_rng(99 + C)
    x = r.standard_normal((N, C, H, W), dtype=np.float32).astype(np.float64)
    w = ((0.3 if C == 3 else 0.05) * r.standard_normal((F, C, 3, 3))).astype(np.float32).astype(np.float64)
    b = (0.5 * r.standard_normal((1, F))).astype(np.float32).astype(np.float64)
    gam = (1 + 0.3 * r.standard_normal(F)).astype(np.float32).astype(np.float64)
    gam[::7] *= -1
    bet = (0.2 * r.standard_normal(F)).astype(np.float32).astype(np.float64)
    gup = r.standard_normal((N, F, H // 2, W // 2), dtype=np.float32).astype(np.float64)
    inp = Input((None, C, H, W))                              # like the stem in the model: nothing upstream wants dX
    conv = Conv2D(F, (3, 3), padding='SAME', activation='relu')(inp)
    bn = BatchNormalization()(conv)
    pool = MaxPooling2D((2, 2))(bn)
    bn._fuse_with_pool = pool
    conv._stats_to = bn
    layers = [inp, conv, bn, pool]
    for l in layers:
        l.precision = precision
    set_params(conv, w, b); set_params(bn, gam, bet)
    inp.input_tensor = x
    for l in layers:
        l.forward(True)
    assert bn._fused_active
    assert conv.output_tensor.dtype == ('bf16' if precision == 'bf16' else 'f32')
    yc = np.asarray(conv.output_tensor)                       # device conv output (what BatchNormalization saw)
    # the convolution itself against the oracle on a 32-image subset (the full batch is test_bench_layer_shapes)
    sub = slice(240, 272)
    yc_ref, _ = O.conv2d_forward(x[sub], w, b, 1, (1, 1), 'relu')
    assert rel_err(yc[sub], yc_ref) < TOL[precision]
    blocked = np.signbit(yc)                                  # -0.0 = blocked by the ReLU (utils/Activator.py:25)
    yb, cb, mm, mv = O.batchnorm_forward(yc, gam, bet, np.zeros(F), np.ones(F))
    yp, cp = O.maxpool_forward(yb, 2, None, 'reshape')
    assert rel_err(np.asarray(pool.output_tensor), yp) < 1e-5
    assert rel_err(bn.moving_mean, mm) < 1e-5 and rel_err(bn.moving_variance, mv) < 1e-5
    pool._seed_grads(as_device(gup))
    pool.backward(); bn.backward()
    dyb = O.maxpool_backward(gup, cp)
    dxb, dg, dbt = O.batchnorm_backward(dyb, cb)
    dxb[blocked] = 0.0                                         # the producing convolution's ReLU mask, applied by the fused pass
    assert rel_err(np.asarray(bn.variables[0].grads), dg) < 1e-5
    assert rel_err(np.asarray(bn.variables[1].grads), dbt) < 1e-5
    assert rel_err(np.asarray(conv.variables[1].grads).reshape(-1), dxb.sum(axis=(0, 2, 3))) < 2e-5     # bias gradient = column sums
    wg16, _ = conv._bf16_grad_plan()
    if wg16:
        # the fused pass wrote dX only as the bf16 copy the convolution's wgrad reads (no fp32 map at all)
        gb = conv._bf16_buf('dy_bf16', N * F * H * W).float().cpu().numpy().reshape(N, H, W, F).transpose(0, 3, 1, 2)
        assert rel_err(gb, dxb) < 4e-3                         # bf16 rounding of an fp32-exact result: 2^-9 of the largest element
    else:
        assert rel_err(np.asarray(conv.grads), dxb) < 1e-5
    conv.backward()
    dw_ref = np.zeros((F, C, 3, 3))
    for a in range(0, N, 64):
        _, cache = O.conv2d_forward(x[a:a + 64], w, b, 1, (1, 1), 'linear')
        _, dwc, _ = O.conv2d_backward(dxb[a:a + 64], cache, need_dx=False)
        dw_ref += dwc
    assert rel_err(np.asarray(conv.variables[0].grads), dw_ref) < TOL[precision]


_ORACLE_TRAJ = {}


def _oracle_trajectory(steps, dense_scale):
    """The float64 oracle on `steps` minibatches of 512 of bench.py's synthetic config-3 data (about
    10 s of host time per step; shared by the tiers).  dense_scale multiplies the classifier's initial
    weights (1.0 = the benchmarked initialisation)."""
    import bench
    key = (steps, dense_scale)
    if key not in _ORACLE_TRAJ:
        X, Y = bench.synth('cfg3', 512 * steps)
        np.random.seed(0)
        spec, ishape = O.spec_cfg3()
        net = O.OracleNet(spec, ishape)
        net.params[-2] *= dense_scale
        p0 = [p.copy() for p in net.params]
        net.fit(X.astype(np.float64), Y.astype(np.float64), 512, 1, shuffle=False)
        _ORACLE_TRAJ[key] = (X, Y, net, p0)
    return _ORACLE_TRAJ[key]


# (initial scale of the classifier weights, steps, bound on the losses after the first update)
#  * 1.0 = the model exactly as bench.py builds it: Dense W ~ N(0, 0.1) on 1024 normalised inputs gives
#    logits of std ~3 and a first loss of ~9; the loss drops by 2-3 per step, and a 0.5 % error in one
#    bf16 update moves the next loss by more than 1 % (measured 2.9 % at the third loss).  The first loss
#    (pure forward) and the loss after one update are held to the tier bound, later ones to 5e-2.
#  * 0.1 = the same network with a well-conditioned head (first loss ~2.4): EVERY loss of five steps is
#    held to the tier bound (north_star: "losses within 1e-2 for bf16").
TRAJ_CASES = [(1.0, 3, 5e-2), (0.1, 5, None)]


@pytest.mark.parametrize('precision', ['bf16', 'tf32'])
@pytest.mark.parametrize('dense_scale,steps,later_bound', TRAJ_CASES)
def test_trajectory_bench_config_at_batch_512(precision, dense_scale, steps, later_bound):
    """THE configuration bench.py times — config 3, batch 512, the same model builder and synthetic
    data, captured-graph replay included (step 1 runs eagerly, step 2 is captured, the rest replay) —
    against the float64 oracle trained on the same minibatches: the losses of the trajectory within
    the tier's bound (see TRAJ_CASES), accuracies equal up to a few near-ties, and the weights after
    the updates within 5e-3 (tf32) / 1e-2 (bf16) of the oracle's (well-conditioned case).

    The accumulated UPDATE W_t - W_0 is only held to a gross-error bound (0.5: a missing term, a wrong
    sign or a wrong scale would show).  It cannot be held tighter, and that is a property of this
    network, not of the kernels: the float64 ORACLE'S OWN first gradient moves by 6-36 % per tensor when
    nothing but the input and the conv filters are truncated to TF32's 10 mantissa bits, and by
    17-61 % for bf16's 7 (tests/test_oracle_golden.py::test_config3_gradient_is_ill_conditioned pins
    that measurement): every rounding flips a few of the ReLU masks and max-pool winners, each flip
    moves a whole term of sums that cancel to ~1/sqrt(n) of their terms, and every BatchNormalization
    backward subtracts the two dominant components of its incoming gradient.  The device tiers land
    inside that band (tf32 4-16 %, bf16 5-25 % after one update; the fp32 tier 1e-3..1e-2, which is
    the resolution of an fp32 master weight for an update of lr * g ~ 1e-5).  Kernel correctness at
    these sizes is what the op-level tests above establish, with the device's own masks."""
    import bench
    X, Y, net, p0 = _oracle_trajectory(steps, dense_scale)
    m = bench.build_model('cfg3', precision)
    dense = m.layers[-1]
    if dense_scale != 1.0:
        dense.variables[0].output_tensor = np.asarray(dense.variables[0].output_tensor) * dense_scale
    m.fit(X, Y, batch_size=512, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert len(m._graphs) >= 1, 'no step graph was captured'
    got, ref = np.asarray(m.train_loss), np.asarray(net.train_loss)
    err = np.abs(got - ref) / np.abs(ref)
    print('losses', precision, dense_scale, got, ref, err)
    tol = TOL[precision]
    assert (err[:2] < tol).all(), (got, ref, err)
    assert (err[2:] < (later_bound or tol)).all(), (got, ref, err)
    assert np.abs(np.asarray(m.train_acc) - np.asarray(net.train_acc)).max() <= 6.0 / 512
    convs = [l for l in m.layers if l.__class__.__name__ == 'Conv2D']
    tensors = [(conv.variables[0], 4 * i, 'conv%d W' % (i + 1)) for i, conv in enumerate(convs)]
    tensors.append((dense.variables[0], len(net.params) - 2, 'dense W'))
    for var, idx, name in tensors:       # oracle parameter order: conv W, b, BN gamma, beta per block, then Dense W, b
        w = np.asarray(var.output_tensor)
        e = rel_err(w, net.params[idx])
        eu = rel_err(w - p0[idx], net.params[idx] - p0[idx])
        print('%s after %d updates: rel err %.2e, of the update itself %.2e' % (name, steps, e, eu))
        # weights: within the tier's bound on the well-conditioned model; the as-benchmarked initialisation
        # takes updates of the size of the weights themselves (loss 9 -> 5.6 in two steps), so there the
        # weights inherit the update's conditioning and are only held to 0.1
        assert e < ({'tf32': 5e-3, 'bf16': 1e-2}[precision] if dense_scale != 1.0 else 0.1), (name, e)
        assert eu < 0.5, (name, eu)
