"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):   python tests/golden/make_golden.py

It imports the unmodified reference package (with a stub for ``matplotlib``,
which ``models.py:5`` imports unconditionally and which is not installed),
drives its own layers / models / optimizers on seeded inputs, and stores the
inputs and outputs as small ``.npz`` files.  ``tests/test_oracle_golden.py``
then pins ``oracle/shinnosuke_oracle.py`` to these outputs, and the ``-m gpu``
tests compare the CUDA path against the same fixtures.

Conventions: ``np.random.seed(0)`` immediately before building a
model (weight init consumes the legacy global stream); data from a fresh
``np.random.default_rng(1234)`` cast float32->float64 so every input is
fp32-representable.
"""
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, '/root')
_mpl, _plt = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
_mpl.pyplot = _plt
sys.modules.setdefault('matplotlib', _mpl)
sys.modules.setdefault('matplotlib.pyplot', _plt)
import reference as ref  # noqa: E402
from reference.layers.Base import Input, Variable  # noqa: E402
from reference.layers.Convolution import Conv2D, MaxPooling2D, ZeroPadding2D  # noqa: E402
from reference.layers.FC import Dense, Flatten  # noqa: E402
from reference.layers.Activation import Activation  # noqa: E402
from reference.layers.Normalization import BatchNormalization  # noqa: E402
from reference.models import Sequential, Model  # noqa: E402
from reference.utils.Objectives import SparseCategoricalCrossEntropy  # noqa: E402
from reference.utils.Optimizers import StochasticGradientDescent, Momentum  # noqa: E402
from reference.utils.MiniBatch import get_batches  # noqa: E402
from reference.utils.Preprocess import to_categorical  # noqa: E402
from reference.utils.ConvCol import im2col, col2im  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **arrs)
    print('wrote', name, {k: np.asarray(v).shape for k, v in arrs.items()})


def quiet(fn, *a, **k):
    """The reference prints a progress bar per minibatch."""
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ---------------------------------------------------------------- conv ----
def conv_case(name, N, C, H, W, F, k, stride, pad, act, first_layer=False, seed=0):
    """Reference Conv2D driven op-level.  pad>0 uses ZeroPadding2D + VALID (the
    reference's only correct SAME)."""
    np.random.seed(seed)
    r = np.random.default_rng(1234 + seed)
    inp = Input((None, C, H, W))
    src = Activation('linear')(inp)           # an inbound that requires grads
    cur = src
    zp = None
    if pad:
        zp = ZeroPadding2D((pad, pad))(cur)
        zp.require_grads = True               # so Conv2D takes the dgrad branch
        cur = zp
    conv = Conv2D(F, (k, k), stride=stride, padding='VALID', activation=act, initializer='normal')(cur)
    Wv, bv = conv.variables
    Wv.output_tensor = f32(Wv.output_tensor)
    bv.output_tensor = f32(bv.output_tensor)
    X = f32(r.standard_normal((N, C, H, W)))
    inp.input_tensor = X
    for l in [inp, src] + ([zp] if zp else []) + [conv]:
        l.forward(True)
    Y = conv.output_tensor.copy()
    G = f32(r.standard_normal(Y.shape))
    conv.grads = G.copy()
    conv.backward()
    dXp = cur.grads.copy()
    dX = dXp[:, :, pad:-pad, pad:-pad] if pad else dXp
    save(name, X=X, W=Wv.output_tensor, b=bv.output_tensor, Y=Y, G=G, dW=Wv.grads, db=bv.grads, dX=dX,
         meta=np.array([N, C, H, W, F, k, stride, pad]), act=np.array(act))


# ---------------------------------------------------------------- pool ----
def pool_case(name, x, pool, mode, g=None, seed=3):
    r = np.random.default_rng(seed)
    N, C, H, W = x.shape
    inp = Input((None, C, H, W))
    src = Activation('linear')(inp)
    mp = MaxPooling2D((pool, pool))(src)
    if mode is not None:
        mp.mode = mode
    inp.input_tensor = x
    for l in (inp, src, mp):
        l.forward(True)
    Y = np.ascontiguousarray(mp.output_tensor)
    G = f32(r.standard_normal(Y.shape)) if g is None else g
    mp.grads = G.copy()
    mp.backward()
    out = dict(X=x, Y=Y, G=G, dX=src.grads.copy(), pool=np.array(pool), mode=np.array(mp.mode))
    if mp.mode == 'im2col':
        out['argmax'] = mp.argmax_index.astype(np.int64)
    save(name, **out)


# --------------------------------------------------------------- dense ----
def dense_case(name, M, K, Nout, act, seed=0):
    np.random.seed(seed)
    r = np.random.default_rng(77 + seed)
    inp = Input((None, K))
    src = Activation('linear')(inp)
    d = Dense(Nout, activation=act, initializer='normal')(src)
    Wv, bv = d.variables
    Wv.output_tensor = f32(Wv.output_tensor)
    bv.output_tensor = f32(r.standard_normal((1, Nout)))
    X = f32(r.standard_normal((M, K)))
    inp.input_tensor = X
    for l in (inp, src, d):
        l.forward(True)
    Y = d.output_tensor.copy()
    G = f32(r.standard_normal(Y.shape))
    d.grads = G.copy()
    d.backward()
    save(name, X=X, W=Wv.output_tensor, b=bv.output_tensor, Y=Y, G=G, dW=Wv.grads, db=bv.grads,
         dX=src.grads.copy(), act=np.array(act))


# ------------------------------------------------------------------ bn ----
def bn_case(name, shape, seed=0):
    r = np.random.default_rng(500 + seed)
    inp = Input((None,) + shape[1:])
    src = Activation('linear')(inp)
    bn = BatchNormalization()(src)
    gam, bet = bn.variables
    C = shape[1]
    gam.output_tensor = f32(1 + 0.2 * r.standard_normal(C))
    bet.output_tensor = f32(0.3 * r.standard_normal(C))
    X = f32(2.0 * r.standard_normal(shape) + 0.5)
    inp.input_tensor = X
    for l in (inp, src, bn):
        l.forward(True)
    Y = np.ascontiguousarray(bn.output_tensor)
    G = f32(r.standard_normal(shape))
    bn.grads = G.copy()
    bn.backward()
    mm, mv = bn.moving_mean.copy(), bn.moving_variance.copy()
    # inference pass with the moving stats just produced
    X2 = f32(r.standard_normal(shape))
    inp.input_tensor = X2
    for l in (inp, src, bn):
        l.forward(False)
    save(name, X=X, gamma=gam.output_tensor, beta=bet.output_tensor, Y=Y, G=G, dX=src.grads.copy(),
         dgamma=gam.grads, dbeta=bet.grads, moving_mean=mm, moving_var=mv, X2=X2,
         Y2=np.ascontiguousarray(bn.output_tensor))


# ---------------------------------------------------------- loss / opt ----
def loss_case():
    r = np.random.default_rng(9)
    Z = f32(3 * r.standard_normal((32, 10)))
    labels = r.integers(0, 10, 32)
    labels[:10] = np.arange(10)
    Yoh = to_categorical(labels)
    from reference.utils.Activator import Softmax
    P = Softmax()._forward(Z.copy())
    L = SparseCategoricalCrossEntropy()
    save('loss_scce', Z=Z, Y=Yoh, P=P, loss=np.array(L.calc_loss(P, Yoh)), acc=np.array(L.calc_acc(P, Yoh)),
         dZ=L.backward(P, Yoh))


def optimizer_case():
    r = np.random.default_rng(11)
    w0 = [f32(r.standard_normal((7, 5))), f32(r.standard_normal((1, 5)))]
    gs = [[f32(r.standard_normal(w.shape)) for w in w0] for _ in range(3)]
    out = {}
    for oname, opt in (('sgd', StochasticGradientDescent(lr=0.05)), ('momentum', Momentum(lr=0.05, rho=0.9))):
        vs = [Variable(w.copy()) for w in w0]
        for v in vs:
            v.grads = np.zeros_like(v.output_tensor)
        for t in range(3):
            for v, g in zip(vs, gs[t]):
                v.grads = v.grads + g          # layers accumulate with +=
            opt.update(vs)
            for i, v in enumerate(vs):
                out['%s_w%d_t%d' % (oname, i, t)] = v.output_tensor.copy()
                out['%s_g%d_t%d' % (oname, i, t)] = v.grads.copy()
    for i, w in enumerate(w0):
        out['w%d' % i] = w
        for t in range(3):
            out['g%d_t%d' % (i, t)] = gs[t][i]
    save('optimizers', **out)


def adaptive_optimizer_case():
    """RMSprop / AdaGrad / AdaDelta / Adam (utils/Optimizers.py:61-157) on the same weights and
    gradient stream as optimizer_case.  None of them zeroes the gradients, and Adam counts two
    iterations per update (its own `+= 1` plus the base class's): both are reproduced as they are."""
    from reference.utils.Optimizers import RMSprop, AdaGrad, AdaDelta, Adam
    r = np.random.default_rng(11)
    w0 = [f32(r.standard_normal((7, 5))), f32(r.standard_normal((1, 5)))]
    gs = [[f32(r.standard_normal(w.shape)) for w in w0] for _ in range(4)]
    out = {}
    cases = (('rmsprop', RMSprop(lr=0.01, rho=0.9)), ('adagrad', AdaGrad(lr=0.05)), ('adadelta', AdaDelta(rho=0.95)),
             ('adam', Adam(lr=0.01, beta1=0.9, beta2=0.999)))
    for oname, opt in cases:
        vs = [Variable(w.copy()) for w in w0]
        for v in vs:
            v.grads = np.zeros_like(v.output_tensor)
        for t in range(4):
            for v, g in zip(vs, gs[t]):
                v.grads = v.grads + g          # layers accumulate with +=; these optimizers never reset
            opt.update(vs)
            for i, v in enumerate(vs):
                out['%s_w%d_t%d' % (oname, i, t)] = v.output_tensor.copy()
        out['%s_iterations' % oname] = np.array(opt.iterations)
    for i, w in enumerate(w0):
        out['w%d' % i] = w
        for t in range(4):
            out['g%d_t%d' % (i, t)] = gs[t][i]
    save('optimizers_adaptive', **out)


def batches_case():
    X = np.arange(23 * 2, dtype=np.float64).reshape(23, 2)
    Y = np.arange(23, dtype=np.float64).reshape(23, 1)
    out = {}
    for epoch in (0, 1):
        mb = get_batches(X, Y, 5, epoch, True)
        out['e%d_n' % epoch] = np.array(len(mb))
        out['e%d_y' % epoch] = np.concatenate([b[1] for b in mb])
        out['e%d_sizes' % epoch] = np.array([b[0].shape[0] for b in mb])
    save('get_batches', **out)


def colcase():
    r = np.random.default_rng(21)
    x = f32(r.standard_normal((2, 3, 7, 6)))
    col = im2col(x, (None, None, 3, 2), (3, 3), 2)
    dcol = f32(r.standard_normal((2 * 6 * 6, 3 * 9)))
    back = col2im((2, 3, 6, 6), (1, 1), (3, 3), 1, dcol)
    save('convcol', x=x, col=col, dcol=dcol, back=back)


# -------------------------------------------------------- trajectories ----
def traj_cfg2():
    np.random.seed(0)
    m = Sequential()
    m.add(Dense(500, activation='relu', n_in=784))
    m.add(Dense(10, activation='softmax'))
    m.compile(loss='sparse_categorical_crossentropy', optimizer='sgd')
    r = np.random.default_rng(1234)
    X = f32(r.random((384, 784)))
    Y = to_categorical(r.integers(0, 10, 384))
    quiet(m.fit, X, Y, batch_size=128, epochs=1, shuffle=False, validation_ratio=0.)
    save('traj_cfg2', train_loss=np.array(m.train_loss), train_acc=np.array(m.train_acc),
         sumW1=np.array(m.layers[0].variables[0].output_tensor.sum()),
         W2=m.layers[1].variables[0].output_tensor, xsum=np.array(X.sum()))
    # momentum variant, shuffled, 2 epochs: exercises get_batches + never-zeroed grads
    np.random.seed(0)
    m = Sequential()
    m.add(Dense(64, activation='relu', n_in=784))
    m.add(Dense(10, activation='softmax'))
    m.compile(loss='sparse_categorical_crossentropy', optimizer=Momentum(lr=0.01))
    quiet(m.fit, X[:200], Y[:200], batch_size=64, epochs=2, shuffle=True, validation_ratio=0.)
    save('traj_mlp_momentum', train_loss=np.array(m.train_loss), train_acc=np.array(m.train_acc),
         W2=m.layers[1].variables[0].output_tensor)


def traj_cfg1():
    np.random.seed(0)
    X_in = Input(shape=(None, 1, 28, 28))
    h = Conv2D(8, (2, 2), padding='VALID', initializer='normal', activation='relu')(X_in)
    h = MaxPooling2D((2, 2))(h)
    h.mode = 'im2col'                          # shipped 'reshape' raises on 27x27
    pool = h
    h = Flatten()(h)
    y = Dense(10, initializer='normal', activation='softmax')(h)
    model = Model(inputs=X_in, outputs=y)
    model.compile(optimizer='sgd', loss='sparse_categorical_crossentropy')
    r = np.random.default_rng(1234)
    X = f32(r.random((768, 1, 28, 28)))
    Y = to_categorical(r.integers(0, 10, 768))
    quiet(model.fit, X, Y, batch_size=256, epochs=1, shuffle=False, validation_ratio=0.)
    conv = X_in.outbound_layers[0]
    save('traj_cfg1', train_loss=np.array(model.train_loss), train_acc=np.array(model.train_acc),
         convW=conv.variables[0].output_tensor, convb=conv.variables[1].output_tensor,
         denseW_sum=np.array(y.variables[0].output_tensor.sum()), xsum=np.array(X.sum()))


def traj_cfg3_small():
    """VGG-style (BASELINE config 3) through the reference with ZeroPadding2D+VALID
    as SAME; full widths, batch 8, 2 SGD steps."""
    np.random.seed(0)
    m = Sequential()
    first = True
    for f in (64, 128, 256, 256):
        if first:
            zp = ZeroPadding2D((1, 1))
            zp.input_shape = (None, 3, 32, 32)
            m.add(_FirstPad(zp))
            first = False
        else:
            m.add(ZeroPadding2D((1, 1)))
        m.add(Conv2D(f, (3, 3), padding='VALID', activation='relu', initializer='normal'))
        m.add(BatchNormalization())
        m.add(MaxPooling2D((2, 2)))
    m.add(Flatten())
    m.add(Dense(10, activation='softmax'))
    m.compile(loss='sparse_categorical_crossentropy', optimizer='sgd')
    r = np.random.default_rng(1234)
    X = f32(r.standard_normal((16, 3, 32, 32)))
    Y = to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 6)]))
    quiet(m.fit, X, Y, batch_size=8, epochs=1, shuffle=False, validation_ratio=0.)
    convs = [l for l in m.layers if isinstance(l, Conv2D)]
    bns = [l for l in m.layers if isinstance(l, BatchNormalization)]
    save('traj_cfg3_small', train_loss=np.array(m.train_loss), train_acc=np.array(m.train_acc),
         conv1W=convs[0].variables[0].output_tensor, conv4W_sum=np.array(convs[3].variables[0].output_tensor.sum()),
         bn1_gamma=bns[0].variables[0].output_tensor, bn1_mm=bns[0].moving_mean, bn4_mv=bns[3].moving_variance,
         denseW=m.layers[-1].variables[0].output_tensor, xsum=np.array(X.sum()))


def traj_residual():
    """Functional residual graph (BASELINE config 4's building block) through the reference with
    the two oracle-side workarounds of SURVEY.md 8(c)(iii): operands of Add get a dummy
    output_tensor before Add()([a, b]) (layers/Base.py:358-359 dereferences it at build time), and
    trainable_variables is filtered to real Variables after compile (models.py:304-307).  1x1 VALID
    convolutions, because the reference cannot back-propagate through a mid-network ZeroPadding2D."""
    from reference.layers.Base import Add
    np.random.seed(0)
    X_in = Input(shape=(None, 3, 8, 8))
    stem = Conv2D(8, (1, 1), padding='VALID', activation='relu', initializer='normal')(X_in)
    r = Conv2D(8, (1, 1), padding='VALID', activation='linear', initializer='normal')(stem)
    r = BatchNormalization()(r)
    r = Activation('relu')(r)
    r = Conv2D(8, (1, 1), padding='VALID', activation='linear', initializer='normal')(r)
    r = BatchNormalization()(r)
    stem.output_tensor = np.zeros((1, 8, 8, 8)); r.output_tensor = np.zeros((1, 8, 8, 8))
    h = Add()([stem, r])
    stem.output_tensor = None; r.output_tensor = None
    h = Activation('relu')(h)
    h = MaxPooling2D((2, 2))(h)
    h = Flatten()(h)
    y = Dense(10, initializer='normal', activation='softmax')(h)
    model = Model(inputs=X_in, outputs=y)
    model.compile(optimizer='sgd', loss='sparse_categorical_crossentropy')
    model.trainable_variables = [v for v in model.trainable_variables if isinstance(v, Variable)]
    r_ = np.random.default_rng(1234)
    X = f32(r_.standard_normal((48, 3, 8, 8)))
    Y = to_categorical(np.concatenate([np.arange(10), r_.integers(0, 10, 38)]))
    quiet(model.fit, X, Y, batch_size=16, epochs=1, shuffle=False, validation_ratio=0.)
    save('traj_residual', train_loss=np.array(model.train_loss), train_acc=np.array(model.train_acc),
         stemW=stem.variables[0].output_tensor, denseW=y.variables[0].output_tensor, xsum=np.array(X.sum()))


class _FirstPad(object):
    """Sequential.compile calls layers[0].connect(None); ZeroPadding2D.connect
    dereferences prev_layer (layers/Convolution.py:219).  This thin proxy gives
    the first ZeroPadding2D an Input to connect to, nothing else."""
    def __new__(cls, zp):
        inp = Input(zp.input_shape)
        orig = zp.connect

        def connect(prev):
            orig(inp)
            zp.inbounds = []
        zp.connect = connect
        return zp


# ---------------------------------------------------------- initializers ----
def initializer_case():
    """utils/Initializers.py under fixed seeds (its fan computation takes the sqrt(prod) branch for
    every shape tuple, :12-22; the host-side initializers must reproduce that)."""
    from reference.utils.Initializers import get_initializer
    out = {}
    for i, (name, size) in enumerate([('glorot_uniform', (4, 6)), ('he_normal', (8, 3, 3, 3)), ('lecun_uniform', (5, 7)),
                                      ('he_uniform', (16, 8, 3, 3)), ('glorot_normal', (12, 12)), ('lecun_normal', (7, 3)),
                                      ('orthogonal', (6, 4)), ('orthogonal', (5, 5)), ('uniform', (3, 4)), ('normal', (3, 4))]):
        np.random.seed(10 + i)
        out['%d_%s' % (i, name)] = get_initializer(name)(size)
    save('initializers', **out)


# ------------------------------------------------------- embedding, lstm ----
def recurrent_case():
    """Embedding forward/backward and LSTM forward/backward of the unmodified reference
    (layers/Embeddings.py, layers/Recurrent.py:216-507).  The LSTM gradients stored here are the
    reference's own — gate gradients concatenated in the wrong order (:358) — and pin the oracle's
    reference_gate_order=True branch."""
    from reference.layers.Embeddings import Embedding
    from reference.layers.Recurrent import LSTM
    np.random.seed(0)
    r = np.random.default_rng(4242)
    ids = r.integers(0, 50, (6, 9))
    inp = Input((None, 9))
    emb = Embedding(50, 8)
    emb(inp)                                   # (the reference's Embedding.__call__ returns None)
    Wv, = emb.variables
    Wv.output_tensor = f32(Wv.output_tensor)
    inp.input_tensor = ids
    for l in (inp, emb):
        l.forward(True)
    Y = emb.output_tensor.copy()
    G = f32(r.standard_normal(Y.shape))
    emb.grads = G.copy()
    emb.backward()
    save('embedding', ids=ids, W=Wv.output_tensor, Y=Y, G=G, dW=Wv.grads)
    out = {}
    for rs in (True, False):
        np.random.seed(1)
        inp = Input((None, 7, 5))
        src = Activation('linear')(inp)
        lstm = LSTM(6, return_sequences=rs)(src)
        W, b = lstm.variables
        W.output_tensor = f32(W.output_tensor)
        b.output_tensor = f32(b.output_tensor + 0.1 * r.standard_normal(b.output_tensor.shape))
        X = f32(r.standard_normal((4, 7, 5)))
        inp.input_tensor = X
        for l in (inp, src, lstm):
            l.forward(True)
        Yl = np.array(lstm.output_tensor)
        Gl = f32(r.standard_normal(Yl.shape))
        lstm.grads = Gl.copy()
        lstm.backward()
        k = 'seq' if rs else 'last'
        out.update({'X_' + k: X, 'W_' + k: W.output_tensor.copy(), 'b_' + k: b.output_tensor.copy(), 'Y_' + k: Yl, 'G_' + k: Gl,
                    'dW_ref_' + k: W.grads.copy(), 'db_ref_' + k: b.grads.copy(), 'dX_ref_' + k: np.array(src.grads)})
    save('lstm', **out)


if __name__ == '__main__':
    if len(sys.argv) > 1:                      # python make_golden.py recurrent_case ... : only these
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    conv_case('conv_katA', 2, 3, 8, 8, 4, 3, 1, 1, 'relu')
    conv_case('conv_valid_s2', 3, 5, 9, 11, 6, 3, 2, 0, 'linear', seed=1)
    conv_case('conv_cfg1_shape', 4, 1, 28, 28, 8, 2, 1, 0, 'relu', seed=2)
    conv_case('conv_c64_same', 2, 64, 8, 8, 128, 3, 1, 1, 'relu', seed=3)
    conv_case('conv_5x5_same', 2, 4, 10, 10, 8, 5, 1, 2, 'linear', seed=4)
    r = np.random.default_rng(1234)
    _ = r.standard_normal((2, 3, 8, 8)); _ = r.standard_normal((2, 4, 8, 8))   # KAT B continues KAT A's stream
    Xp = r.integers(0, 4, (2, 3, 6, 6)).astype(np.float64)
    Gp = r.standard_normal((2, 3, 3, 3))
    pool_case('pool_katB_reshape', Xp, 2, None, g=Gp)
    pool_case('pool_katB_im2col', Xp, 2, 'im2col', g=Gp)
    r2 = np.random.default_rng(5)
    pool_case('pool_odd27_im2col', f32(r2.standard_normal((2, 8, 27, 27))), 2, 'im2col')
    pool_case('pool_32_reshape', f32(r2.standard_normal((2, 16, 32, 32))), 2, None)
    pool_case('pool_3x3_ties', r2.integers(0, 3, (2, 4, 9, 9)).astype(np.float64), 3, None)
    dense_case('dense_relu', 16, 40, 24, 'relu')
    dense_case('dense_softmax', 12, 33, 10, 'softmax', seed=1)
    dense_case('dense_linear', 5, 7, 3, 'linear', seed=2)
    bn_case('bn_4d', (4, 6, 5, 5))
    bn_case('bn_2d', (16, 12), seed=1)
    loss_case()
    optimizer_case()
    adaptive_optimizer_case()
    batches_case()
    colcase()
    traj_cfg2()
    traj_cfg1()
    traj_cfg3_small()
    traj_residual()
