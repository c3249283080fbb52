"""CPU: host-side logic of the reference-facing API (shape inference, parameter creation order,
graph ordering, batching) checked against the golden fixtures.  No GPU work."""
import numpy as np
import pytest

from conftest import load_golden
import shinnosuke_b200 as sb
from shinnosuke_b200 import (Sequential, Model, Input, Dense, Conv2D, MaxPooling2D, Flatten, BatchNormalization,
                             ZeroPadding2D, Activation)
from shinnosuke_b200.utils.MiniBatch import get_batches, batch_slices
from shinnosuke_b200.utils.Preprocess import to_categorical
from shinnosuke_b200.utils import Initializers as I
from oracle import shinnosuke_oracle as O


def test_initial_weights_match_reference_stream():
    """np.random.seed(0) + compile must reproduce the reference's initial weights (KAT A values)."""
    np.random.seed(0)
    x = Input((None, 3, 8, 8))
    c = Conv2D(4, (3, 3), padding='SAME', activation='relu', initializer='normal')(x)
    W, b = c.variables
    assert np.allclose(W.output_tensor[0, 0, 0, :], [0.176405234597, 0.040015720837, 0.097873798411], atol=1e-12)
    assert np.allclose(b.output_tensor[0, :2], [1.92294202648, 1.480514791434], atol=1e-11)
    assert c.output_shape == (None, 4, 8, 8) and c.pad_size == (1, 1)


def test_sequential_compile_shapes_cfg3():
    np.random.seed(0)
    m = Sequential()
    m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    for f in (128, 256, 256):
        m.add(Conv2D(f, (3, 3), padding='SAME', activation='relu')); m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Flatten()); m.add(Dense(10, activation='softmax'))
    m.compile(optimizer='sgd', loss='sparse_categorical_crossentropy', precision='tf32')
    assert m.layers[-1].output_shape == (None, 10)
    assert m.layers[-2].output_shape == (None, 1024)
    assert sum(v.size for v in m.trainable_variables) == 972554          # SURVEY.md / BASELINE.md
    assert all(l.precision == 'tf32' for l in m.layers)
    # same RNG consumption as the oracle's builder (which is pinned to the reference)
    np.random.seed(0)
    spec, ishape = O.spec_cfg3()
    net = O.OracleNet(spec, ishape)
    ours = [np.asarray(v.output_tensor) for v in m.trainable_variables]
    assert len(ours) == len(net.params)
    for a, b in zip(ours, net.params):
        assert a.shape == b.shape and np.array_equal(a, b)


def test_functional_graph_order_cfg1():
    np.random.seed(0)
    X_in = Input(shape=(None, 1, 28, 28))
    h = Conv2D(8, (2, 2), padding='VALID', initializer='normal', activation='relu')(X_in)
    h = MaxPooling2D((2, 2))(h)
    h = Flatten()(h)
    y = Dense(10, initializer='normal', activation='softmax')(h)
    model = Model(inputs=X_in, outputs=y)
    model.compile(optimizer='sgd', loss='sparse_categorical_crossentropy')
    assert [l.__class__.__name__ for l in model.forward_graph] == ['Input', 'Conv2D', 'MaxPooling2D', 'Flatten', 'Dense']
    assert [l.__class__.__name__ for l in model.backward_graph] == ['Dense', 'Flatten', 'MaxPooling2D', 'Conv2D', 'Input']
    assert sum(v.size for v in model.trainable_variables) == 13570
    assert model.forward_graph[2].output_shape == (None, 8, 13, 13) and model.forward_graph[2].mode == 'reshape'
    g = load_golden('traj_cfg1')     # not the initial weights, but shapes must agree
    assert model.forward_graph[1].variables[0].output_shape == g['convW'].shape


def test_mlp_param_count_matches_readme():
    m = Sequential([Dense(500, activation='relu', n_in=784), Dense(10, activation='softmax')])
    m.compile('sgd', 'sparse_categorical_crossentropy')
    assert sum(v.size for v in m.trainable_variables) == 397510            # README.md:81


def test_pool_mode_selection_and_floor():
    inp = Input((None, 8, 27, 27))
    p = MaxPooling2D((2, 2))(inp)
    assert p.mode == 'reshape' and p.output_shape == (None, 8, 13, 13)
    p2 = MaxPooling2D((3, 3), stride=2)(inp)
    assert p2.mode == 'im2col' and p2.output_shape == (None, 8, 13, 13)


def test_error_behaviour_mirrors_reference():
    with pytest.raises(TypeError):
        Conv2D(4, (3, 3), padding='full')(Input((None, 1, 8, 8)))
    with pytest.raises(ValueError):
        Sequential([Dense(3)]).compile('sgd', 'sparse_categorical_crossentropy')
    with pytest.raises(ValueError):
        sb.get_optimizer(3.5)
    with pytest.raises(ValueError):
        Flatten(out_dim=0)
    with pytest.raises(AssertionError):
        MaxPooling2D((2, 3))


def test_get_batches_matches_reference():
    g = load_golden('get_batches')
    X = np.arange(46, dtype=np.float64).reshape(23, 2)
    Y = np.arange(23, dtype=np.float64).reshape(23, 1)
    for epoch in (0, 1):
        mb = get_batches(X, Y, 5, epoch, True)
        assert len(mb) == int(g['e%d_n' % epoch])
        assert np.array_equal(np.concatenate([b[1] for b in mb]), g['e%d_y' % epoch])
    assert batch_slices(23, 5) == [(0, 5), (5, 10), (10, 15), (15, 20), (20, 23)]
    assert batch_slices(0, 5) == []


def test_to_categorical():
    assert np.array_equal(to_categorical(np.array([0, 2, 1])), np.eye(3)[[0, 2, 1]])
    assert np.array_equal(to_categorical(np.array([[1], [0]])), np.eye(2)[[1, 0]])
    with pytest.raises(ValueError):
        to_categorical(np.zeros((2, 2, 2), dtype=int))


def test_initializers_consume_rng_like_reference():
    """Every initializer against the reference's own output under the same seed
    (tests/golden/initializers.npz) — including its fan computation, which takes the sqrt(prod)
    branch for every shape tuple (utils/Initializers.py:12-22)."""
    from conftest import load_golden
    g = load_golden('initializers')
    assert len(g) == 10
    for key, ref in g.items():
        i, name = key.split('_', 1)
        np.random.seed(10 + int(i))
        assert np.array_equal(I.get_initializer(name)(ref.shape), ref), key
    np.random.seed(3)
    a = I.get_initializer('glorot_uniform')((4, 6))
    np.random.seed(3)
    s = np.sqrt(6. / (4 + 4))                       # int(sqrt(24)) = 4 for fan_in and fan_out
    assert np.array_equal(a, np.random.uniform(-s, s, size=(4, 6)))


def test_no_cpu_fallback():
    """Without a CUDA device the compute path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    m = Sequential([Dense(4, activation='softmax', n_in=3)])
    m.compile('sgd', 'sparse_categorical_crossentropy')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m.fit(np.zeros((2, 3)), np.eye(4)[:2], batch_size=2, epochs=1, validation_ratio=0.)
    with pytest.raises(RuntimeError):
        m.predict(np.zeros((2, 3)))


def test_allreduce_overlap_split_cfg3():
    """Data-parallel backward starts reducing as soon as the arena tail is final: for config 3 that
    is after conv3.backward (conv3, conv4, their BatchNorms and the Dense layer = 92 % of the
    parameters), leaving conv2 / conv1 to overlap the NCCL call (models._overlap_split)."""
    from types import SimpleNamespace
    np.random.seed(0)
    m = Sequential()
    m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    for f in (128, 256, 256):
        m.add(Conv2D(f, (3, 3), padding='SAME', activation='relu')); m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Flatten()); m.add(Dense(10, activation='softmax'))
    m.compile(optimizer='sgd', loss='sparse_categorical_crossentropy')
    offs, total = [], 0
    for v in m.trainable_variables:               # device.Arena's layout rule: 32-float aligned, forward order
        offs.append(total)
        total += (v.size + 31) // 32 * 32
    m._arena = SimpleNamespace(offsets=offs, param_floats=total, ALIGN=32)
    m._split = None
    k, lo = m._overlap_split()
    order = m._backward_order()
    assert order[k] is m.layers[6]                # conv3 (layers: conv,bn,pool x4, flatten, dense)
    conv3_w = m.trainable_variables.index(m.layers[6].variables[0])
    assert lo == offs[conv3_w]
    assert (total - lo) / total > 0.9
    # nothing that still has to run its backward owns arena space at or above lo
    for layer in order[k + 1:]:
        for v in layer.variables:
            i = m.trainable_variables.index(v)
            assert offs[i] + v.size <= lo
    m._arena = None


def test_sequential_compile_cfg5_matches_the_oracle_builder():
    """Config 5 (Embedding -> LSTM -> Dense) compiles on the host with the reference's shapes and the
    same RNG consumption as the oracle's builder (LSTM weights: per gate glorot-uniform input block,
    then orthogonal recurrent block; forget-gate bias one; layers/Recurrent.py:229-258)."""
    from shinnosuke_b200 import Embedding, LSTM
    np.random.seed(0)
    m = Sequential()
    m.add(Embedding(1000, 64, input_length=64)); m.add(LSTM(128)); m.add(Dense(2, activation='softmax'))
    m.compile(optimizer='sgd', loss='sparse_categorical_crossentropy', precision='bf16')
    assert m.layers[0].output_shape == (None, 64, 64)
    assert m.layers[1].output_shape == (None, 128)
    assert m.layers[2].output_shape == (None, 2)
    assert sum(v.size for v in m.trainable_variables) == 1000 * 64 + 192 * 512 + 512 + 128 * 2 + 2
    np.random.seed(0)
    spec, ishape = O.spec_cfg5()
    net = O.OracleNet(spec, ishape)
    ours = [np.asarray(v.output_tensor) for v in m.trainable_variables]
    assert len(ours) == len(net.params)
    for a, b in zip(ours, net.params):
        assert a.shape == b.shape and np.array_equal(a, b)
    W, b = m.layers[1].variables
    assert np.array_equal(np.asarray(b.output_tensor)[0, :128], np.ones(128)) and not np.asarray(b.output_tensor)[0, 128:].any()
    with pytest.raises(NotImplementedError):
        LSTM(8, return_state=True)
    with pytest.raises(NotImplementedError):
        LSTM(8, activation='relu')


def test_lstm_weight_storage_layout_round_trip():
    """('lstm_w', H): logical (H + D, 4H) rows [recurrent; input] <-> device storage [4H][H] then [4H][D]."""
    from shinnosuke_b200.device import _to_storage, _from_storage
    r = np.random.default_rng(0)
    H, D = 6, 5
    W = r.standard_normal((H + D, 4 * H)).astype(np.float32).astype(np.float64)
    st = _to_storage(W, ('lstm_w', H))
    assert st.dtype == np.float32 and st.shape == (4 * H * (H + D),)
    wh, wx = st[:4 * H * H].reshape(4 * H, H), st[4 * H * H:].reshape(4 * H, D)
    assert np.array_equal(wh, W[:H].T) and np.array_equal(wx, W[H:].T)
    assert np.array_equal(_from_storage(st, W.shape, ('lstm_w', H)), W)


def test_relu_folding_and_fusion_flags_cfg4():
    """compile() folds an Activation('relu') into the BatchNormalization / Add that feeds it when that
    layer has no other consumer, and leaves everything alone with fuse=False."""
    from shinnosuke_b200 import Model, Input, Activation, Add
    def build(fuse):
        np.random.seed(0)
        x_in = Input(shape=(None, 3, 8, 8))
        h = Conv2D(8, (3, 3), padding='SAME', activation='relu')(x_in)
        r = Conv2D(8, (3, 3), padding='SAME', activation='linear')(h)
        bn1 = BatchNormalization()(r)
        a1 = Activation('relu')(bn1)
        r = Conv2D(8, (3, 3), padding='SAME', activation='linear')(a1)
        bn2 = BatchNormalization()(r)
        add = Add()([h, bn2])
        a2 = Activation('relu')(add)
        y = Dense(10, activation='softmax')(Flatten()(MaxPooling2D((2, 2))(a2)))
        m = Model(inputs=x_in, outputs=y)
        m.compile(optimizer='sgd', loss='sparse_categorical_crossentropy', precision='bf16', fuse=fuse)
        return bn1, a1, bn2, add, a2
    bn1, a1, bn2, add, a2 = build(True)
    assert bn1._relu_out is a1 and a1._folded_into is bn1
    assert add._relu_out is a2 and a2._folded_into is add
    assert bn2._relu_out is None                      # its consumer is the Add, not a ReLU
    bn1, a1, bn2, add, a2 = build(False)
    assert bn1._relu_out is None and a1._folded_into is None and add._relu_out is None and a2._folded_into is None


def test_layers_with_every_initializer_pickle():
    """save() pickles the layers (models.py:208-220), and layers hold their initializer objects: every
    initializer — the fan-scaled ones included — must survive a round trip, and a config-5 model's
    layer list must come back with the same weights and storage layouts."""
    import pickle
    from shinnosuke_b200 import Embedding, LSTM
    for name in ('zeros', 'ones', 'uniform', 'normal', 'glorot_uniform', 'glorot_normal', 'he_uniform', 'he_normal',
                 'lecun_uniform', 'lecun_normal', 'orthogonal'):
        init = pickle.loads(pickle.dumps(I.get_initializer(name)))
        np.random.seed(1); a = init((6, 4))
        np.random.seed(1); b = I.get_initializer(name)((6, 4))
        assert np.array_equal(a, b), name
    np.random.seed(0)
    m = Sequential()
    m.add(Embedding(50, 8, input_length=6)); m.add(LSTM(64)); m.add(Dense(2, activation='softmax', initializer='glorot_uniform'))
    m.compile('sgd', 'sparse_categorical_crossentropy')
    layers, opt, loss = pickle.loads(pickle.dumps([m.layers, m.optimizer, m.loss]))
    for a, b in zip(layers, m.layers):
        assert type(a) is type(b)
        for va, vb in zip(a.variables, b.variables):
            assert va.layout == vb.layout and np.array_equal(np.asarray(va.output_tensor), np.asarray(vb.output_tensor))


def test_gradient_hand_over_rules():
    """Layer._hand_over_grads: a pass-through layer gives its gradient buffer to an operand only when that
    operand has no other consumer and nothing has been written to its gradient yet; Add serves the
    adopting operand last (layers/Base.py)."""
    from shinnosuke_b200.layers.Base import Layer, Add

    class Buf(object):
        def __init__(self, shape, layout='nhwc'):
            self.shape, self.layout, self.size = tuple(shape), layout, int(np.prod(shape))
            self.log = []

        def view(self, shape, layout):
            v = Buf(shape, layout); v.base = self
            return v

        def copy_from(self, other):
            self.log.append(('copy', other))

        def add_(self, other):
            self.log.append(('add', other))

    def layer(n_consumers):
        l = Layer()
        l.outbound_layers = [object()] * n_consumers
        l.grads = Buf((2, 3, 4, 4))
        return l

    src, g = Layer(), Buf((2, 3, 4, 4))
    single, shared, written = layer(1), layer(2), layer(1)
    written._grads_written = True
    assert src._hand_over_grads(single, g) and single.grads is g and single._grads_written
    own = shared.grads
    assert not src._hand_over_grads(shared, g) and shared.grads is own
    assert not src._hand_over_grads(written, g)
    frozen = layer(1); frozen.require_grads = False
    assert not src._hand_over_grads(frozen, g)
    # Add: the multi-consumer operand gets its copy first, the single-consumer operand adopts the buffer
    a, b = layer(2), layer(1)
    add = Add()
    add.inbounds = [b, a]
    add.grads = Buf((2, 3, 4, 4))
    a_own = a.grads
    add.backward()
    assert b.grads is add.grads and a.grads is a_own and a_own.log == [('copy', add.grads)]
    # the same layer on both sides: one copy, one accumulation, no hand-over
    c = layer(2)
    add2 = Add(); add2.inbounds = [c, c]; add2.grads = Buf((2, 3, 4, 4))
    c_own = c.grads
    add2.backward()
    assert c.grads is c_own and [k for k, _ in c_own.log] == ['copy', 'add']
