"""-m gpu: save / load (reference models.py:208-220 Sequential, :530-543 Model) keeps what the
reference's pickle keeps — weights, accumulated gradients, optimizer state — as NumPy arrays only,
so that a resumed run continues the SAME trajectory; plus host-side behaviours around the captured
step graph (hyper-parameter changes, Embedding id semantics)."""
import os
import pickle

import numpy as np
import pytest

from conftest import rel_err
from oracle import shinnosuke_oracle as O

pytestmark = pytest.mark.gpu


def _mlp(optimizer, hidden=64):
    from shinnosuke_b200 import Sequential, Dense
    np.random.seed(0)
    m = Sequential()
    m.add(Dense(hidden, activation='relu', n_in=784))
    m.add(Dense(10, activation='softmax'))
    m.compile(loss='sparse_categorical_crossentropy', optimizer=optimizer)
    return m


def _data(n=256):
    r = np.random.default_rng(7)
    X = r.random((n, 784)).astype(np.float32).astype(np.float64)
    Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, n - 10)]))
    return X, Y


@pytest.mark.parametrize('optimizer', ['sgd', 'momentum', 'rmsprop', 'adagrad', 'adadelta', 'adam'])
def test_resume_from_checkpoint_continues_the_trajectory(tmp_path, optimizer):
    """4 minibatches, save, load into a fresh model, 4 more == 8 uninterrupted minibatches: Momentum's
    velocity and never-zeroed gradients, the adaptive optimizers' moments and Adam's device tick (its
    bias correction) all come back.  The uninterrupted run replays a captured graph from step 3 on, the
    resumed one runs its first step eagerly again: also a graph-vs-eager equivalence check."""
    from shinnosuke_b200 import Sequential
    X, Y = _data(512)
    ref = _mlp(optimizer)
    ref.fit(X, Y, batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    a = _mlp(optimizer)
    a.fit(X[:256], Y[:256], batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    path = str(tmp_path / 'ckpt')
    a.save(path)
    # the file is NumPy only (unpickles with no CUDA state) and about parameter-sized: weights (float64 host
    # copies) + the optimizer's state, not one copy of the arena per variable
    n_params = sum(v.size for v in a.trainable_variables)
    assert os.path.getsize(path + '.pkl') < 8 * n_params * 4 + 65536
    b = Sequential()
    b.load(path)
    assert b.optimizer.iterations == a.optimizer.iterations == {'adam': 8}.get(optimizer, 4)
    b.fit(X[256:], Y[256:], batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.allclose(a.train_loss + b.train_loss, ref.train_loss, rtol=2e-6), (a.train_loss + b.train_loss, ref.train_loss)
    for v, w in zip(b.trainable_variables, ref.trainable_variables):
        assert rel_err(np.asarray(v.output_tensor), np.asarray(w.output_tensor)) < 2e-6


def test_checkpoint_holds_no_device_objects(tmp_path):
    m = _mlp('momentum')
    X, Y = _data()
    m.fit(X, Y, batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    path = str(tmp_path / 'm')
    m.save(path)

    class NoTorch(pickle.Unpickler):
        def find_class(self, module, name):
            assert not module.startswith('torch'), 'the checkpoint references %s.%s' % (module, name)
            return super().find_class(module, name)
    with open(path + '.pkl', 'rb') as f:
        layers, optimizer, loss = NoTorch(f).load()
    assert layers[0].variables[0]._dev is None and isinstance(layers[0].variables[0]._host, np.ndarray)


def test_functional_model_save_load_roundtrip(tmp_path):
    """Model.save / Model.load (reference models.py:530-543) on a residual graph."""
    from shinnosuke_b200 import Model, Input, Conv2D, BatchNormalization, Activation, Add, MaxPooling2D, Flatten, Dense
    np.random.seed(0)
    x_in = Input(shape=(None, 3, 8, 8))
    stem = Conv2D(8, (3, 3), padding='SAME', activation='relu')(x_in)
    r = BatchNormalization()(Conv2D(8, (3, 3), padding='SAME', activation='linear')(stem))
    h = MaxPooling2D((2, 2))(Activation('relu')(Add()([stem, r])))
    y = Dense(10, activation='softmax')(Flatten()(h))
    m = Model(inputs=x_in, outputs=y)
    m.compile(optimizer='momentum', loss='sparse_categorical_crossentropy')
    rr = np.random.default_rng(3)
    X = rr.standard_normal((64, 3, 8, 8)).astype(np.float32).astype(np.float64)
    Y = O.to_categorical(np.concatenate([np.arange(10), rr.integers(0, 10, 54)]))
    m.fit(X, Y, batch_size=16, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    acc, loss = m.evaluate(X, Y)
    path = str(tmp_path / 'res')
    m.save(path)
    m2 = Model()
    m2.load(path)
    acc2, loss2 = m2.evaluate(X, Y)
    assert acc2 == acc and abs(loss2 - loss) < 1e-6 * abs(loss)
    # and both continue identically
    m.fit(X, Y, batch_size=16, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    m2.fit(X, Y, batch_size=16, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.allclose(m.train_loss[-4:], m2.train_loss[-4:], rtol=1e-5)


def test_learning_rate_change_is_not_ignored_by_the_captured_graph():
    """The reference reads self.lr on every update (utils/Optimizers.py:29-32); the captured step graph
    bakes it in as a kernel scalar, so the graph key carries the optimizer's hyper-parameters."""
    X, Y = _data(256)
    m = _mlp('sgd')
    m.fit(X, Y, batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)        # captures at lr = 0.01
    n_graphs = len(m._graphs)
    w0 = np.asarray(m.trainable_variables[0].output_tensor).copy()
    m.optimizer.lr = 0.0
    m.fit(X, Y, batch_size=64, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    assert np.array_equal(np.asarray(m.trainable_variables[0].output_tensor), w0), 'an lr of 0 still moved the weights'
    assert len(m._graphs) > n_graphs


def test_embedding_ids_follow_numpy_indexing():
    """W[idx] in the reference (layers/Embeddings.py:60): negative ids wrap, ids outside the table raise."""
    from shinnosuke_b200 import Embedding, Input
    from shinnosuke_b200.device import as_device
    from shinnosuke_b200._lib import lib
    np.random.seed(0)
    inp = Input((None, 6))
    emb = Embedding(10, 8)(inp)
    W = np.asarray(emb.variables[0].output_tensor)
    ids = np.array([[0, 9, -1, -10, 3, -3]], dtype=np.float64)
    inp.input_tensor = ids
    inp.forward(True); emb.forward(True)
    assert np.array_equal(np.asarray(emb.output_tensor)[0], W[ids[0].astype(int)].astype(np.float32))
    assert lib.embedding_oob_count() == 0
    g = np.ones((1, 6, 8))
    emb._seed_grads(as_device(g)); emb.backward()
    ref = np.zeros_like(W); np.add.at(ref, ids[0].astype(int), g[0])
    assert np.array_equal(np.asarray(emb.variables[0].grads), ref)
    inp.input_tensor = np.array([[0, 10, -11, 1, 2, 3]], dtype=np.float64)
    inp.forward(True); emb.forward(True)
    with pytest.raises(IndexError):
        Embedding.check_ids()
    assert lib.embedding_oob_count() == 0            # the check clears the counter
    with pytest.raises(ValueError):
        Embedding(1 << 24, 4)
