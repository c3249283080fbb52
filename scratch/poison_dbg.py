import sys, os, numpy as np
os.environ['SN_DEBUG_POISON'] = '1'
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import shinnosuke_oracle as O
from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
from shinnosuke_b200.device import as_device
r = np.random.default_rng(5)
X = r.standard_normal((16, 3, 32, 32))
Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 6)]))
for fuse in (True, False):
    np.random.seed(0)
    m = Sequential()
    m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Conv2D(128, (3, 3), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Flatten()); m.add(Dense(10, activation='softmax'))
    m.compile('sgd', 'sparse_categorical_crossentropy', fuse=fuse)
    m._materialize()
    y_hat = m._forward_device(as_device(X), True)
    print('fuse', fuse, 'forward nan:', np.isnan(np.asarray(y_hat)).sum())
    m.calc_gradients(y_hat, Y)
    for li, layer in enumerate(m.layers):
        for v in layer.variables:
            g = np.asarray(v.grads)
            print('   layer', li, type(layer).__name__, v.name, 'grad nan', int(np.isnan(g).sum()), 'of', g.size)
        if layer.grads is not None and li > 0:
            try:
                g = np.asarray(layer.grads)
                print('   layer', li, type(layer).__name__, 'dX-in (layer.grads) nan', int(np.isnan(g).sum()), 'of', g.size)
            except Exception as e:
                print('   layer', li, 'grads unreadable', e)

print('---- fit path')
for graph in (False, True):
    np.random.seed(0)
    m = Sequential()
    m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Flatten()); m.add(Dense(10, activation='softmax'))
    m.compile('sgd', 'sparse_categorical_crossentropy')
    m.use_cuda_graph = graph
    m.fit(X, Y, batch_size=8, epochs=1, shuffle=False, validation_ratio=0., verbose=0)
    print('graph', graph, 'losses', m.train_loss)
    for li, layer in enumerate(m.layers):
        for v in layer.variables:
            w = np.asarray(v.output_tensor)
            print('   layer', li, type(layer).__name__, v.name, 'weight nan', int(np.isnan(w).sum()), 'of', w.size)
    inp = m._input_layer()
    for k, t in inp._bufs.items():
        print('   input buf', k, 'nan', int(np.isnan(np.asarray(t.buf.cpu().numpy())).sum()))
