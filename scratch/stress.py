import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import shinnosuke_oracle as O
from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
r = np.random.default_rng(5)
X = r.standard_normal((32, 3, 32, 32))
Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 22)]))
def run(fuse, graph=True, shuffle=True):
    np.random.seed(0)
    m = Sequential()
    m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Conv2D(128, (3, 3), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Flatten()); m.add(Dense(10, activation='softmax'))
    m.compile('sgd', 'sparse_categorical_crossentropy', fuse=fuse)
    m.use_cuda_graph = graph
    m.fit(X, Y, batch_size=16, epochs=2, shuffle=shuffle, validation_ratio=0., verbose=0)
    return np.array(m.train_loss), np.asarray(m.layers[0].variables[0].output_tensor)
mode = sys.argv[1] if len(sys.argv) > 1 else 'default'
kw = dict(default={}, nograph=dict(graph=False), noshuffle=dict(shuffle=False))[mode]
ref = None
for trial in range(12):
    for fuse in (True, False):
        l, w = run(fuse, **kw)
        if ref is None: ref = (l, w)
        e = np.abs(w - ref[1]) / np.abs(ref[1]).max()
        print(mode, 'trial', trial, 'fuse', fuse, 'loss', ['%.6f' % v for v in l], 'w q99 %.1e max %.1e' % (np.quantile(e, 0.99), e.max()))
