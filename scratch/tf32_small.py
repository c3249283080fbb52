import sys, numpy as np
sys.path.insert(0, '.')
from oracle import shinnosuke_oracle as O
from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
r = np.random.default_rng(7)
X = r.standard_normal((64, 3, 16, 16)); Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 54)]))
np.random.seed(0)
spec = [('conv', 16, 3, 1, 'SAME', 'relu'), ('bn',), ('pool', 2), ('conv', 32, 3, 1, 'SAME', 'relu'), ('bn',), ('flatten',), ('dense', 10, 'softmax')]
net = O.OracleNet(spec, (None, 3, 16, 16), optimizer='sgd')
for ep in range(2):
    for a in range(0, 64, 32):
        net.step(X[a:a + 32], Y[a:a + 32])
for prec in ('fp32', 'tf32', 'bf16'):
    for fuse in (True, False):
        np.random.seed(0)
        m = Sequential()
        m.add(Conv2D(16, (3, 3), input_shape=(None, 3, 16, 16), padding='SAME', activation='relu'))
        m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
        m.add(Conv2D(32, (3, 3), padding='SAME', activation='relu'))
        m.add(BatchNormalization())
        m.add(Flatten()); m.add(Dense(10, activation='softmax'))
        m.compile('sgd', 'sparse_categorical_crossentropy', precision=prec, fuse=fuse)
        m.fit(X, Y, batch_size=32, epochs=2, shuffle=False, validation_ratio=0., verbose=0)
        ws = [np.asarray(v.output_tensor) for v in m.trainable_variables]
        err = [np.abs(w - p).max() / max(np.abs(p).max(), 1e-12) for w, p in zip(ws, net.params)]
        lerr = [abs(a - b) / abs(b) for a, b in zip(m.train_loss, net.train_loss)]
        print(prec, 'fuse', fuse, 'w err', ['%.1e' % e for e in err], 'loss err', ['%.1e' % e for e in lerr])
