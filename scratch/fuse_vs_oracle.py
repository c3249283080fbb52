import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import shinnosuke_oracle as O
from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
from test_gpu_train import _data
r, X = _data((32, 3, 32, 32), kind='normal')
X = (0.25 * X).astype(np.float32).astype(np.float64)
Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 22)]))
np.random.seed(0)
spec = [('conv', 64, 3, 1, 'SAME', 'relu'), ('bn',), ('pool', 2), ('conv', 128, 3, 1, 'SAME', 'relu'), ('bn',), ('pool', 2), ('flatten',), ('dense', 10, 'softmax')]
net = O.OracleNet(spec, (None, 3, 32, 32), optimizer='sgd')
net.fit(X, Y, 16, epochs=2, shuffle=True)
print('oracle losses', net.train_loss)
for trial in range(2):
  for fuse in (True, False):
    np.random.seed(0)
    m = Sequential()
    m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Conv2D(128, (3, 3), padding='SAME', activation='relu'))
    m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
    m.add(Flatten()); m.add(Dense(10, activation='softmax'))
    m.compile('sgd', 'sparse_categorical_crossentropy', fuse=fuse)
    m.fit(X, Y, batch_size=16, epochs=2, shuffle=True, validation_ratio=0., verbose=0)
    errs = [np.abs(np.asarray(v.output_tensor) - p).max() / np.abs(p).max() for v, p in zip(m.trainable_variables, net.params)]
    print('fuse', fuse, 'losses', m.train_loss, 'w err vs oracle', ['%.1e' % e for e in errs])
