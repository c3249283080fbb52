import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import shinnosuke_oracle as O
from conftest import rel_err
from shinnosuke_b200 import Sequential, Conv2D, BatchNormalization, MaxPooling2D, Flatten, Dense
r = np.random.default_rng(5)
X = r.standard_normal((32, 3, 32, 32))
Y = O.to_categorical(np.concatenate([np.arange(10), r.integers(0, 10, 22)]))
for trial in range(4):
    runs = []
    for fuse in (True, False, True, False):
        np.random.seed(0)
        m = Sequential()
        m.add(Conv2D(64, (3, 3), input_shape=(None, 3, 32, 32), padding='SAME', activation='relu'))
        m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
        m.add(Conv2D(128, (3, 3), padding='SAME', activation='relu'))
        m.add(BatchNormalization()); m.add(MaxPooling2D((2, 2)))
        m.add(Flatten()); m.add(Dense(10, activation='softmax'))
        m.compile('sgd', 'sparse_categorical_crossentropy', fuse=fuse)
        m.fit(X, Y, batch_size=16, epochs=2, shuffle=True, validation_ratio=0., verbose=0)
        runs.append((np.array(m.train_loss), np.asarray(m.layers[0].variables[0].output_tensor)))
    print('fused vs unfused  loss rel', np.max(np.abs(runs[0][0]-runs[1][0])/np.abs(runs[1][0])), 'w', rel_err(runs[0][1], runs[1][1]),
          '| fused vs fused  loss', np.max(np.abs(runs[0][0]-runs[2][0])/np.abs(runs[2][0])), 'w', rel_err(runs[0][1], runs[2][1]), '| unfused vs unfused w', rel_err(runs[1][1], runs[3][1]))
