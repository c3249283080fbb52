"""Mini-batch order (reference: utils/MiniBatch.py:6-34)."""
import numpy as np


def batch_permutation(m, seed, shuffle):
    """The reference reseeds the GLOBAL NumPy RNG with the epoch index and draws one
    permutation; reproduced so both frameworks visit the samples in the same order."""
    np.random.seed(seed)
    return np.random.permutation(m) if shuffle else None


def batch_slices(m, batch_size):
    """Complete batches followed by the ragged tail, if any (:23-32)."""
    return [(s, min(s + batch_size, m)) for s in range(0, m, batch_size)]


def get_batches(X, y, batch_size, seed, shuffle):
    perm = batch_permutation(X.shape[0], seed, shuffle)
    if perm is not None:
        X, y = X[perm], y[perm]
    return [(X[a:b], y[a:b]) for a, b in batch_slices(X.shape[0], batch_size)]
