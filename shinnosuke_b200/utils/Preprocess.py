"""Host-side helpers (reference: utils/Preprocess.py)."""
import numpy as np


def to_categorical(inputs):
    """utils/Preprocess.py:5-20: integer labels -> one-hot rows."""
    inputs = np.asarray(inputs)
    if inputs.ndim > 2:
        raise ValueError('only accept 1-d or 2-d inputs')
    if inputs.ndim == 2:
        if inputs.shape[-1] != 1:
            raise ValueError('can not convert %s to one-hot vector' % (inputs.__class__))
        inputs = inputs[:, 0]
    n_class = int(np.max(inputs)) + 1
    return np.eye(n_class)[inputs].reshape(-1, n_class)
