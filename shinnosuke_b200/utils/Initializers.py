"""Weight initialisers (reference: utils/Initializers.py).  Host side, run once; they draw from
the legacy global NumPy stream with the same calls as the reference so that, after
``np.random.seed(s)``, a model built here starts from bit-identical weights."""
import copy

import numpy as np


def _fans(size):
    """utils/Initializers.py:12-22, bug-compatible: the reference tests `np.array(size).ndim`, which is
    1 for every shape TUPLE, so its 2-D / 4-D branches never run and every scaled initializer uses
    fan_in = fan_out = int(sqrt(prod(size))).  Identical initial weights need the same arithmetic."""
    size = np.array(size)
    if size.ndim == 2:
        return size[0], size[1]
    if size.ndim in (4, 5):
        field = np.prod(size[2:])
        return size[1] * field, size[0] * field
    n = int(np.sqrt(np.prod(size)))
    return n, n


class Initializer(object):
    def __init__(self, seed=None):
        if seed is not None:
            np.random.seed(seed)

    def decompose_size(self, size):
        return _fans(size)


class Zeros(Initializer):
    def __call__(self, size):
        return np.zeros(size)


class Ones(Initializer):
    def __call__(self, size):
        return np.ones(size)


class Uniform(Initializer):
    def __init__(self, scale=0.05, seed=None):
        self.scale = scale
        super(Uniform, self).__init__(seed)

    def __call__(self, size):
        return np.random.uniform(-self.scale, self.scale, size=size)


class Normal(Initializer):
    def __init__(self, std=0.1, mean=0.0, seed=None):
        self.std, self.mean = std, mean
        super(Normal, self).__init__(seed)

    def __call__(self, size):
        return np.random.normal(loc=self.mean, scale=self.std, size=size)


class _Scaled(Initializer):
    """utils/Initializers.py:39-121: a Uniform / Normal whose scale comes from the fans.  (Module-level
    classes: instances live on layers, and layers are pickled by save().)"""
    base = None

    @staticmethod
    def scale(fan_in, fan_out):
        raise NotImplementedError

    def __call__(self, size):
        fan_in, fan_out = _fans(size)
        return self.base(self.scale(fan_in, fan_out))(size)


class LecunUniform(_Scaled):
    base = Uniform
    scale = staticmethod(lambda i, o: np.sqrt(3. / i))


class GlorotUniform(_Scaled):
    base = Uniform
    scale = staticmethod(lambda i, o: np.sqrt(6. / (i + o)))


class HeUniform(_Scaled):
    base = Uniform
    scale = staticmethod(lambda i, o: np.sqrt(6. / i))


class LecunNormal(_Scaled):
    base = Normal
    scale = staticmethod(lambda i, o: np.sqrt(1. / i))


class GlorotNormal(_Scaled):
    base = Normal
    scale = staticmethod(lambda i, o: np.sqrt(2. / (i + o)))


class HeNormal(_Scaled):
    base = Normal
    scale = staticmethod(lambda i, o: np.sqrt(2. / i))


class Orthogonal(Initializer):
    def __init__(self, gain=1.0, seed=None):
        self.gain = np.sqrt(2) if gain == 'relu' else gain
        super(Orthogonal, self).__init__(seed)

    def __call__(self, size):
        size = [int(s) for s in np.array(size).tolist()]
        flat = (size[0], int(np.prod(size[1:])))
        a = Normal(1.)(flat)
        u, _, v = np.linalg.svd(a, full_matrices=False)
        q = u if u.shape == flat else v
        return self.gain * q.reshape(size)


_BY_NAME = {
    'zeros': Zeros, 'ones': Ones, 'uniform': Uniform, 'normal': Normal,
    'heuniform': HeUniform, 'he_uniform': HeUniform, 'henormal': HeNormal, 'he_normal': HeNormal,
    'lecunnormal': LecunNormal, 'lecun_normal': LecunNormal, 'lecununiform': LecunUniform,
    'lecun_uniform': LecunUniform, 'orthogonal': Orthogonal,
    'glorotnoraml': GlorotNormal, 'glorot_normal': GlorotNormal,      # first spelling as in the reference (:180)
    'glorotuniform': GlorotUniform, 'glorot_uniform': GlorotUniform,
}


def get_initializer(initializer):
    if isinstance(initializer, str):
        cls = _BY_NAME.get(initializer.lower())
        return cls() if cls is not None else None
    if isinstance(initializer, Initializer):
        return copy.deepcopy(initializer)
    raise ValueError('unknown initializer type!')
