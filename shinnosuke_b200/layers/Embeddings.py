"""Embedding on the device (reference: layers/Embeddings.py:7-76)."""
import numpy as np

from .Base import Layer, Variable
from ..device import current_stream
from .._lib import lib
from ..utils.Initializers import get_initializer


class Embedding(Layer):
    """Rows of W (input_dim, output_dim) gathered by integer ids.  The ids travel through the
    minibatch pipeline as float32 values (exact below 2^24, enforced on input_dim), `mask_zero` is stored
    and unused as in the reference.  Negative ids wrap like NumPy's W[idx]; an id outside
    [-input_dim, input_dim) raises IndexError in the reference — here it is counted on the device and
    `check_ids()` (called by fit() at the end of every epoch) raises the IndexError."""

    @staticmethod
    def check_ids():
        n = lib.embedding_oob_count(reset=True)
        if n:
            raise IndexError('%d Embedding ids were out of bounds for the table (layers/Embeddings.py:60 raises on the first)' % n)

    def __init__(self, input_dim, output_dim, embeddings_initializer='uniform', mask_zero=False, input_length=None, **kwargs):
        super(Embedding, self).__init__(**kwargs)
        if input_dim >= (1 << 24):
            raise ValueError('Embedding ids travel as float32 values: input_dim must be below 2**24 (got %d)' % input_dim)
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.initializer = get_initializer(embeddings_initializer)
        self.mask_zero = mask_zero
        self.input_length = input_length

    def __call__(self, prev_layer):
        assert self.output_dim is not None
        super(Embedding, self).__call__(prev_layer)
        self._initial_params()
        self.output_shape = self.compute_output_shape()
        return self

    def connect(self, inbound_layer):
        assert self.output_dim is not None
        if inbound_layer is None:
            assert self.input_length is not None
            self.input_shape = (None, self.input_length)
        else:
            self.input_shape = inbound_layer.output_shape
        self._initial_params()
        self.output_shape = self.compute_output_shape()
        Layer.connect(self, inbound_layer)

    def _initial_params(self):
        """layers/Embeddings.py:38-43."""
        self.variables.append(Variable(self.initializer((self.input_dim, self.output_dim)), name='embedding_w', layout='flat'))

    def compute_output_shape(self):
        return tuple(self.input_shape) + (self.output_dim,)

    def forward(self, is_training=True):
        """:54-63: output = W[input]."""
        W, = self._param_ptrs()
        x = self.input_tensor
        assert x.ndim == 2
        out = self._buf('out', tuple(x.shape) + (self.output_dim,), 'flat')
        lib.sn_embedding_fwd(W.output_tensor.ptr, x.ptr, out.ptr, x.size, self.input_dim, self.output_dim, current_stream())
        self.output_tensor = out
        super(Embedding, self).forward(is_training)

    def backward(self):
        """:66-76: np.add.at(W.grads, input, grads).  Integer ids carry no gradient."""
        W, = self.variables
        if W.require_grads:
            x = self.input_tensor
            lib.sn_embedding_bwd(x.ptr, self.grads.ptr, W.grads.ptr, x.size, self.input_dim, self.output_dim, current_stream())
