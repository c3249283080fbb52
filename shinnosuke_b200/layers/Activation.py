"""Stand-alone activation layer (reference: layers/Activation.py:7-42)."""
from .Base import Layer
from ..device import current_stream
from .._lib import lib
from ..utils.Activator import get_activator


class Activation(Layer):
    def __init__(self, act_name='relu'):
        self.activator = get_activator(act_name)
        super(Activation, self).__init__()

    def connect(self, prev_layer):
        self.output_shape = prev_layer.output_shape
        Layer.connect(self, prev_layer)

    def __call__(self, prev_layer):
        super(Activation, self).__call__(prev_layer)
        self.output_shape = self.input_shape
        return self

    def _kind(self):
        name = self.activator.__class__.__name__.lower()
        if name not in ('relu', 'linear', 'softmax'):
            raise NotImplementedError('Activation on the device supports relu/linear/softmax (got %s)' % name)
        return name

    def forward(self, is_training=True):
        x = self.input_tensor
        kind = self._kind()
        st = current_stream()
        if kind == 'linear' or getattr(self, '_folded_into', None) is not None:
            # 'relu' folded into the producing BatchNormalization / Add by compile(): that kernel already
            # wrote relu(x) (blocked elements as -0.0f), this layer only hands the buffer on
            self.output_tensor = x
        elif kind == 'relu':
            y = self._buf('out', x.shape, x.layout)
            lib.sn_relu_fwd(x.ptr, y.ptr, x.size, st)
            self.output_tensor = y
        else:
            y = self._buf('out', x.shape, x.layout)
            lib.sn_softmax_fwd(x.ptr, y.ptr, x.size // x.shape[-1], x.shape[-1], st)
            self.output_tensor = y
        super(Activation, self).forward(is_training)

    def backward(self):
        g = self.grads
        st = current_stream()
        if self._kind() == 'relu' and not self._grads_premasked:
            lib.sn_relu_bwd(g.ptr, self.output_tensor.ptr, g.size, st)
        for layer in self.inbounds:
            if self._hand_over_grads(layer, g):
                continue
            self._push_grad(layer, lambda dst, acc: dst.add_(g) if acc else dst.copy_from(g))
