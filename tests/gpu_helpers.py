"""Helpers shared by the -m gpu tests: drive ONE device layer the way the golden script drives
the reference layer (Input -> Activation('linear') -> layer), through the public Layer API."""
import numpy as np

from shinnosuke_b200 import Input, Activation
from shinnosuke_b200.device import as_device


def drive(layer, x, upstream_grad, is_training=True, precision='fp32'):
    """Returns (y, dx) as float64 arrays in the reference's layout; parameter grads stay on the
    layer's variables."""
    inp = Input((None,) + tuple(x.shape[1:]))
    src = Activation('linear')(inp)
    layer(src)
    layer.precision = precision
    inp.input_tensor = x
    for l in (inp, src, layer):
        l.forward(is_training)
    y = np.asarray(layer.output_tensor)
    dx = None
    if upstream_grad is not None:
        layer._seed_grads(as_device(upstream_grad))
        layer.backward()
        dx = np.asarray(src.grads)
    return y, dx


def set_params(layer, *values):
    for v, val in zip(layer.variables, values):
        v.output_tensor = np.asarray(val, dtype=np.float64).reshape(v.output_shape)
