"""Activation tags (reference: utils/Activator.py).  On the device the activations are fused
into the producing kernel's epilogue (ReLU, linear) or run as one small kernel (softmax); these
classes only name the choice, mirroring `get_activator` (:143-155)."""


class Activator(object):
    pass


class Relu(Activator):
    """forward max(0, x); backward zeroes the gradient where the PRE-activation was < 0
    (strict, utils/Activator.py:25).  The device stores negative pre-activations as -0.0f."""


class Linear(Activator):
    pass


class Softmax(Activator):
    """row softmax; backward is the identity (utils/Activator.py:121-124)."""


class Sigmoid(Activator):
    pass


class Tanh(Activator):
    pass


def get_activator(activator):
    if isinstance(activator, str):
        table = {'relu': Relu, 'softmax': Softmax, 'linear': Linear, 'sigmoid': Sigmoid, 'tanh': Tanh}
        key = activator.lower()
        if key in table:
            return table[key]()
        return None
    if isinstance(activator, Activator):
        return activator
    return None
