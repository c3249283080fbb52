"""SGD / Momentum on the device (reference: utils/Optimizers.py:8-57).

When the variables live in a model's parameter arena the update is ONE fused kernel over the
flat arena; stand-alone variables are updated one kernel each.  Bug-compatible details kept:
SGD zeroes the gradients (:32), Momentum never does (:52-55), `decay` is accepted and unused."""
import copy

from ..device import current_stream, zeros
from .._lib import lib


class Optimizer(object):
    def __init__(self, lr, decay):
        self.lr = lr
        self.decay = decay
        self.iterations = 0
        self.grad_scale = 1.0       # 1/world_size under data parallelism (set by the model)

    def update(self, trainable_variables):
        self.iterations += 1

    @staticmethod
    def _arena(trainable_variables):
        return getattr(trainable_variables, 'arena', None)


class StochasticGradientDescent(Optimizer):
    def __init__(self, lr=0.01, decay=0.0, *args, **kwargs):
        super(StochasticGradientDescent, self).__init__(lr, decay)

    def update(self, trainable_variables):
        st = current_stream()
        arena = self._arena(trainable_variables)
        if arena is not None:
            lib.sn_sgd_update(arena.w.data_ptr(), arena.g.data_ptr(), arena.param_floats, float(self.lr),
                              float(self.grad_scale), st)
        else:
            for var in trainable_variables:
                var.ensure_device()
                lib.sn_sgd_update(var.output_tensor.ptr, var.grads.ptr, var.size, float(self.lr),
                                  float(self.grad_scale), st)
        super(StochasticGradientDescent, self).update(trainable_variables)


class Momentum(Optimizer):
    def __init__(self, lr=0.01, decay=0.0, rho=0.9, *args, **kwargs):
        self.rho = rho
        self.velocity = None
        super(Momentum, self).__init__(lr, decay)

    def update(self, trainable_variables):
        st = current_stream()
        arena = self._arena(trainable_variables)
        if arena is not None:
            if arena.v is None:
                arena.v = zeros(arena.total)
            gstep = arena.gstep.data_ptr() if arena.gstep is not None else None
            lib.sn_momentum_update(arena.w.data_ptr(), arena.g.data_ptr(), arena.v.data_ptr(), gstep,
                                   arena.param_floats, float(self.lr), float(self.rho), float(self.grad_scale), st)
        else:
            if self.velocity is None:
                self.velocity = [zeros(v.size) for v in trainable_variables]
            for vel, var in zip(self.velocity, trainable_variables):
                var.ensure_device()
                lib.sn_momentum_update(var.output_tensor.ptr, var.grads.ptr, vel.data_ptr(), None, var.size,
                                       float(self.lr), float(self.rho), 1.0, st)
        super(Momentum, self).update(trainable_variables)

    def __getstate__(self):
        d = dict(self.__dict__)
        d['velocity'] = None
        return d


def get_optimizer(optimizer):
    if isinstance(optimizer, str):
        key = optimizer.lower()
        if key == 'sgd':
            return StochasticGradientDescent()
        if key == 'momentum':
            return Momentum()
        if key in ('adam', 'rmsprop', 'adagrad', 'adadelta'):
            raise NotImplementedError('optimizer %r is off the accelerated hot path (SURVEY.md section 2 row 6)' % optimizer)
        raise ValueError('unknown optimizer type!')
    if isinstance(optimizer, Optimizer):
        return copy.deepcopy(optimizer)
    raise ValueError('unknown optimizer type!')
