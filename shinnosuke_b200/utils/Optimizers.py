"""The reference's optimizers on the device (utils/Optimizers.py:8-157): SGD, Momentum, RMSprop,
AdaGrad, AdaDelta, Adam.

When the variables live in a model's parameter arena the update is ONE fused kernel over the
flat arena; stand-alone variables are updated one kernel each.  Bug-compatible details kept:
SGD zeroes the gradients (:32), Momentum never does (:52-55), `decay` is accepted and unused."""
import copy

from ..device import current_stream, zeros
from .._lib import lib


def _to_dev(host_array):
    from ..device import torch
    return torch().from_numpy(host_array).to('cuda')


class Optimizer(object):
    def __init__(self, lr, decay):
        self.lr = lr
        self.decay = decay
        self.iterations = 0
        self.grad_scale = 1.0       # 1/world_size under data parallelism (set by the model)

    ITER_PER_UPDATE = 1             # what one update adds to `iterations` (a graph replay adds the same)

    def update(self, trainable_variables):
        self.iterations += 1

    def hyper_key(self):
        """Everything a captured step graph bakes in as kernel scalars: a change of lr / rho / betas /
        epsilon between two fit() calls must re-capture (the reference reads them on every update)."""
        return (self.__class__.__name__,) + tuple(sorted((k, float(v)) for k, v in self.__dict__.items()
                                                         if isinstance(v, (int, float)) and not isinstance(v, bool)
                                                         and k not in ('iterations', 'grad_scale')))

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
                pending = self.__dict__.pop('_v_host', None)
                if pending is not None and pending.size == arena.param_floats:      # resumed from a checkpoint
                    arena.v[:arena.param_floats].copy_(_to_dev(pending))
            self._arena_ref = arena
            gstep = arena.gstep.data_ptr() if arena.gstep is not None else None
            lib.sn_momentum_update(arena.w.data_ptr(), arena.g.data_ptr(), arena.v.data_ptr(), gstep,
                                   arena.param_floats, float(self.lr), float(self.rho), float(self.grad_scale), st)
        else:
            if self.velocity is None:
                self.velocity = [zeros(v.size) for v in trainable_variables]
                pending = self.__dict__.pop('_vel_host', None)
                if pending is not None and len(pending) == len(self.velocity):
                    for dst, src in zip(self.velocity, pending):
                        dst.copy_(_to_dev(src))
            for vel, var in zip(self.velocity, trainable_variables):
                var.ensure_device()
                lib.sn_momentum_update(var.output_tensor.ptr, var.grads.ptr, vel.data_ptr(), None, var.size,
                                       float(self.lr), float(self.rho), 1.0, st)
        super(Momentum, self).update(trainable_variables)

    def __getstate__(self):
        """The reference pickles its velocity arrays with the optimizer (models.py:208-220 dumps
        [layers, optimizer, loss]); here they come off the device as NumPy arrays."""
        d = dict(self.__dict__)
        arena = d.pop('_arena_ref', None)
        if arena is not None and arena.v is not None:
            d['_v_host'] = arena.v[:arena.param_floats].cpu().numpy()
        if self.velocity is not None:
            d['_vel_host'] = [v.cpu().numpy() for v in self.velocity]
        d['velocity'] = None
        return d


class _Adaptive(Optimizer):
    """RMSprop / AdaGrad / AdaDelta / Adam share one kernel family (csrc/optimizers.cu).  Like
    Momentum — and like the reference — they never zero the gradients."""
    KIND = None
    TWO_STATES = False

    def _params(self):
        raise NotImplementedError

    def _state(self, owner, n):
        st = self.__dict__.setdefault('_states', {})
        key = (id(owner), n)
        if key not in st:
            st[key] = (zeros(n), zeros(n) if self.TWO_STATES else None, zeros(4) if self.KIND == 3 else None, owner)
            pending = self.__dict__.get('_states_host')
            if pending:                  # resumed from a checkpoint: states come back in creation order
                for i, (pn, h1, h2, ht) in enumerate(pending):
                    if pn == n:
                        s1, s2, tick = st[key][:3]
                        s1[:h1.size].copy_(_to_dev(h1))
                        if s2 is not None and h2 is not None:
                            s2[:h2.size].copy_(_to_dev(h2))
                        if tick is not None and ht is not None:
                            tick.copy_(_to_dev(ht))
                        del pending[i]
                        break
        return st[key][:3]

    def __getstate__(self):
        """The reference pickles ms / vs with the optimizer; the device state (and Adam's device tick,
        which is what the bias correction reads) travels as NumPy arrays."""
        d = dict(self.__dict__)
        host = []
        for (_, n), (s1, s2, tick, _owner) in self.__dict__.get('_states', {}).items():
            host.append((n, s1.cpu().numpy(), s2.cpu().numpy() if s2 is not None else None,
                         tick.cpu().numpy() if tick is not None else None))
        d['_states_host'] = host + list(self.__dict__.get('_states_host') or [])
        d['_states'] = {}
        return d

    def update(self, trainable_variables):
        st = current_stream()
        lr, p1, p2, eps = self._params()
        arena = self._arena(trainable_variables)
        if arena is not None:
            s1, s2, tick = self._state(arena, arena.total)
            gstep = arena.gstep.data_ptr() if arena.gstep is not None else None
            lib.sn_adaptive_update(self.KIND, arena.w.data_ptr(), arena.g.data_ptr(), s1.data_ptr(),
                                   s2.data_ptr() if s2 is not None else None, gstep,
                                   tick.data_ptr() if tick is not None else None, arena.param_floats,
                                   float(lr), float(p1), float(p2), float(eps), float(self.grad_scale), st)
        else:
            for var in trainable_variables:
                var.ensure_device()
                s1, s2, tick = self._state(var, var.size)
                lib.sn_adaptive_update(self.KIND, var.output_tensor.ptr, var.grads.ptr, s1.data_ptr(),
                                       s2.data_ptr() if s2 is not None else None, None,
                                       tick.data_ptr() if tick is not None else None, var.size,
                                       float(lr), float(p1), float(p2), float(eps), 1.0, st)
        super(_Adaptive, self).update(trainable_variables)

    needs_gstep = True          # data parallel: gradients persist, so each step's go through arena.gstep


class RMSprop(_Adaptive):
    KIND = 0

    def __init__(self, lr=0.001, decay=0.0, rho=0.9, epsilon=1e-7, *args, **kwargs):
        self.rho = rho
        self.epsilon = epsilon
        super(RMSprop, self).__init__(lr, decay)

    def _params(self):
        return self.lr, self.rho, 0.0, self.epsilon


class AdaGrad(_Adaptive):
    KIND = 1

    def __init__(self, lr=0.01, decay=0.0, epsilon=1e-7):
        super(AdaGrad, self).__init__(lr, decay)
        self.epsilon = epsilon

    def _params(self):
        return self.lr, 0.0, 0.0, self.epsilon


class AdaDelta(_Adaptive):
    KIND = 2
    TWO_STATES = True

    def __init__(self, decay=0.0, lr=1.0, rho=0.95, epsilon=1e-7):
        super(AdaDelta, self).__init__(lr, decay)
        self.rho = rho
        self.epsilon = epsilon

    def _params(self):
        return self.lr, self.rho, 0.0, self.epsilon          # lr is accepted and unused (:110-123)


class Adam(_Adaptive):
    KIND = 3
    TWO_STATES = True
    ITER_PER_UPDATE = 2

    def __init__(self, lr=0.001, decay=0.0, beta1=0.9, beta2=0.999, epsilon=1e-7, *args, **kwargs):
        self.beta1 = beta1
        self.beta2 = beta2
        self.epsilon = epsilon
        super(Adam, self).__init__(lr, decay)

    def _params(self):
        return self.lr, self.beta1, self.beta2, self.epsilon

    def update(self, trainable_variables):
        self.iterations += 1            # utils/Optimizers.py:139: two increments per update (the device tick mirrors it)
        super(Adam, self).update(trainable_variables)


def get_optimizer(optimizer):
    if isinstance(optimizer, str):
        key = optimizer.lower()
        if key == 'sgd':
            return StochasticGradientDescent()
        if key == 'momentum':
            return Momentum()
        if key == 'adam':
            return Adam()
        if key == 'rmsprop':
            return RMSprop()
        if key == 'adagrad':
            return AdaGrad()
        if key == 'adadelta':
            return AdaDelta()
        raise ValueError('unknown optimizer type!')
    if isinstance(optimizer, Optimizer):
        return copy.deepcopy(optimizer)
    raise ValueError('unknown optimizer type!')
