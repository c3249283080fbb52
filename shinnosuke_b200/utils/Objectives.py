"""Losses on the device (reference: utils/Objectives.py).  Only the loss the hot path uses is
implemented: SparseCategoricalCrossEntropy (:72-91), which — despite its name — takes ONE-HOT
labels.  backward, calc_loss and calc_acc are one kernel (sn_scce_bwd)."""
import copy

import numpy as np

from ..device import DeviceTensor, as_device, current_stream, zeros
from .._lib import lib


class Objective(object):
    pass


class SparseCategoricalCrossEntropy(Objective):
    def __init__(self):
        self._dz = None
        self._metrics = None
        self._rows = 0

    def _run(self, y_hat, y_true, metrics_ptr=None, denom_rows=0):
        p = as_device(y_hat)
        y = as_device(y_true)
        rows, classes = p.size // p.shape[-1], p.shape[-1]
        # one gradient buffer per logits shape, kept for the life of the loss: captured step graphs hold
        # its address, and a ragged last minibatch alternates with the full ones
        bufs = self.__dict__.setdefault('_dz_by_shape', {})
        if p.shape not in bufs:
            bufs[p.shape] = DeviceTensor.empty(p.shape, 'flat')
        self._dz = bufs[p.shape]
        st = current_stream()
        if metrics_ptr is None:
            if self._metrics is None:
                self._metrics = zeros(2)
            lib.sn_memset(self._metrics.data_ptr(), 0, 8, st)
            metrics_ptr = self._metrics.data_ptr()
        lib.sn_scce_bwd(p.ptr, y.ptr, self._dz.ptr, metrics_ptr, rows, classes, int(denom_rows), st)
        self._rows = rows
        return self._dz

    def backward(self, y_hat, y_true, metrics_ptr=None, denom_rows=0):
        """(y_hat - y)/N with N = prod(shape[:-1]) (:89-91); also leaves sum(-y log y_hat) and the
        number of correct rows on the device for calc_loss / calc_acc."""
        return self._run(y_hat, y_true, metrics_ptr, denom_rows)

    def _host_metrics(self):
        m = np.empty(2, dtype=np.float32)
        lib.sn_memcpy_d2h(m.ctypes.data, self._metrics.data_ptr(), 8, current_stream())
        lib.sn_stream_sync(current_stream())
        return m

    def calc_loss(self, y_hat, y_true):
        self._run(y_hat, y_true)
        return float(self._host_metrics()[0]) / self._rows

    def calc_acc(self, y_hat, y_true):
        self._run(y_hat, y_true)
        return float(self._host_metrics()[1]) / self._rows

    def calc_loss_acc(self, y_hat, y_true):
        self._run(y_hat, y_true)
        m = self._host_metrics()
        return float(m[0]) / self._rows, float(m[1]) / self._rows

    def __getstate__(self):
        return {'_dz': None, '_metrics': None, '_rows': 0}


def get_objective(objective):
    if isinstance(objective, str):
        key = objective.lower()
        if key in ('sparsecategoricalcrossentropy', 'sparse_categorical_crossentropy',
                   'sparse_categorical_cross_entropy'):
            return SparseCategoricalCrossEntropy()
        raise NotImplementedError('objective %r is off the accelerated hot path (SURVEY.md section 2 row 5)' % objective)
    if isinstance(objective, Objective):
        return copy.deepcopy(objective)
    raise ValueError('unknown objective type!')
