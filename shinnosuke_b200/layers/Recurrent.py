"""LSTM on the device (reference: layers/Recurrent.py:216-507; SimpleRNN / GRU are not built — the
reference's GRU backward raises, :602).

Same parameters as the reference: W (H + D, 4H) = rows [recurrent; input], gate columns
(f, u, c~, o), b (1, 4H) with the forget-gate bias at one (:225-262).  The input half of the fused
product is one GEMM over all timesteps; the recurrence is ONE launch for the whole sequence when
H is 64 or 128 (thread-block clusters, Wh in registers, h_t exchanged through distributed shared
memory; in the bf16 tier with H = 128 the per-step product runs on the tensor pipe with Wh resident
in shared memory as the UMMA A operand, csrc/recurrent_tc.cu), else one fused launch per step
(csrc/recurrent.cu).  The backward pass is the CORRECT back-propagation through time: the reference
concatenates its gate gradients as (o, c~, u, f) (:358, :406) against forward columns (f, u, c~, o),
so its dW / dX do not match its own forward; the oracle reproduces the
reference's numbers with `reference_gate_order=True` and is checked against numeric gradients
without it, and the device follows the latter."""
import numpy as np

from .Base import Layer, Variable
from ..device import current_stream
from .._lib import lib, PRECISION
from ..utils.Activator import get_activator
from ..utils.Initializers import get_initializer


class Recurrent(Layer):
    """layers/Recurrent.py:11-38."""

    def __init__(self, cell, return_sequences=False, return_state=False, stateful=False, input_length=None, **kwargs):
        super(Recurrent, self).__init__(**kwargs)
        self.cell = cell
        self.return_sequences = return_sequences
        self.return_state = return_state
        self.stateful = stateful
        self.input_length = input_length
        if return_state:
            raise NotImplementedError('return_state=True makes output_tensor a list in the reference (:487); not on the device path')

    def __call__(self, prev_layer):
        assert len(prev_layer.output_shape) == 3, 'Only support batch input'
        return super(Recurrent, self).__call__(prev_layer)


class LSTMCell(object):
    """Parameter factory of layers/Recurrent.py:216-262 (the step itself runs in csrc/recurrent.cu)."""

    def __init__(self, units, activation='tanh', recurrent_activation='sigmoid', initializer='glorotuniform',
                 recurrent_initializer='orthogonal', unit_forget_bias=True):
        self.units = units
        act, ract = get_activator(activation), get_activator(recurrent_activation)
        if act.__class__.__name__ != 'Tanh' or ract.__class__.__name__ != 'Sigmoid':
            raise NotImplementedError('LSTM on the device runs activation=tanh, recurrent_activation=sigmoid')
        self.initializer = get_initializer(initializer)
        self.recurrent_initializer = get_initializer(recurrent_initializer)
        self.unit_forget_bias = unit_forget_bias

    def _initial_params(self, n_in, n_out):
        # draw order of the reference (:229-240): per gate f, u, c~, o: input block, then recurrent block
        cols = []
        for _ in range(4):
            w_l = self.initializer((n_in, n_out))
            w_r = self.recurrent_initializer((n_out, n_out))
            cols.append(np.concatenate((w_r, w_l), axis=0))
        W = Variable(np.concatenate(cols, axis=1), name='lstm_w', layout=('lstm_w', n_out))
        b = np.zeros((1, 4 * n_out))
        if self.unit_forget_bias:
            b[:, :n_out] = 1.0
        return [W, Variable(b, name='lstm_b', layout='flat')]


class LSTM(Recurrent):
    def __init__(self, units, activation='tanh', recurrent_activation='sigmoid', initializer='glorotuniform',
                 recurrent_initializer='orthogonal', unit_forget_bias=True, return_sequences=False, return_state=False,
                 stateful=False, **kwargs):
        cell = LSTMCell(units=units, activation=activation, recurrent_activation=recurrent_activation, initializer=initializer,
                        recurrent_initializer=recurrent_initializer, unit_forget_bias=unit_forget_bias)
        super(LSTM, self).__init__(cell, return_sequences=return_sequences, return_state=return_state, stateful=stateful, **kwargs)
        self._carry = None

    def _shapes(self, in_shape):
        batch, timesteps, n_vec = in_shape
        H = self.cell.units
        # (the reference's connect() declares (batch, 1, H) without return_sequences, :470; its forward
        # produces (batch, H), :481-484, which is what the next layer receives and what is declared here)
        self.output_shape = (batch, timesteps, H) if self.return_sequences else (batch, H)
        self.variables = self.cell._initial_params(n_vec, H)

    def __call__(self, prev_layer):
        self._shapes(prev_layer.output_shape)
        return super(LSTM, self).__call__(prev_layer)

    def connect(self, inbound_layer):
        self._shapes(self.input_shape if inbound_layer is None else inbound_layer.output_shape)
        Layer.connect(self, inbound_layer)

    def forward(self, is_training=True):
        """LSTMCell._forward (:265-318) + LSTM.forward (:476-489)."""
        W, b = self._param_ptrs()
        x = self.input_tensor
        N, T, D = x.shape
        H = self.cell.units
        st = current_stream()
        prec = min(PRECISION[self.precision], 1)
        wh = W.output_tensor.ptr
        wx = wh + 4 * H * H * 4
        # time-major copy of the input: the per-step slices and the three backward GEMMs want rows (t, n)
        xt = self._buf('x_tm', (T, N, D), 'flat')
        lib.sn_swap_leading_axes(x.ptr, xt.ptr, N, T, D, 0, st)
        zx = self._buf('zx', (T, N, 4 * H), 'flat')
        lib.sn_dense_fprop(xt.ptr, wx, b.output_tensor.ptr, zx.ptr, T * N, D, 4 * H, 0, prec, st)
        h_all = self._buf('h_all', (T + 1, N, H), 'flat')
        c_all = self._buf('c_all', (T + 1, N, H), 'flat')
        acts = self._buf('acts', (T, N, 4 * H), 'flat')
        step = N * H * 4
        carry = self.stateful and is_training and self._carry == (N, T)   # :274-276
        if carry:
            lib.sn_memcpy_d2d(h_all.ptr, h_all.ptr + T * step, step, st)
            lib.sn_memcpy_d2d(c_all.ptr, c_all.ptr + T * step, step, st)
        hseq = self._buf('h_seq', (N, T, H), 'flat') if self.return_sequences else None
        self._seq = bool(lib.load().sn_lstm_seq_supported(N, H))
        # bf16 tier, H = 128: the per-step product on the tensor pipe (csrc/recurrent_tc.cu)
        self._seq_tc = bool(self._seq and self.precision == 'bf16' and lib.load().sn_lstm_seq_tc_supported(N, H))
        if self._seq:
            # the whole recurrence in one launch: clusters keep Wh in registers, h_t travels through DSMEM
            (lib.sn_lstm_seq_fwd_tc if self._seq_tc else lib.sn_lstm_seq_fwd)(zx.ptr, wh, h_all.ptr if carry else None, c_all.ptr if carry else None, acts.ptr, c_all.ptr,
                                h_all.ptr, None if hseq is None else hseq.ptr, T, N, H, st)
        for t in range(0 if not self._seq else T, T):
            first = t == 0 and not carry
            lib.sn_lstm_step_fwd(zx.ptr + t * 4 * step, 4 * H, wh,
                                 None if first else h_all.ptr + t * step, None if first else c_all.ptr + t * step,
                                 acts.ptr + t * 4 * step, c_all.ptr + (t + 1) * step, h_all.ptr + (t + 1) * step,
                                 None if hseq is None else hseq.ptr + t * H * 4, T * H, N, H, st)
        self._first_zero = not carry
        if is_training:
            self._carry = (N, T)
        self.output_tensor = hseq if hseq is not None else self._last_view(h_all, T, N, H)
        super(LSTM, self).forward(is_training)

    def _last_view(self, h_all, T, N, H):
        from ..device import DeviceTensor
        return DeviceTensor(h_all.buf[T * N * H:(T + 1) * N * H], (N, H), 'flat')

    def backward(self):
        """LSTMCell._backward (:321-437) with the gate gradients in forward column order, + LSTM.backward."""
        W, b = self.variables
        x = self.input_tensor
        N, T, D = x.shape
        H = self.cell.units
        st = current_stream()
        prec = min(PRECISION[self.precision], 1)
        wh = W.output_tensor.ptr
        wx = wh + 4 * H * H * 4
        g = self.grads
        h_all, c_all = self._buf('h_all', (T + 1, N, H), 'flat'), self._buf('c_all', (T + 1, N, H), 'flat')
        acts = self._buf('acts', (T, N, 4 * H), 'flat')
        xt = self._buf('x_tm', (T, N, D), 'flat')
        dz = self._buf('dz', (T, N, 4 * H), 'flat')
        dc = self._buf('dc', (N, H), 'flat')
        step = N * H * 4
        if self._seq:
            seq_bwd = lib.sn_lstm_seq_bwd_tc if getattr(self, '_seq_tc', False) else lib.sn_lstm_seq_bwd
            if self.return_sequences:
                seq_bwd(None, g.ptr, wh, acts.ptr, c_all.ptr, 1 if self._first_zero else 0, dz.ptr, T, N, H, st)
            else:
                seq_bwd(g.ptr, None, wh, acts.ptr, c_all.ptr, 1 if self._first_zero else 0, dz.ptr, T, N, H, st)
        for t in range(T - 1 if not self._seq else -1, -1, -1):
            last = t == T - 1
            if self.return_sequences:
                dh_ext, dh_stride = g.ptr + t * H * 4, T * H
            else:
                dh_ext, dh_stride = (g.ptr, H) if last else (None, 0)
            c_prev = None if (t == 0 and self._first_zero) else c_all.ptr + t * step
            lib.sn_lstm_step_bwd(dh_ext, dh_stride, None if last else dz.ptr + (t + 1) * 4 * step, wh,
                                 acts.ptr + t * 4 * step, c_all.ptr + (t + 1) * step, c_prev, dc.ptr,
                                 dz.ptr + t * 4 * step, N, H, 1 if last else 0, st)
        # W.grads += [h_{t-1}; x_t]^T . dz_t over all steps (:359-362), as two GEMMs on the two blocks of W
        if W.require_grads or b.require_grads:
            dwh = W.grads.ptr
            dwx = dwh + 4 * H * H * 4
            lib.sn_dense_wgrad(xt.ptr, dz.ptr, dwx, b.grads.ptr if b.require_grads else None, T * N, D, 4 * H, 1, prec, st)
            if self._first_zero:
                if T > 1:       # h_0 = 0 contributes nothing
                    lib.sn_dense_wgrad(h_all.ptr + step, dz.ptr + 4 * step, dwh, None, (T - 1) * N, H, 4 * H, 1, prec, st)
            else:
                lib.sn_dense_wgrad(h_all.ptr, dz.ptr, dwh, None, T * N, H, 4 * H, 1, prec, st)
        targets = [layer for layer in self.inbounds if layer.require_grads]
        if targets:
            dxt = self._buf('dx_tm', (T, N, D), 'flat')
            ws_bytes = lib.sn_conv2d_dgrad_workspace(D, 4 * H, 1, 1, prec)
            ws = self._buf('dgrad_ws', (ws_bytes // 4,), 'flat').ptr if ws_bytes else None
            lib.sn_dense_dgrad(dz.ptr, wx, dxt.ptr, T * N, D, 4 * H, 0, prec, ws, st)
            for layer in targets:
                self._push_grad(layer, lambda dst, acc: lib.sn_swap_leading_axes(dxt.ptr, dst.ptr, T, N, D, acc, st))
