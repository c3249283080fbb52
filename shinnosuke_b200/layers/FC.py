"""Flatten / Dense on the device (reference: layers/FC.py, Flatten :10-58, Dense :64-153)."""
import numpy as np

from .Base import Layer, Variable
from ..device import current_stream
from .._lib import lib, ACT, PRECISION
from ..utils.Initializers import get_initializer
from ..utils.Activator import get_activator
from ..utils.Regularizers import get_regularizer


class Flatten(Layer):
    def __init__(self, out_dim=2):
        if out_dim < 1:
            raise ValueError('out_dim must be > 0')
        if out_dim != 2:
            raise NotImplementedError('the device path flattens to (N, features) only (out_dim=2)')
        self.out_dim = out_dim
        super(Flatten, self).__init__()

    def _infer(self, input_shape):
        flatten_shape = np.prod(np.array(input_shape[self.out_dim - 1:])).tolist()
        self.output_shape = tuple(input_shape[:self.out_dim - 1]) + (flatten_shape,)

    def connect(self, prev_layer):
        assert len(prev_layer.output_shape) >= 3
        self._infer(prev_layer.output_shape)
        Layer.connect(self, prev_layer)

    def __call__(self, layers):
        super(Flatten, self).__call__(layers)
        self._infer(self.input_shape)
        return self

    def forward(self, is_training=True):
        """layers/FC.py:39-48.  The reference flattens NCHW in (c,i,j) order, which fixes the row
        order of the next Dense's weights; the device map is NHWC, so this is a real transpose."""
        x = self.input_tensor
        self.input_shape = x.shape
        if x.ndim == 4 and getattr(self, '_hwc_passthrough', False):
            # the only consumer is a Dense whose weight rows are stored in (i,j,c) order (set up by
            # compile()): the NHWC map IS its input, no transpose kernel either way
            N, C, H, W = x.shape
            y = x.view((N, C * H * W), ('flat_hwc', C, H, W))
        elif x.ndim == 4:
            N, C, H, W = x.shape
            y = self._buf('out', (N, C * H * W), 'flat')
            lib.sn_nhwc_to_nchw(x.ptr, y.ptr, N, C, H, W, 0, current_stream())
        else:
            y = x.view((x.shape[0], int(np.prod(x.shape[1:]))))
        self.output_tensor = y
        super(Flatten, self).forward(is_training)
        self._grads_alias = None
        if is_training and self.require_grads and isinstance(y.layout, tuple) and len(self.inbounds) == 1:
            prev = self.inbounds[0]
            if prev.require_grads and len(prev.outbound_layers) == 1 and prev.grads is not None:
                # our gradient buffer IS the producer's: the Dense behind us writes dX straight into it
                self.grads = prev.grads.view(y.shape, y.layout)
                self._grads_alias = prev

    def backward(self):
        """layers/FC.py:53-58."""
        g = self.grads
        st = current_stream()
        if getattr(self, '_grads_alias', None) is not None:
            self._grads_alias._grads_written = True      # nothing to move
            return
        for layer in self.inbounds:
            if len(self.input_shape) == 4 and isinstance(g.layout, tuple) and g.layout[0] == 'flat_hwc':
                def writer(dst, acc):                   # the gradient already is the NHWC map
                    dst.add_(g) if acc else dst.copy_from(g)
            elif len(self.input_shape) == 4:
                N, C, H, W = self.input_shape

                def writer(dst, acc):
                    if acc:
                        tmp = self._buf('bwd_tmp', self.input_shape, 'nhwc')
                        lib.sn_nchw_to_nhwc(g.ptr, tmp.ptr, N, C, H, W, st)
                        dst.add_(tmp)
                    else:
                        lib.sn_nchw_to_nhwc(g.ptr, dst.ptr, N, C, H, W, st)
            else:
                def writer(dst, acc):
                    dst.add_(g) if acc else dst.copy_from(g)
            self._push_grad(layer, writer)


class Dense(Layer):
    def __init__(self, n_out, n_in=None, initializer='Normal', activation='linear', kernel_regularizer=None, **kwargs):
        self.n_out = n_out
        self.n_in = n_in
        self.initializer = get_initializer(initializer)
        self.activator = get_activator(activation)
        self.kernel_regularizer = get_regularizer(kernel_regularizer)   # stored, never applied (as in the reference)
        super(Dense, self).__init__(**kwargs)

    def connect(self, prev_layer):
        if prev_layer is None:
            if self.n_in is None and self.input_shape is None:
                raise ValueError('must specify n_in or input_shape')
            elif self.input_shape is None:
                self.input_shape = (None, self.n_in)
        else:
            self.input_shape = prev_layer.output_shape
        self._initial_params()
        Layer.connect(self, prev_layer)
        self.output_shape = self.compute_output_shape()

    def __call__(self, prev_layer):
        super(Dense, self).__call__(prev_layer)
        self._initial_params()
        self.output_shape = self.compute_output_shape()
        return self

    def _initial_params(self):
        """layers/FC.py:105-113: W (n_in, n_out) from the initializer, b zeros (1, n_out).  On the
        device W is stored [n_out][n_in] so Dense is the 1x1 case of the convolution kernels."""
        n_in = self.input_shape[-1]
        W = Variable(self.initializer((n_in, self.n_out)), name='dense_w', layout='dense_w')
        b = Variable(np.zeros((1, self.n_out)), name='dense_b', layout='flat')
        self.variables.append(W)
        self.variables.append(b)

    def compute_output_shape(self):
        return tuple(self.input_shape[:-1]) + (self.n_out,)

    def _act_name(self):
        name = self.activator.__class__.__name__.lower()
        if name not in ('linear', 'relu', 'softmax'):
            raise NotImplementedError('Dense on the device supports linear/relu/softmax (got %s)' % name)
        return name

    def _use_hwc_rows(self, c, h, w):
        """compile() hook: the Flatten in front of this layer passes its NHWC map through, so the
        weight rows (and their gradients) are stored in (i,j,c) order on the device."""
        Wv = self.variables[0]
        assert Wv.output_shape[0] == c * h * w
        Wv.relayout(('dense_w_hwc', c, h, w))

    def _bf16_ok(self, op, M, K):
        key = (op, M, K)
        cache = self.__dict__.setdefault('_bf16_support', {})
        if key not in cache:
            cache[key] = bool(lib.load().sn_conv2d_bf16_supported(op, M, K, 1, 1, self.n_out, 1, 1, 1, 0, 0))
        return cache[key]

    def _bf16_buf(self, name, n):
        from ..device import empty, torch
        key = (name, int(n))
        t = self._bufs.get(key)
        if t is None:
            t = empty(int(n), torch().bfloat16)
            self._bufs[key] = t
        return t

    def _head(self, M, K):
        return self._act_name() in ('linear', 'softmax') and bool(lib.load().sn_dense_head_supported(M, K, self.n_out))

    def forward(self, is_training=True):
        """layers/FC.py:124-134."""
        W, b = self._param_ptrs()
        x = self.input_tensor
        if x.ndim != 2:
            raise NotImplementedError('Dense on the device takes (N, features) inputs')
        M, K = x.shape
        hwc_w = isinstance(W.output_tensor.layout, tuple)
        hwc_x = isinstance(x.layout, tuple)
        if hwc_w != hwc_x or (hwc_w and tuple(W.output_tensor.layout[1:]) != tuple(x.layout[1:])):
            raise RuntimeError('Dense: input feature order %r does not match the weight row order %r'
                               % (x.layout, W.output_tensor.layout))
        act = self._act_name()
        st = current_stream()
        if self._head(M, K):
            # thin classifier head (n_out <= 16): one launch for x.W + b (+ softmax), csrc/dense_head.cu
            out = self._buf('probs' if act == 'softmax' else 'out', (M, self.n_out), 'flat')
            lib.sn_dense_head_fprop(x.ptr, W.output_tensor.ptr, b.output_tensor.ptr, out.ptr, M, K, self.n_out,
                                    1 if act == 'softmax' else 0, st)
            self.output_tensor = out
            super(Dense, self).forward(is_training)
            return
        z = self._buf('out', (M, self.n_out), 'flat')
        act_code = ACT['relu'] if act == 'relu' else ACT['linear']
        self._xb = None
        if self.precision == 'bf16' and self._bf16_ok(0, M, K):
            # bf16 tier: Dense is the 1x1 convolution on bf16 operand copies (tcgen05 kind::f16, fp32 accumulation);
            # K need not be a multiple of the 64-element TMA box (the tail is zero-filled)
            self._xb = self._bf16_buf('x_bf16', x.size)
            wb = self._bf16_buf('w_bf16', W.size)
            lib.sn_cast_f32_to_bf16(x.ptr, self._xb.data_ptr(), x.size, st)
            lib.sn_cast_f32_to_bf16(W.output_tensor.ptr, wb.data_ptr(), W.size, st)
            lib.sn_conv2d_fprop_bf16(self._xb.data_ptr(), wb.data_ptr(), b.output_tensor.ptr, z.ptr, M, K, 1, 1, self.n_out, 1, 1, 1, 0, 0,
                                     act_code, st)
        else:
            lib.sn_dense_fprop(x.ptr, W.output_tensor.ptr, b.output_tensor.ptr, z.ptr, M, K, self.n_out, act_code,
                               PRECISION[self.precision], st)
        if act == 'softmax':
            p = self._buf('probs', (M, self.n_out), 'flat')
            lib.sn_softmax_fwd(z.ptr, p.ptr, M, self.n_out, st)
            self.output_tensor = p
        else:
            self.output_tensor = z
        super(Dense, self).forward(is_training)

    def backward(self):
        """layers/FC.py:140-153 (Softmax backward is the identity, utils/Activator.py:121-124)."""
        W, b = self.variables
        x, g = self.input_tensor, self.grads
        M, K = x.shape
        st = current_stream()
        prec = PRECISION[self.precision]
        if self._act_name() == 'relu' and not self._grads_premasked:
            lib.sn_relu_bwd(g.ptr, self.output_tensor.ptr, g.size, st)
        if self._head(M, K):
            # dW += x^T.dZ, db += sum dZ and dX in ONE launch that reads x once
            dw = W.grads.ptr if W.require_grads else None
            db = b.grads.ptr if b.require_grads else None
            targets = [layer for layer in self.inbounds if layer.require_grads]
            if not targets:
                if dw is not None or db is not None:
                    lib.sn_dense_head_bwd(x.ptr, g.ptr, W.output_tensor.ptr, dw, db, None, M, K, self.n_out, 0, st)
                return
            for i, layer in enumerate(targets):
                first = i == 0

                def writer(dst, acc, first=first):
                    lib.sn_dense_head_bwd(x.ptr, g.ptr, W.output_tensor.ptr, dw if first else None, db if first else None,
                                          dst.ptr, M, K, self.n_out, acc, st)
                self._push_grad(layer, writer)
            return
        if W.require_grads or b.require_grads:
            lib.sn_dense_wgrad(x.ptr, g.ptr, W.grads.ptr, b.grads.ptr if b.require_grads else None,
                               M, K, self.n_out, 1, prec, st)
        if not any(layer.require_grads for layer in self.inbounds):
            return
        if self.precision == 'bf16' and self._bf16_ok(1, M, K):
            gb = self._bf16_buf('dy_bf16', g.size)
            wtb = self._bf16_buf('wt_bf16', W.size)
            lib.sn_cast_f32_to_bf16(g.ptr, gb.data_ptr(), g.size, st)
            lib.sn_flip_transpose_filters_bf16(W.output_tensor.ptr, wtb.data_ptr(), self.n_out, K, 1, 1, st)
            for layer in self.inbounds:
                self._push_grad(layer, lambda dst, acc: lib.sn_conv2d_dgrad_bf16(
                    gb.data_ptr(), wtb.data_ptr(), dst.ptr, M, K, 1, 1, self.n_out, 1, 1, 1, 0, 0, acc, st))
            return
        ws_bytes = lib.sn_conv2d_dgrad_workspace(K, self.n_out, 1, 1, prec)
        ws = self._buf('dgrad_ws', (ws_bytes // 4,), 'flat').ptr if ws_bytes else None
        for layer in self.inbounds:
            self._push_grad(layer, lambda dst, acc: lib.sn_dense_dgrad(
                g.ptr, W.output_tensor.ptr, dst.ptr, M, K, self.n_out, acc, prec, ws, st))
