"""Conv2D / ZeroPadding2D / MaxPooling2D on the device.

Operator interface of the reference's layers/Convolution.py (Conv2D :8-205, ZeroPadding2D
:209-251, MaxPooling2D :254-362); the arithmetic is done by the sm_100a kernels behind
sn_conv2d_* / sn_maxpool2d_* (include/shinnosuke_b200.h).  Maps are NHWC on the device and the
filters [F][kh][kw][C]; np.asarray() of any tensor gives the reference's NCHW view.
"""
import numpy as np

from .Base import Layer, Variable, defer
from ..device import DeviceTensor, current_stream
from .._lib import lib, ACT, PRECISION, POOL_MODE
from ..utils.Initializers import get_initializer
from ..utils.Activator import get_activator


class Conv2D(Layer):
    def __init__(self, filter_nums, filter_size, input_shape=None, stride=1, padding='VALID', activation='linear',
                 initializer='Normal'):
        self.filter_nums = filter_nums
        self.filter_size = filter_size
        self.stride = stride
        self.padding = padding
        self.activator = get_activator(activation)
        self.initializer = get_initializer(initializer)
        super(Conv2D, self).__init__(input_shape=input_shape)

    # shape inference + parameter creation: layers/Convolution.py:20-57 / :59-90
    def _infer(self, input_shape):
        assert len(input_shape) == 4
        batch_nums, n_C_prev, n_H_prev, n_W_prev = input_shape
        filter_h, filter_w = self.filter_size
        assert isinstance(self.padding, str)
        self.padding = self.padding.upper()
        s = self.stride
        if self.padding == 'SAME':
            pad_h = (s * (n_H_prev - 1) - n_H_prev + filter_h) // 2
            pad_w = (s * (n_W_prev - 1) - n_W_prev + filter_w) // 2
            n_H, n_W = n_H_prev, n_W_prev
            true_h = (n_H_prev + 2 * pad_h - filter_h) // s + 1
            true_w = (n_W_prev + 2 * pad_w - filter_w) // s + 1
            if (true_h, true_w) != (n_H, n_W):
                # the reference declares oh = H for SAME but its pad formula only delivers it for
                # odd filters at stride 1; anything else has no well-defined result there either
                raise ValueError('padding=SAME needs an odd filter and stride 1 (got %r, stride %d)'
                                 % (self.filter_size, s))
            self.pad_size = (pad_h, pad_w)
        elif self.padding == 'VALID':
            n_H = (n_H_prev - filter_h) // s + 1
            n_W = (n_W_prev - filter_w) // s + 1
            self.pad_size = (0, 0)
        else:
            raise TypeError('Unknown padding type!plz inputs SAME or VALID')
        self.output_shape = (batch_nums, self.filter_nums, n_H, n_W)
        # RNG order of the reference: initializer(W.shape) then randn(1, F) (:52-53)
        W = Variable(self.initializer((self.filter_nums, n_C_prev, filter_h, filter_w)), name='conv_w', layout='krsc')
        b = Variable(np.random.randn(1, self.filter_nums), name='conv_b', layout='flat')
        self.variables.append(W)
        self.variables.append(b)

    def connect(self, prev_layer):
        if prev_layer is None:
            assert self.input_shape is not None
            input_shape = self.input_shape
        else:
            input_shape = prev_layer.output_shape
            Layer.connect(self, prev_layer)
        self._infer(input_shape)

    def __call__(self, prev_layer):
        super(Conv2D, self).__call__(prev_layer)
        self._infer(self.input_shape)
        return self

    def _act_code(self):
        name = self.activator.__class__.__name__.lower()
        if name not in ACT:
            raise NotImplementedError('Conv2D on the device fuses linear/relu epilogues only (got %s)' % name)
        return ACT[name]

    def forward(self, is_training=True):
        """layers/Convolution.py:131-161 with true zero padding (SAME policy)."""
        W, b = self._param_ptrs()
        x = self.input_tensor
        N, C, H, Wd = x.shape
        self.input_shape = x.shape
        F = self.filter_nums
        kh, kw = self.filter_size
        oh, ow = self.output_shape[2:]
        y = self._buf('out', (N, F, oh, ow), 'nhwc')
        geom = (N, C, H, Wd, F, kh, kw, self.stride, self.pad_size[0], self.pad_size[1])
        st = current_stream()
        self._xb = None
        self._x_lo = None
        if self.precision == 'tf32x3' and self._x3_ok(0, geom):
            # near-fp32 tier on the tensor pipe: x*w = xh*wh + xh*wl + xl*wh (3xTF32), lo parts split off here
            self._x_lo = self._buf('x_lo', x.shape, x.layout)
            w_lo = self._buf('w_lo', (W.size,), 'flat')
            lib.sn_split_tf32_lo(x.ptr, self._x_lo.ptr, x.size, st)
            lib.sn_split_tf32_lo(W.output_tensor.ptr, w_lo.ptr, W.size, st)
            lib.sn_conv2d_fprop_x3(x.ptr, self._x_lo.ptr, W.output_tensor.ptr, w_lo.ptr, b.output_tensor.ptr, y.ptr, *geom,
                                   self._act_code(), st)
            bn = getattr(self, '_stats_to', None)
            if bn is not None:
                bn._stats_from = None
            self.output_tensor = y
            super(Conv2D, self).forward(is_training)
            return
        self._x4 = None
        if self.precision in ('tf32', 'bf16') and self._stem_ok(geom):
            self._forward_stem(x, W, b, geom, is_training, st)
            return
        bf16 = self.precision == 'bf16' and self._bf16_ok(0, geom)
        # the BatchNormalization fed by this layer alone takes its batch statistics from our epilogue
        bn = getattr(self, '_stats_to', None) if is_training else None
        scratch = None
        if bn is not None and self._stats_ok(bf16, geom):
            scratch = bn._device_stats(F)['scratch'].data_ptr()
            bn._stats_from = self
        elif bn is not None:
            bn._stats_from = None
        # bf16 tier: when that BatchNormalization runs fused with its 2x2 pool (it reads this map twice and
        # nobody else ever does), the map itself is stored in bf16 — written once, read twice, half the bytes
        if scratch and self.precision == 'bf16' and bn._takes_bf16_map((N, F, oh, ow)) and self._ybf16_ok(bf16, geom):
            yb = self._buf('out', (N, F, oh, ow), 'nhwc', 'bf16')
            if bf16:
                self._xb = self._bf16_buf('x_bf16', x.size)
                if getattr(self, '_xb_filled_for', None) is not x:
                    lib.sn_cast_f32_to_bf16(x.ptr, self._xb.data_ptr(), x.size, st)
                self._xb_filled_for = None
                prepared = getattr(self, '_w_prepared', None)
                if prepared is not None:
                    wb_ptr = prepared[0]
                else:
                    wb = self._bf16_buf('w_bf16', W.size)
                    lib.sn_cast_f32_to_bf16(W.output_tensor.ptr, wb.data_ptr(), W.size, st)
                    wb_ptr = wb.data_ptr()
                lib.sn_conv2d_fprop_bf16_stats_ybf16(self._xb.data_ptr(), wb_ptr, b.output_tensor.ptr, yb.ptr, *geom,
                                                     self._act_code(), scratch, st)
            else:
                lib.sn_conv2d_fprop_stats_ybf16(x.ptr, W.output_tensor.ptr, b.output_tensor.ptr, yb.ptr, *geom,
                                                self._act_code(), scratch, st)
            self.output_tensor = yb
            super(Conv2D, self).forward(is_training)
            return
        if bf16:
            # bf16 tier: bf16 copies of the input map and the filters, fp32 accumulation and output
            self._xb = self._bf16_buf('x_bf16', x.size)
            if getattr(self, '_xb_filled_for', None) is not x:       # else the producing kernel already wrote the copy
                lib.sn_cast_f32_to_bf16(x.ptr, self._xb.data_ptr(), x.size, st)
            self._xb_filled_for = None
            prepared = getattr(self, '_w_prepared', None)            # the model cast every filter of this step in one launch
            if prepared is not None:
                wb_ptr = prepared[0]
            else:
                wb = self._bf16_buf('w_bf16', W.size)
                lib.sn_cast_f32_to_bf16(W.output_tensor.ptr, wb.data_ptr(), W.size, st)
                wb_ptr = wb.data_ptr()
            if scratch:
                lib.sn_conv2d_fprop_bf16_stats(self._xb.data_ptr(), wb_ptr, b.output_tensor.ptr, y.ptr, *geom,
                                               self._act_code(), scratch, st)
            else:
                lib.sn_conv2d_fprop_bf16(self._xb.data_ptr(), wb_ptr, b.output_tensor.ptr, y.ptr, *geom,
                                         self._act_code(), st)
        elif scratch:
            lib.sn_conv2d_fprop_stats(x.ptr, W.output_tensor.ptr, b.output_tensor.ptr, y.ptr, *geom, self._act_code(),
                                      PRECISION[self.precision], scratch, st)
        else:
            lib.sn_conv2d_fprop(x.ptr, W.output_tensor.ptr, b.output_tensor.ptr, y.ptr, *geom, self._act_code(),
                                PRECISION[self.precision], st)
        self.output_tensor = y
        super(Conv2D, self).forward(is_training)

    def _stem_ok(self, geom):
        key = ('stem',) + tuple(geom)
        cache = self.__dict__.setdefault('_bf16_support', {})
        if key not in cache:
            cache[key] = bool(lib.load().sn_conv2d_stem_supported(*geom))
        return cache[key]

    def _forward_stem(self, x, W, b, geom, is_training, st):
        """Tiny-C convolution (the RGB stem): csrc/conv_stem_im2col.cu — the TMA unit gathers the column
        matrix (im2col mode) from an NHWC4 copy of the input, tcgen05 contracts it (TF32 products in both
        tensor tiers: K = 27 is far too short for operand precision to matter for speed)."""
        N, C, H, Wd, F = geom[:5]
        oh, ow = self.output_shape[2:]
        x4 = self._buf('x_nhwc4', (N * H * Wd * 4,), 'flat')
        lib.sn_conv2d_stem_prepare(x.ptr, x4.ptr, N, C, H, Wd, st)
        self._x4 = x4
        bn = getattr(self, '_stats_to', None) if is_training else None
        scratch = None
        if bn is not None:
            pow2 = F % 4 == 0 and (F // 4) & (F // 4 - 1) == 0 and F <= 1024     # what the BN apply kernels take
            if pow2:
                scratch = bn._device_stats(F)['scratch'].data_ptr()
                bn._stats_from = self
            else:
                bn._stats_from = None
        ybf16 = bool(scratch and self.precision == 'bf16' and bn._takes_bf16_map((N, F, oh, ow)) and F % 8 == 0
                     and __import__('os').environ.get('SN_BF16_MAPS', '1') != '0')
        y = self._buf('out', (N, F, oh, ow), 'nhwc', 'bf16' if ybf16 else 'f32')
        lib.sn_conv2d_stem_fprop(x4.ptr, W.output_tensor.ptr, b.output_tensor.ptr, y.ptr, *geom, self._act_code(), int(ybf16), scratch, st)
        self.output_tensor = y
        Layer.forward(self, is_training)

    def _stats_ok(self, bf16, geom):
        key = ('stats', bool(bf16)) + tuple(geom)
        cache = self.__dict__.setdefault('_bf16_support', {})
        if key not in cache:
            F = geom[4]
            pow2 = F % 4 == 0 and (F // 4) & (F // 4 - 1) == 0 and F <= 1024     # what the BN apply kernels take
            cache[key] = bool(pow2 and lib.load().sn_conv2d_fprop_stats_supported(PRECISION[self.precision], int(bf16), *geom))
        return cache[key]

    def _ybf16_ok(self, bf16_operands, geom):
        if __import__('os').environ.get('SN_BF16_MAPS', '1') == '0':
            return False
        key = ('ybf16', bool(bf16_operands)) + tuple(geom)
        cache = self.__dict__.setdefault('_bf16_support', {})
        if key not in cache:
            cache[key] = bool(lib.load().sn_conv2d_fprop_ybf16_supported(int(bool(bf16_operands)), *geom))
        return cache[key]

    def _bf16_input_buffer(self, shape):
        """For a producer that can write our bf16 operand copy itself (the fused BN+pool kernel): the
        buffer sn_conv2d_fprop_bf16 will read for an input of `shape`, or None when this layer's
        forward stays on the fp32 pointers.  The producer then sets `_xb_filled_for` to the tensor."""
        N, C, H, Wd = shape
        kh, kw = self.filter_size
        geom = (N, C, H, Wd, self.filter_nums, kh, kw, self.stride, self.pad_size[0], self.pad_size[1])
        if self.precision == 'bf16' and self._bf16_ok(0, geom):
            return self._bf16_buf('x_bf16', N * C * H * Wd)
        return None

    def _bf16_grad_plan(self):
        """(wgrad on bf16 operands?, dgrad on bf16 operands?) for the current input."""
        x = self.input_tensor
        N, C, H, Wd = x.shape
        kh, kw = self.filter_size
        geom = (N, C, H, Wd, self.filter_nums, kh, kw, self.stride, self.pad_size[0], self.pad_size[1])
        bf16 = self.precision == 'bf16'
        needs_dx = any(layer.require_grads for layer in self.inbounds)
        return (bool(bf16 and ((self._xb is not None and self._bf16_ok(2, geom)) or self._stem_wg16(geom))),
                bool(bf16 and needs_dx and self._bf16_ok(1, geom)))

    def _stem_wg16(self, geom):
        """The tiny-C stem: weight gradient from the fp32 input and the bf16 gradient map (csrc/conv_stem_tc.cu)."""
        if self._xb is not None:
            return False
        key = ('stem_wg16',) + tuple(geom)
        cache = self.__dict__.setdefault('_bf16_support', {})
        if key not in cache:
            cache[key] = bool(lib.load().sn_conv2d_stem_wgrad_dybf16_supported(*geom))
        return cache[key]

    def _x3_ok(self, op, geom):
        key = ('x3', op) + tuple(geom)
        cache = self.__dict__.setdefault('_bf16_support', {})
        if key not in cache:
            cache[key] = bool(lib.load().sn_conv2d_x3_supported(op, *geom))
        return cache[key]

    def _bf16_ok(self, op, geom):
        key = (op,) + tuple(geom)
        cache = self.__dict__.setdefault('_bf16_support', {})
        if key not in cache:
            cache[key] = bool(lib.load().sn_conv2d_bf16_supported(op, *geom))
        return cache[key]

    def _bf16_buf(self, name, n):
        from ..device import empty, torch
        key = (name, int(n))
        t = self._bufs.get(key)
        if t is None:
            t = empty(int(n), torch().bfloat16)
            self._bufs[key] = t
        return t

    def backward(self):
        """layers/Convolution.py:182-205."""
        W, b = self.variables
        x, y, g = self.input_tensor, self.output_tensor, self.grads
        N, C, H, Wd = x.shape
        F = self.filter_nums
        kh, kw = self.filter_size
        st = current_stream()
        prec = PRECISION[self.precision]
        if self._act_code() == ACT['relu'] and not self._grads_premasked:
            lib.sn_relu_bwd(g.ptr, y.ptr, g.size, st)
        geom = (N, C, H, Wd, F, kh, kw, self.stride, self.pad_size[0], self.pad_size[1])
        needs_dx = any(layer.require_grads for layer in self.inbounds)
        if self.precision == 'tf32x3':
            needs_dx = any(layer.require_grads for layer in self.inbounds)
            w3 = (W.require_grads or b.require_grads) and self._x3_ok(2, geom)
            d3 = needs_dx and self._x3_ok(1, geom)
            if w3 or d3:
                g_lo = self._buf('g_lo', g.shape, g.layout)
                lib.sn_split_tf32_lo(g.ptr, g_lo.ptr, g.size, st)
                db = b.grads.ptr if (b.require_grads and not self.__dict__.pop('_db_done', False)) else None
                if w3:
                    x_lo = self._x_lo
                    if x_lo is None:
                        x_lo = self._buf('x_lo', x.shape, x.layout)
                        lib.sn_split_tf32_lo(x.ptr, x_lo.ptr, x.size, st)
                    lib.sn_conv2d_wgrad_x3(x.ptr, x_lo.ptr, g.ptr, g_lo.ptr, W.grads.ptr, db, *geom, st)
                elif W.require_grads or db is not None:
                    lib.sn_conv2d_wgrad(x.ptr, g.ptr, W.grads.ptr, db, *geom, 1, prec, st)
                if d3:
                    ws3 = self._buf('dgrad_ws3', (2 * W.size,), 'flat').ptr
                    for layer in self.inbounds:
                        self._push_grad(layer, lambda dst, acc: lib.sn_conv2d_dgrad_x3(
                            g.ptr, g_lo.ptr, W.output_tensor.ptr, dst.ptr, *geom, acc, ws3, st))
                else:
                    ws_bytes = lib.sn_conv2d_dgrad_workspace(C, F, kh, kw, prec)
                    ws = self._buf('dgrad_ws', (ws_bytes // 4,), 'flat').ptr if ws_bytes else None
                    for layer in self.inbounds:
                        self._push_grad(layer, lambda dst, acc: lib.sn_conv2d_dgrad(
                            g.ptr, W.output_tensor.ptr, dst.ptr, *geom, acc, prec, ws, st))
                return
        wgrad_bf16, dgrad_bf16 = self._bf16_grad_plan()
        # a fused BN+pool backward may already have delivered the bias gradient (column sums of
        # our gradient map) and the bf16 copy of that map while it wrote it
        db_done = bool(self.__dict__.pop('_db_done', False))
        gb_ready = bool(self.__dict__.pop('_gb_ready', False))
        stem16 = bool(wgrad_bf16 and self._xb is None)
        if stem16 and not gb_ready:
            wgrad_bf16 = stem16 = False      # only worth it when the bf16 map comes for free (fused BN+pool backward)
        gb = None
        if wgrad_bf16 or dgrad_bf16:
            gb = self._bf16_buf('dy_bf16', g.size)
            if not gb_ready:
                lib.sn_cast_f32_to_bf16(g.ptr, gb.data_ptr(), g.size, st)
        db = b.grads.ptr if (b.require_grads and not db_done) else None
        if W.require_grads or db is not None:
            xb = self._xb

            def wgrad():
                if stem16:
                    lib.sn_conv2d_stem_wgrad_dybf16(x.ptr, gb.data_ptr(), W.grads.ptr, db, *geom, st)
                elif wgrad_bf16:
                    lib.sn_conv2d_wgrad_bf16(xb.data_ptr(), gb.data_ptr(), g.ptr if db is not None else None, W.grads.ptr, db, *geom, st)
                else:
                    lib.sn_conv2d_wgrad(x.ptr, g.ptr, W.grads.ptr, db, *geom, 1, prec, st)
            if getattr(self, '_defer_wgrad', False) and needs_dx:
                defer(wgrad)         # dX first; the weight gradient fills the gap in the upstream BatchNormalization's exchange
            else:
                wgrad()
        self._used_wt_bf16 = bool(dgrad_bf16)
        if dgrad_bf16:
            wtb = self._bf16_buf('wt_bf16', W.size)
            prepared = getattr(self, '_w_prepared', None)
            if prepared is None or not prepared[1]:
                lib.sn_flip_transpose_filters_bf16(W.output_tensor.ptr, wtb.data_ptr(), F, C, kh, kw, st)
            for layer in self.inbounds:
                self._push_grad(layer, lambda dst, acc: lib.sn_conv2d_dgrad_bf16(
                    gb.data_ptr(), wtb.data_ptr(), dst.ptr, *geom, acc, st))
            return
        ws_bytes = lib.sn_conv2d_dgrad_workspace(C, F, kh, kw, prec)
        ws = self._buf('dgrad_ws', (ws_bytes // 4,), 'flat').ptr if ws_bytes else None
        for layer in self.inbounds:
            self._push_grad(layer, lambda dst, acc: lib.sn_conv2d_dgrad(
                g.ptr, W.output_tensor.ptr, dst.ptr, *geom, acc, prec, ws, st))


class ZeroPadding2D(Layer):
    """layers/Convolution.py:209-251.  The reference needs this layer to get a correct SAME
    convolution.  Deviation: the reference marks it require_grads=False,
    which makes a Conv2D behind it hand over its raw output gradient (:204) so that the crop in
    :248 only type-checks when nothing upstream needs a gradient; here the layer carries the
    true dX through (crop of the padded gradient)."""

    def __init__(self, pad_size):
        assert len(pad_size) == 2
        self.pad_h, self.pad_w = pad_size
        super(ZeroPadding2D, self).__init__()

    def connect(self, prev_layer):
        (batch_nums, C, H, W) = prev_layer.output_shape
        self.output_shape = (batch_nums, C, H + 2 * self.pad_h, W + 2 * self.pad_w)
        Layer.connect(self, prev_layer)

    def __call__(self, prev_layer):
        super(ZeroPadding2D, self).__call__(prev_layer)
        (batch_nums, C, H, W) = self.input_shape
        self.output_shape = (batch_nums, C, H + 2 * self.pad_h, W + 2 * self.pad_w)
        return self

    def forward(self, is_training=True):
        x = self.input_tensor
        N, C, H, W = x.shape
        y = self._buf('out', (N, C, H + 2 * self.pad_h, W + 2 * self.pad_w), 'nhwc')
        lib.sn_pad2d_fwd(x.ptr, y.ptr, N, H, W, C, self.pad_h, self.pad_w, current_stream())
        self.output_tensor = y
        super(ZeroPadding2D, self).forward(is_training)

    def backward(self):
        if not self._grads_written:
            return
        g = self.grads
        N, C, Hp, Wp = g.shape
        st = current_stream()
        for layer in self.inbounds:
            self._push_grad(layer, lambda dst, acc: lib.sn_pad2d_bwd(
                g.ptr, dst.ptr, N, Hp - 2 * self.pad_h, Wp - 2 * self.pad_w, C, self.pad_h, self.pad_w, acc, st))


class MaxPooling2D(Layer):
    def __init__(self, pool_size, stride=None):
        assert len(pool_size) == 2
        assert pool_size[0] == pool_size[1]
        self.pool_size = pool_size
        self.stride = pool_size[0] if stride is None else stride
        super(MaxPooling2D, self).__init__()

    def _infer(self, input_shape):
        batch_nums, n_C, n_H_prev, n_W_prev = input_shape
        pool_h, pool_w = self.pool_size
        assert pool_w == pool_h
        # layers/Convolution.py:266-269
        self.mode = 'reshape' if pool_h == pool_w == self.stride else 'im2col'
        n_H, n_W = (n_H_prev - pool_h) // self.stride + 1, (n_W_prev - pool_w) // self.stride + 1
        self.output_shape = (batch_nums, n_C, n_H, n_W)

    def connect(self, prev_layer):
        self._infer(prev_layer.output_shape)
        Layer.connect(self, prev_layer)

    def __call__(self, prev_layer):
        super(MaxPooling2D, self).__call__(prev_layer)
        self._infer(self.input_shape)
        return self

    def _fused_buffers(self, N, C, H, W):
        """Output and code buffers for the BatchNormalization layer that computes this pool's forward
        as part of its own (csrc/bn_pool.cu)."""
        from ..device import empty, torch
        oh, ow = H // 2, W // 2
        y = self._buf('out', (N, C, oh, ow), 'nhwc')
        key = ('code', N * C * oh * ow)
        code = self._bufs.get(key)
        if code is None:
            code = empty(key[1], torch().uint8)
            self._bufs[key] = code
        self._code, self._code_mode = code, self.mode
        self.input_shape = (N, C, H, W)
        return y, code

    def forward(self, is_training=True):
        """layers/Convolution.py:287-324.  'reshape' mode keeps a tie bit-mask, 'im2col' mode the
        first-argmax index; both are one byte per output element in (n,i,j,c) order.  On maps
        whose side is not a multiple of the pool the reference's 'reshape' mode raises; here it
        floors like the layer's own declared output_shape."""
        fused = getattr(self, '_fused_output', None)
        if fused is not None:                       # already computed by the preceding BatchNormalization
            self._fused_output = None
            self._fused_bwd = True
            self.output_tensor = fused
            super(MaxPooling2D, self).forward(is_training)
            return
        self._fused_bwd = False
        x = self.input_tensor
        N, C, H, W = x.shape
        self.input_shape = x.shape
        p, s = self.pool_size[0], self.stride
        oh, ow = (H - p) // s + 1, (W - p) // s + 1
        y = self._buf('out', (N, C, oh, ow), 'nhwc')
        code = None
        if is_training:
            wide = 2 if (self.mode == 'reshape' and p * p > 8) else 1
            key = ('code', N * C * oh * ow * wide)
            code = self._bufs.get(key)
            if code is None:
                from ..device import empty, torch
                code = empty(key[1], torch().uint8)
                self._bufs[key] = code
        lib.sn_maxpool2d_fwd(x.ptr, y.ptr, code.data_ptr() if code is not None else None, N, H, W, C, p, s,
                             POOL_MODE[self.mode], current_stream())
        self._code, self._code_mode = code, self.mode
        self.output_tensor = y
        super(MaxPooling2D, self).forward(is_training)

    @property
    def argmax_index(self):
        """The reference's int64 argmax vector (layers/Convolution.py:303,309), order (n,i,j,c)."""
        if self._code_mode != 'im2col':
            raise AttributeError('argmax_index exists in im2col mode only')
        from ..device import torch
        torch().cuda.current_stream().synchronize()
        return self._code.cpu().numpy().astype(np.int64)

    def backward(self):
        """layers/Convolution.py:326-362."""
        if getattr(self, '_fused_bwd', False):
            return                                   # the fused BatchNormalization.backward consumes self.grads
        g = self.grads
        N, C, H, W = self.input_shape
        p, s = self.pool_size[0], self.stride
        st = current_stream()
        for layer in self.inbounds:
            self._push_grad(layer, lambda dst, acc: lib.sn_maxpool2d_bwd(
                g.ptr, self._code.data_ptr(), dst.ptr, N, H, W, C, p, s, POOL_MODE[self._code_mode], acc, st))

    def __getstate__(self):
        d = Layer.__getstate__(self)
        d.pop('_code', None)
        d.pop('_fused_output', None)
        return d
