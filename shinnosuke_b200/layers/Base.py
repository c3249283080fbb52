"""Graph plumbing of the CUDA path: Variable / Layer / Input.

Mirrors the operator interface of the reference's layers/Base.py (Node :7-29, Variable
:107-115, Layer :168-276, Input :283-308): the same attributes (inbounds, outbound_layers,
input/output_shape, input/output_tensor, grads, variables, require_grads), the same
push-forward / accumulate-backward data flow.  Tensors are DeviceTensor handles instead of
ndarrays; ``np.asarray(t)`` gives the reference's float64 NCHW view back.
"""
import numpy as np

from ..device import DeviceTensor, as_device, current_stream
from .._lib import lib


# Work a layer's backward() has postponed (closures that only launch kernels): a Conv2D whose dX feeds a
# data-parallel BatchNormalization hands its weight-gradient launch over, so that it runs BETWEEN the two halves of
# that BatchNormalization's cross-replica exchange (layers/Normalization.py).  The model flushes what is left at
# the end of the backward sweep.  One list per host thread (fit(n_gpus=...) runs one replica per thread).
import threading as _threading
_deferred = _threading.local()


def defer(fn):
    if not hasattr(_deferred, 'work'):
        _deferred.work = []
    _deferred.work.append(fn)


def flush_deferred():
    work = getattr(_deferred, 'work', None)
    while work:
        work.pop(0)()


class Variable(object):
    """A trainable tensor.  Holds a host float64 array (the reference's initial value) until
    the model places it in its parameter arena; afterwards ``output_tensor`` / ``grads`` are
    DeviceTensor views into the arena (layers/Base.py:107-115)."""

    def __init__(self, initial_value=None, shape=None, name='variable', layout=None):
        self.inbounds = []
        self.outbound_layers = []
        self.require_grads = True
        self.name = name
        self.layout = layout            # storage layout on the device ('krsc', 'dense_w', 'flat')
        self._host = None if initial_value is None else np.array(initial_value, dtype=np.float64)
        self._host_grads = None
        self.output_shape = shape if self._host is None else self._host.shape
        self._dev = None
        self._grads = None

    # -- value ------------------------------------------------------------
    @property
    def output_tensor(self):
        return self._dev if self._dev is not None else self._host

    @output_tensor.setter
    def output_tensor(self, value):
        if isinstance(value, DeviceTensor):
            self._dev = value
            return
        value = np.array(value, dtype=np.float64)
        if self._dev is not None:
            self._dev.set(value)
        else:
            self._host = value
            self.output_shape = value.shape

    @property
    def grads(self):
        if self._grads is None and self._dev is None and self._host is not None and self.require_grads:
            return np.zeros_like(self._host)
        return self._grads

    @grads.setter
    def grads(self, value):
        if isinstance(value, DeviceTensor) or value is None:
            self._grads = value
        elif self._grads is not None:
            self._grads.set(np.asarray(value))
        # host-side zero arrays assigned before the device exists carry no information

    def get_value(self):
        return np.asarray(self.output_tensor)

    def get_shape(self):
        return tuple(self.output_shape)

    @property
    def size(self):
        return int(np.prod(self.output_shape))

    # -- device placement -------------------------------------------------
    def bind(self, wbuf, gbuf):
        """Adopt arena slices (torch views) as storage and upload the host value."""
        layout = self.layout or 'flat'
        self._dev = DeviceTensor(wbuf, self.output_shape, layout)
        self._grads = DeviceTensor(gbuf, self.output_shape, layout)
        self._dev.set(self._host)
        pending = getattr(self, '_host_grads', None)
        if pending is not None:                 # unpickled mid-training: the accumulated gradients come back too
            self._grads.set(pending)
            self._host_grads = None

    def ensure_device(self):
        """Stand-alone placement for op-level use of a layer outside a model."""
        if self._dev is None:
            layout = self.layout or 'flat'
            self._dev = DeviceTensor.from_numpy(self._host, layout)
            self._grads = DeviceTensor.zeros(self.output_shape, layout)
            pending = getattr(self, '_host_grads', None)
            if pending is not None:
                self._grads.set(pending)
                self._host_grads = None
        return self._dev

    def relayout(self, layout):
        """Change the device storage order (before or after placement; the logical value is kept)."""
        if layout == self.layout:
            return
        if self._dev is not None:
            self._host = self._dev.numpy()
            grads = self._grads.numpy() if self._grads is not None else None
            self.layout = layout
            self._dev = DeviceTensor(self._dev.buf, self.output_shape, layout)
            self._dev.set(self._host)
            if self._grads is not None:
                self._grads = DeviceTensor(self._grads.buf, self.output_shape, layout)
                self._grads.set(grads)
        else:
            self.layout = layout

    def to_host(self):
        """Pull the current value back into the host copy (pickling, inspection)."""
        if self._dev is not None:
            self._host = self._dev.numpy()
        return self._host

    # -- pickling: the reference's format is NumPy arrays only (models.py:208-220 pickles the layers,
    # whose Variables hold `output_tensor` and `grads` ndarrays).  Device views into the parameter arena
    # must never reach the pickle: torch would serialise the whole backing storage once per view.
    def __getstate__(self):
        d = dict(self.__dict__)
        if self._dev is not None:
            d['_host'] = self._dev.numpy()
        g = None
        if self._grads is not None:
            g = self._grads.numpy()
            # SGD leaves the gradients zeroed (utils/Optimizers.py:32); Momentum and the adaptive optimizers
            # never zero them (:52-55), so a resumed run needs them back
            g = g.astype(np.float32) if np.any(g) else None
        d['_host_grads'] = g
        d['_dev'] = None
        d['_grads'] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.__dict__.setdefault('_host_grads', None)


class Layer(object):
    def __init__(self, inbounds=None, outbound_layers=None, input_shape=None, output_shape=None,
                 input_tensor=None, output_tensor=None, variables=None):
        self.inbounds = [] if inbounds is None else [inbounds]
        self.outbound_layers = [] if outbound_layers is None else [outbound_layers]
        self.input_shape = input_shape
        self.output_shape = output_shape
        self.input_tensor = input_tensor
        self.output_tensor = output_tensor
        self.variables = [] if variables is None else list(variables)
        self.counts = 0
        self.require_grads = True
        self.grads = None
        self.precision = 'fp32'
        self._bufs = {}
        self._grads_written = False      # first writer assigns, later writers accumulate
        self._grads_premasked = False    # the consumer already applied this layer's ReLU mask

    # -- wiring (layers/Base.py:190-212) -----------------------------------
    def __call__(self, inbound):
        if not isinstance(inbound, (list, tuple)):
            inbound = [inbound]
        for layer in inbound:
            self.inbounds.append(layer)
            self.input_shape = layer.output_shape
            layer.outbound_layers.append(self)
        return self

    def connect(self, inbound):
        if inbound is not None:
            self.inbounds.append(inbound)
            self.input_shape = inbound.output_shape
            inbound.outbound_layers.append(self)

    def _initial_params(self, *args):
        pass

    def compute_output_shape(self, *args):
        pass

    # -- buffers ----------------------------------------------------------
    def _buf(self, name, shape, layout=None, dtype='f32'):
        """A device buffer owned by this layer, one per (name, shape), never freed while the
        layer lives: stable addresses are what make the training step capturable in a CUDA
        graph, also when a ragged last minibatch alternates with full ones."""
        shape = tuple(int(s) for s in shape)
        key = (name, shape) if dtype == 'f32' else (name, shape, dtype)
        t = self._bufs.get(key)
        if t is None:
            t = DeviceTensor.empty(shape, layout, dtype)
            self._bufs[key] = t
        return t

    def _param_ptrs(self):
        out = []
        for v in self.variables:
            v.ensure_device()
            out.append(v)
        return out

    # -- data flow (layers/Base.py:224-229) ---------------------------------
    def forward(self, is_training=True):
        for layer in self.outbound_layers:
            layer.input_tensor = self.output_tensor
        if is_training and self.require_grads:
            # The reference allocates zeros here and consumers `+=` into it.  On the device the
            # first consumer overwrites and later ones accumulate: same values, no memset pass.
            self.grads = self._buf('grads', self.output_tensor.shape, self.output_tensor.layout)
            self._grads_written = False
            self._grads_premasked = False

    def backward(self):
        raise NotImplementedError

    def _fused_relu(self):
        """True when this layer's output carries a ReLU whose backward mask a consumer may apply."""
        act = getattr(self, 'activator', None)
        return act is not None and act.__class__.__name__ == 'Relu'

    def _push_grad(self, layer, writer):
        """Deliver dX to an inbound layer: ``writer(dst, accumulate)`` launches the kernel.
        Reference rule (e.g. layers/FC.py:148-153): `layer.grads += dX` if it requires grads."""
        if not layer.require_grads:
            return
        acc = 1 if layer._grads_written else 0
        writer(layer.grads, acc)
        layer._grads_written = True

    def _hand_over_grads(self, layer, g):
        """dX of a pass-through layer IS its own gradient buffer: when `layer` has no other consumer
        (nobody else will add to its gradient) it adopts the buffer instead of receiving a copy.  The
        caller must not need `g` afterwards (its backward is over).  Returns False when a copy /
        accumulation is needed."""
        if not layer.require_grads or layer._grads_written or len(layer.outbound_layers) != 1:
            return False
        if layer.grads is None or layer.grads.size != g.size:
            return False
        layer.grads = g.view(layer.grads.shape, layer.grads.layout) if (layer.grads.shape != g.shape or layer.grads.layout != g.layout) else g
        layer._grads_written = True
        return True

    def _seed_grads(self, value):
        """Used by the loss (models.py:80): `outputs.grads = loss.backward(...)`."""
        self.grads = value
        self._grads_written = True

    def get_value(self):
        return np.asarray(self.output_tensor)

    def feed(self, inputs, objects='inputs'):
        """Test hook of the reference (layers/Base.py:263-276)."""
        if isinstance(inputs, (Variable, Layer)):
            inputs = inputs.output_tensor
        t = as_device(inputs)
        if objects == 'inputs':
            self.input_tensor = t
        elif objects == 'outputs':
            self.output_tensor = t
        elif objects == 'grads':
            self._seed_grads(t)
        elif objects[:-1] == 'variable':
            self.variables[int(objects[-1])].output_tensor = np.asarray(inputs)
        else:
            raise ValueError('unknown objects type,only support for - inputs/outputs/grads/variable*')

    # -- pickling: drop device state, keep host copies ---------------------
    def __getstate__(self):
        d = dict(self.__dict__)
        for v in self.variables:
            if isinstance(v, Variable):
                v.to_host()
        d['_bufs'] = {}
        for k in ('input_tensor', 'output_tensor', 'grads'):
            d[k] = None
        # device handles and per-step hand-over flags of the layers (bf16 operand copies, 3xTF32 lo parts,
        # prepared filters, fused-pool outputs): none of them belongs in a checkpoint
        for k in ('_xb', '_bf16_support', '_x_lo', '_w_prepared', '_xb_filled_for', '_code', '_fused_output', '_stats_from',
                  '_db_done', '_gb_ready', '_par', '_x4'):
            d.pop(k, None)
        return d


class Input(Layer):
    """layers/Base.py:283-308.  Accepts host arrays in the reference's layout (N,C,H,W) or
    (N,features) and hands the device copy (NHWC) to its consumers."""

    def __init__(self, shape, value=None):
        super(Input, self).__init__()
        self.input_shape = shape
        self.output_shape = self.input_shape
        self.output_tensor = value
        self.require_grads = False

    def connect(self, prev_layer):
        if prev_layer is not None:
            self.input_shape = prev_layer.output_shape
            self.output_shape = self.input_shape

    def forward(self, is_training=True):
        self.output_tensor = as_device(self.input_tensor)
        super(Input, self).forward(is_training)

    def backward(self):
        pass


def upload_batch(layer, host_array):
    """Host (N,C,H,W)/(N,K) float array -> device tensor in storage order, through the
    layer-owned staging buffers (NCHW upload + sn_nchw_to_nhwc on the device)."""
    a = np.ascontiguousarray(host_array, dtype=np.float32)
    st = current_stream()
    if a.ndim == 4 and a.shape[1] > 1 and a.shape[2] * a.shape[3] > 1:
        n, c, h, w = a.shape
        stage = layer._buf('stage_nchw', a.shape, 'flat')
        out = layer._buf('stage_nhwc', a.shape, 'nhwc')
        lib.sn_memcpy_h2d(stage.ptr, a.ctypes.data, a.nbytes, st)
        lib.sn_nchw_to_nhwc(stage.ptr, out.ptr, n, c, h, w, st)
        lib.sn_stream_sync(st)
        return out
    out = layer._buf('stage_nhwc', a.shape, 'nhwc' if a.ndim == 4 else 'flat')
    lib.sn_memcpy_h2d(out.ptr, a.ctypes.data, a.nbytes, st)
    lib.sn_stream_sync(st)
    return out


class Add(Layer):
    """Element-wise sum of two layers (reference: layers/Base.py:352-411), the residual connection
    of BASELINE config 4.  The reference's `Add.__call__` dereferences `output_tensor.shape` of its
    operands at graph-build time, when it is still None (:358-359), and its model then
    registers the operand LAYERS as trainable variables (:324); here the shape comes from
    `output_shape` and the layer has no variables.  Operands must have equal shapes (no
    broadcasting: the broadcast-reduction branch of :392-409 is off the training path)."""

    def __call__(self, inbounds):
        if not isinstance(inbounds, (list, tuple)) or len(inbounds) != 2:
            raise ValueError('Add takes a list of two layers')
        x, y = inbounds
        if tuple(x.output_shape[1:]) != tuple(y.output_shape[1:]):
            raise NotImplementedError('device Add needs operands of equal shape, got %s and %s'
                                      % (x.output_shape, y.output_shape))
        super(Add, self).__call__(list(inbounds))
        self.output_shape = tuple(x.output_shape)
        return self

    def connect(self, inbound):
        raise TypeError('Add has two operands: use the functional API, Add()([a, b])')

    def forward(self, is_training=True):
        a, b = (layer.output_tensor for layer in self.inbounds)
        out = self._buf('out', a.shape, a.layout)
        if getattr(self, '_relu_out', None) is not None:      # Add -> Activation('relu'), folded by compile()
            lib.sn_add_relu(a.ptr, b.ptr, out.ptr, a.size, current_stream())
        else:
            lib.sn_add(a.ptr, b.ptr, out.ptr, a.size, current_stream())
        self.output_tensor = out
        super(Add, self).forward(is_training)

    def backward(self):
        g = self.grads
        # the operand with no other consumer adopts the buffer instead of receiving a copy; it comes
        # LAST, so the other operand's copy is enqueued before the adopter's backward can touch the buffer
        order = sorted(self.inbounds, key=lambda l: len(l.outbound_layers) == 1)
        for i, layer in enumerate(order):
            if i == len(order) - 1 and self._hand_over_grads(layer, g):
                continue
            self._push_grad(layer, lambda dst, acc: dst.add_(g) if acc else dst.copy_from(g))
