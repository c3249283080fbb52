"""Device buffers for the CUDA path.

PyTorch is used here only as plumbing: its caching allocator owns the HBM, its current stream
is the stream every C-ABI call is issued on, and ``torch.distributed`` carries the gradient
all-reduce.  All arithmetic is done by libshinnosuke_b200.so.

A :class:`DeviceTensor` is an fp32 buffer plus the *logical* shape the reference would see and a
layout tag saying how that logical tensor is stored:

    'flat'    : stored exactly as the logical C-order array
    'nhwc'    : logical (N, C, H, W)         stored as [N][H][W][C]
    'krsc'    : logical (F, C, kh, kw)       stored as [F][kh][kw][C]
    'dense_w' : logical (n_in, n_out)        stored as [n_out][n_in]

and two parametrised layouts (tuples) that let Flatten hand an NHWC map to the Dense behind it
without a transpose kernel (the reference flattens NCHW in (c,i,j) order, layers/FC.py:39-48):

    ('flat_hwc', C, H, W)    : logical (N, C*H*W) in (c,i,j) order    stored as [N][H][W][C]
    ('dense_w_hwc', C, H, W) : logical (C*H*W, n_out), rows (c,i,j)   stored as [n_out][H][W][C]

and the LSTM weight matrix of layers/Recurrent.py:246-250, rows [recurrent; input], columns (f,u,c~,o):

    ('lstm_w', H)            : logical (H + D, 4H)                    stored as [4H][H] then [4H][D]
"""
import ctypes

import numpy as np

from ._lib import lib, ShinnosukeLibError

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise ShinnosukeLibError('shinnosuke_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    lib.load()
    return t


def current_stream():
    """cudaStream_t of torch's current stream as an int (what the C ABI takes as void*)."""
    return torch().cuda.current_stream().cuda_stream


_POISON = bool(int(__import__('os').environ.get('SN_DEBUG_POISON', '0')))


def empty(n, dtype=None):
    t = require_cuda()
    buf = t.empty(int(n), dtype=dtype or t.float32, device='cuda')
    if _POISON and t.cuda.is_current_stream_capturing():
        raise RuntimeError('device buffer of %d elements allocated while a CUDA graph is being captured' % int(n))
    if _POISON and buf.is_floating_point():
        # debugging aid (tests/test_gpu_train.py::test_no_uninitialised_reads): fresh buffers start as
        # NaN, so a kernel that reads memory nobody wrote shows up in the results
        buf.fill_(float('nan'))
        t.cuda.synchronize()            # the fill runs on the current stream; other streams may write the buffer next
    return buf


def zeros(n, dtype=None):
    buf = empty(n, dtype)
    lib.sn_memset(buf.data_ptr(), 0, buf.numel() * buf.element_size(), current_stream())
    return buf


def _to_storage(arr, layout):
    """logical float64/float32 ndarray -> C-contiguous float32 ndarray in storage order."""
    a = np.asarray(arr)
    if layout == 'nhwc':
        a = a.transpose(0, 2, 3, 1)
    elif layout == 'krsc':
        a = a.transpose(0, 2, 3, 1)
    elif layout == 'dense_w':
        a = a.T
    elif isinstance(layout, tuple) and layout[0] == 'flat_hwc':
        _, c, h, w = layout
        a = a.reshape(a.shape[0], c, h, w).transpose(0, 2, 3, 1)
    elif isinstance(layout, tuple) and layout[0] == 'dense_w_hwc':
        _, c, h, w = layout
        a = a.T.reshape(a.shape[1], c, h, w).transpose(0, 2, 3, 1)
    elif isinstance(layout, tuple) and layout[0] == 'lstm_w':
        h = layout[1]
        a = np.concatenate((a[:h].T.ravel(), a[h:].T.ravel()))
    return np.ascontiguousarray(a, dtype=np.float32)


def _from_storage(a, shape, layout):
    if layout == 'nhwc':
        n, c, h, w = shape
        a = a.reshape(n, h, w, c).transpose(0, 3, 1, 2)
    elif layout == 'krsc':
        f, c, kh, kw = shape
        a = a.reshape(f, kh, kw, c).transpose(0, 3, 1, 2)
    elif layout == 'dense_w':
        n_in, n_out = shape
        a = a.reshape(n_out, n_in).T
    elif isinstance(layout, tuple) and layout[0] == 'flat_hwc':
        _, c, h, w = layout
        a = a.reshape(shape[0], h, w, c).transpose(0, 3, 1, 2).reshape(shape)
    elif isinstance(layout, tuple) and layout[0] == 'dense_w_hwc':
        _, c, h, w = layout
        n_in, n_out = shape
        a = a.reshape(n_out, h, w, c).transpose(0, 3, 1, 2).reshape(n_out, n_in).T
    elif isinstance(layout, tuple) and layout[0] == 'lstm_w':
        h = layout[1]
        rows, cols = shape
        a = np.concatenate((a[:cols * h].reshape(cols, h).T, a[cols * h:].reshape(cols, rows - h).T), axis=0)
    else:
        a = a.reshape(shape)
    return np.ascontiguousarray(a, dtype=np.float64)


def layout_for(shape, kind=None):
    if kind is not None:
        return kind
    return 'nhwc' if len(shape) == 4 else 'flat'


class DeviceTensor(object):
    """`dtype` is 'f32' for everything except the bf16 tier's convolution output maps in front of a fused
    BatchNormalization ('bf16': the map is written once and read twice, at half the bytes); host views
    are float64 either way."""
    __slots__ = ('buf', 'shape', 'layout', 'ptr', 'dtype')

    def __init__(self, buf, shape, layout='flat', dtype='f32'):
        self.buf = buf
        self.shape = tuple(int(s) for s in shape)
        self.layout = layout
        self.ptr = buf.data_ptr()
        self.dtype = dtype
        assert buf.numel() == int(np.prod(self.shape)), (buf.numel(), self.shape)

    # ---- construction -------------------------------------------------
    @classmethod
    def empty(cls, shape, layout=None, dtype='f32'):
        shape = tuple(int(s) for s in shape)
        if dtype == 'bf16':
            return cls(empty(int(np.prod(shape)), torch().bfloat16), shape, layout_for(shape, layout), 'bf16')
        return cls(empty(int(np.prod(shape))), shape, layout_for(shape, layout))

    @classmethod
    def zeros(cls, shape, layout=None):
        shape = tuple(int(s) for s in shape)
        return cls(zeros(int(np.prod(shape))), shape, layout_for(shape, layout))

    @classmethod
    def from_numpy(cls, arr, layout=None):
        arr = np.asarray(arr)
        layout = layout_for(arr.shape, layout)
        out = cls.empty(arr.shape, layout)
        out.set(arr)
        return out

    def set(self, arr):
        """Upload a logical-layout host array (any float dtype) into this buffer."""
        arr = np.asarray(arr)
        assert tuple(arr.shape) == self.shape, (arr.shape, self.shape)
        assert self.dtype == 'f32', 'bf16 maps are written by kernels only'
        host = _to_storage(arr, self.layout)
        lib.sn_memcpy_h2d(self.ptr, host.ctypes.data, host.nbytes, current_stream())
        lib.sn_stream_sync(current_stream())        # `host` is pageable and about to be freed

    # ---- host views ----------------------------------------------------
    def numpy(self):
        """float64 array in the reference's layout (synchronises)."""
        if self.dtype == 'bf16':
            raw = np.empty(self.buf.numel(), dtype=np.uint16)
            lib.sn_memcpy_d2h(raw.ctypes.data, self.ptr, raw.nbytes, current_stream())
            lib.sn_stream_sync(current_stream())
            host = (raw.astype(np.uint32) << 16).view(np.float32)
            return _from_storage(host, self.shape, self.layout)
        host = np.empty(self.buf.numel(), dtype=np.float32)
        lib.sn_memcpy_d2h(host.ctypes.data, self.ptr, host.nbytes, current_stream())
        lib.sn_stream_sync(current_stream())
        return _from_storage(host, self.shape, self.layout)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    @property
    def size(self):
        return self.buf.numel()

    @property
    def ndim(self):
        return len(self.shape)

    def view(self, shape, layout='flat'):
        return DeviceTensor(self.buf, shape, layout, self.dtype)

    def zero_(self):
        lib.sn_memset(self.ptr, 0, self.buf.numel() * self.buf.element_size(), current_stream())
        return self

    def copy_from(self, other):
        assert other.buf.numel() == self.buf.numel() and other.dtype == self.dtype
        lib.sn_memcpy_d2d(self.ptr, other.ptr, self.buf.numel() * self.buf.element_size(), current_stream())
        return self

    def add_(self, other, alpha=1.0):
        assert self.dtype == 'f32' and other.dtype == 'f32'
        lib.sn_axpy(float(alpha), other.ptr, self.ptr, self.buf.numel(), current_stream())
        return self

    def __repr__(self):
        return 'DeviceTensor(shape=%s, layout=%s)' % (self.shape, self.layout)


def as_device(x, layout=None):
    """Accepts a DeviceTensor or anything array-like in the reference's layout."""
    if isinstance(x, DeviceTensor):
        return x
    return DeviceTensor.from_numpy(np.asarray(x), layout)


class Arena(object):
    """One flat fp32 allocation holding every trainable parameter (and a twin for the
    gradients), so the optimizer is ONE kernel and the data-parallel all-reduce ONE call."""
    ALIGN = 32          # floats (128 bytes): TMA global addresses want >= 16-byte alignment

    @classmethod
    def layout(cls, sizes, extra=0):
        offs, total = [], 0
        for s in sizes:
            offs.append(total)
            total += (int(s) + cls.ALIGN - 1) // cls.ALIGN * cls.ALIGN
        return offs, total, total + extra

    def __init__(self, sizes, extra=0, g_alloc=None):
        """g_alloc(n) -> zero-filled float32 tensor: lets the gradient arena live in the data-parallel
        communicator's peer-mapped window, where it is all-reduced in place (comm.py)."""
        self.offsets, self.param_floats, self.total = self.layout(sizes, extra)
        self.w = zeros(self.total)
        self.g = g_alloc(self.total) if g_alloc is not None else zeros(self.total)
        self.v = None       # momentum velocity, allocated on demand
        self.gstep = None   # data-parallel Momentum: per-step gradients

    def slice(self, which, i, size):
        buf = getattr(self, which)
        return buf[self.offsets[i]:self.offsets[i] + int(size)]
