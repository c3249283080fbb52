"""ctypes binding of libshinnosuke_b200.so (the C ABI declared in include/shinnosuke_b200.h).

There is no CPU fallback: if the library cannot be loaded every compute entry point raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'shinnosuke_b200.h')
LIB_PATH = os.path.join(_HERE, 'lib', 'libshinnosuke_b200.so')

_CTYPES = {
    'int': ctypes.c_int, 'long long': ctypes.c_longlong, 'float': ctypes.c_float,
    'size_t': ctypes.c_size_t, 'void': None,
}


def parse_header(path=HEADER):
    """Returns {name: (restype, [argtypes])} for every SN_API prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'SN_API\s+([\w\s\*]+?)\s*\b(sn_\w+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if '*' in ret:
            restype = ctypes.c_char_p if 'char' in ret else ctypes.c_void_p
        else:
            restype = _CTYPES[ret]
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    base = re.sub(r'\b(const|unsigned)\b', '', a).strip()
                    base = ' '.join(base.split()[:-1])      # drop the parameter name
                    argtypes.append(_CTYPES[base])
        protos[name] = (restype, argtypes)
    return protos


class ShinnosukeLibError(RuntimeError):
    pass


class KernelProfiler(object):
    """Brackets every compute entry point with CUDA events on the launching stream (the last
    argument of every such call).  bench.py uses it for the per-kernel roofline; off by default."""
    SKIP = ('sn_event_', 'sn_stream_', 'sn_malloc', 'sn_free', 'sn_set_device', 'sn_device_', 'sn_memcpy_h2d',
            'sn_memcpy_d2h', 'sn_embedding_oob_count', 'sn_comm_', 'sn_graph_', 'sn_igemm_')

    # optional output pointers that change a call's algorithmic bytes: their presence (1 / 0) is appended
    # to the integer key, in this order (bench.kernel_work reads them from the end of the tuple)
    PTR_FLAGS = {'sn_bn_pool_bwd': (8, 9), 'sn_bn_pool_bwd_apply': (3, 4), 'sn_bn_pool_fwd_apply': (2,), 'sn_bn_pool_fwd_train': (4,),
                 'sn_bn_pool_bwd_xbf16': (8, 9), 'sn_bn_pool_bwd_apply_xbf16': (3, 4), 'sn_bn_pool_fwd_apply_xbf16': (2,)}

    def __init__(self, dll):
        self.dll = dll
        self.records = []          # (name, args, start_event, stop_event)
        self._pool = []

    def _event(self):
        if self._pool:
            return self._pool.pop()
        ev = ctypes.c_void_p()
        self.dll.sn_event_create(ctypes.byref(ev))
        return ev

    def call(self, name, fn, args):
        stream = args[-1]
        a, b = self._event(), self._event()
        self.dll.sn_event_record(a, stream)
        rc = fn(*args)
        self.dll.sn_event_record(b, stream)
        self.records.append((name, args, a, b))
        return rc

    def summary(self):
        """{(name, int-args): [ms, ...]}; synchronises.  Pointer arguments are dropped from the key."""
        out = {}
        ms = ctypes.c_float()
        for name, args, a, b in self.records:
            self.dll.sn_event_elapsed_ms(a, b, ctypes.byref(ms))
            ints = tuple(x for x in args[:-1] if isinstance(x, (int, float)) and not isinstance(x, bool) and abs(x) < (1 << 40))
            ints += tuple(1 if args[i] else 0 for i in self.PTR_FLAGS.get(name, ()))
            key = (name, ints)
            out.setdefault(key, []).append(ms.value)
            self._pool += [a, b]
        self.records = []
        return out


class _Lib(object):
    def __init__(self):
        self._dll = None
        self._err = None
        self.profiler = None

    def start_profiling(self):
        self.profiler = KernelProfiler(self.load())
        return self.profiler

    def stop_profiling(self):
        prof, self.profiler = self.profiler, None
        return prof.summary() if prof is not None else {}

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise ShinnosukeLibError(
                'shinnosuke_b200: CUDA library %s is missing; build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` (there is no CPU fallback)' % LIB_PATH)
        dll = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in parse_header().items():
            fn = getattr(dll, name)          # AttributeError if the .so is stale
            fn.restype = restype
            fn.argtypes = argtypes
        if dll.sn_abi_version() != 1:
            raise ShinnosukeLibError('libshinnosuke_b200.so ABI version mismatch')
        self._dll = dll
        return dll

    def comm_timeouts(self, comm_handle, reset=True):
        n = ctypes.c_longlong(0)
        self.sn_comm_timeouts(comm_handle, ctypes.byref(n), 1 if reset else 0)
        return int(n.value)

    def embedding_oob_count(self, reset=True):
        n = ctypes.c_longlong(0)
        self.sn_embedding_oob_count(ctypes.byref(n), 1 if reset else 0)
        return int(n.value)

    def __getattr__(self, name):
        dll = self.load()
        fn = getattr(dll, name)
        if fn.restype is ctypes.c_int and name not in ('sn_abi_version',):
            timed = not name.startswith(KernelProfiler.SKIP)

            def checked(*args, _fn=fn, _name=name, _timed=timed):
                prof = self.profiler
                rc = prof.call(_name, _fn, args) if (prof is not None and _timed) else _fn(*args)
                if rc != 0:
                    raise ShinnosukeLibError('%s failed (%d): %s' % (_name, rc, dll.sn_last_error().decode()))
            setattr(self, name, checked)
            return checked
        setattr(self, name, fn)
        return fn


lib = _Lib()

ACT = {'linear': 0, 'relu': 1}
# 'fp32'   : 1e-5 tier, CUDA-core FFMA kernels
# 'tf32x3' : 3xTF32 on the tensor pipe for the convolutions (every operator within 1e-5, trajectories 5e-5), fp32 elsewhere
PRECISION = {'fp32': 0, 'tf32x3': 0, 'tf32': 1, 'bf16': 2}
POOL_MODE = {'reshape': 0, 'im2col': 1}
