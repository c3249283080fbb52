"""CPU: the C-ABI library builds, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import subprocess

import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from shinnosuke_b200._lib import parse_header, LIB_PATH
    assert os.path.exists(LIB_PATH), 'run __graft_entry__.build() first'
    protos = parse_header()
    assert len(protos) >= 40
    out = subprocess.check_output(['nm', '-D', '--defined-only', LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if ' T ' in line}
    missing = sorted(set(protos) - exported)
    assert not missing, 'declared in the header but not exported: %s' % missing
    stray = sorted(s for s in exported if s.startswith('sn_') and s not in protos)
    assert not stray, 'exported but not declared: %s' % stray
    # only the C ABI is visible
    assert all(s.startswith('sn_') for s in exported if not s.startswith('_')), exported


def test_library_loads_and_reports_errors_without_aborting():
    from shinnosuke_b200._lib import lib, ShinnosukeLibError
    assert lib.sn_abi_version() == 1
    n = ctypes.c_int(-7)
    try:
        lib.sn_device_count(ctypes.byref(n))
    except ShinnosukeLibError as e:          # no driver in the CPU container
        assert 'cudaGetDeviceCount' in str(e)
        assert n.value == 0
    # argument validation happens before any CUDA call
    with pytest.raises(ShinnosukeLibError, match='bad shape|bad arguments|null'):
        lib.sn_maxpool2d_fwd(None, None, None, 1, 4, 4, 1, 2, 2, 0, None)
    with pytest.raises(ShinnosukeLibError, match='not supported'):
        lib.sn_maxpool2d_fwd(8, 8, None, 1, 40, 40, 1, 5, 5, 0, None)


def test_header_compiles_as_c():
    """include/shinnosuke_b200.h is plain C (no torch / C++ types in the signatures)."""
    src = '#include "shinnosuke_b200.h"\nint main(void){return sn_abi_version()==SN_ABI_VERSION?0:1;}\n'
    p = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-I', os.path.join(ROOT, 'include'),
                        '-x', 'c', '-'], input=src, text=True, capture_output=True)
    assert p.returncode == 0, p.stderr


def test_product_does_not_import_the_oracle():
    import re
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'shinnosuke_b200')):
        for f in files:
            if f.endswith('.py'):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M) or '/root/reference' in text:
                    bad.append(f)
    assert not bad


def test_every_dependent_launch_target_waits_for_its_predecessor():
    """Programmatic dependent launch (csrc/common.cuh): a kernel launched through launch_dependent() may
    start before the previous kernel of the stream has finished, so its body must call pdl_wait() before
    it touches global memory.  Static check over the sources: every kernel name passed to
    launch_dependent() has pdl_wait() in its definition, ahead of the first __ldg / global pointer
    dereference pattern we can recognise (the call must be in the first statements of the body or right
    after the shared-memory / barrier prologue)."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'shinnosuke_b200', 'csrc')
    text = {p: open(p).read() for p in glob.glob(os.path.join(root, '*.cu'))}
    launched = set()
    for src in text.values():
        launched.update(re.findall(r'launch_dependent\(\s*([A-Za-z_0-9]+)', src))
    assert len(launched) >= 20, launched
    for name in sorted(launched):
        bodies = []
        for src in text.values():
            for m in re.finditer(r'__global__[^;{]*?\b%s\s*\(' % re.escape(name), src):
                start = src.index('{', m.end())
                depth, i = 0, start
                while True:
                    depth += src[i] == '{'
                    depth -= src[i] == '}'
                    if depth == 0:
                        break
                    i += 1
                bodies.append(src[start:i])
        assert bodies, 'definition of %s not found' % name
        for body in bodies:
            assert 'pdl_wait();' in body, '%s is launched as a dependent but never calls pdl_wait()' % name
            head = body[:body.index('pdl_wait();')]
            assert '__ldg(' not in head and 'ldg4(' not in head and 'atomicAdd(' not in head, \
                '%s reads or updates global memory before pdl_wait()' % name
