# NOTE: needs the throw-away SN_IGEMM_DEBUG switches (skip loads / MMAs / stores) that were patched into conv_tc.cu for the
# measurement in profiles/r01_igemm_pair_vs_single.txt and removed again; with the committed library every column is the full kernel.
# timing decomposition of conv_igemm_kernel with the TEMPORARY SN_IGEMM_DEBUG switches (results are garbage, times are not)
import os, sys, torch, ctypes
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
libc = ctypes.CDLL(None)
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=20):
    for i in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
names = {16: 'full (TMA store)', 0: 'full', 1: 'no store', 2: 'no B', 8: 'no A', 10: 'no loads', 4: 'no MMA', 5: 'no MMA, no store', 11: 'no loads, no store', 14: 'barriers+epilogue', 15: 'barriers only'}
def conv(N, C, H, W, F):
    x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(F, 3, 3, C, device='cuda') * 0.05; b = torch.randn(F, device='cuda')
    y = torch.empty(N, H, W, F, device='cuda'); xb = x.to(torch.bfloat16); wb = w.to(torch.bfloat16)
    g = (N, C, H, W, F, 3, 3, 1, 1, 1)
    out = []
    for dbg in [int(v) for v in os.environ.get('DBG', '0,16,1').split(',')]:
        libc.setenv(b'SN_IGEMM_DEBUG', str(dbg).encode(), 1)
        a = t(lambda: lib.sn_conv2d_fprop_bf16(xb.data_ptr(), wb.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, st))
        d = t(lambda: lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, 1, st))
        out.append('%s: bf16 %.1f tf32 %.1f' % (names[dbg], a, d))
    print('pair=%s %s | ' % (os.environ.get('SN_IGEMM_PAIR', '1'), (N, C, H, W, F)) + ' | '.join(out), flush=True)
conv(512, 64, 16, 16, 128)
conv(512, 128, 8, 8, 256)
conv(2048, 128, 8, 8, 256)
