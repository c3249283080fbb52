# NOTE: needs the throw-away SN_LSTM_DBG switches (skip the FMA phase / the HBM stores) that were patched into recurrent.cu for
# the measurement quoted in DESIGN.md section 9 and removed again; with the committed library every line is the full kernel.
import os, sys, torch, ctypes
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
libc = ctypes.CDLL(None)
st = torch.cuda.current_stream().cuda_stream
T, N, H = 64, 256, 128
zx = torch.randn(T, N, 4 * H, device='cuda'); wh = torch.randn(4 * H, H, device='cuda') * 0.05
acts = torch.empty(T, N, 4 * H, device='cuda'); c_all = torch.zeros(T + 1, N, H, device='cuda'); h_all = torch.zeros(T + 1, N, H, device='cuda')
def t(fn, n=20):
    for i in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for dbg, name in ((0, 'full'), (1, 'no FMA phase'), (4, 'no HBM stores'), (5, 'no FMA, no HBM stores')):
    libc.setenv(b'SN_LSTM_DBG', str(dbg).encode(), 1)
    us = t(lambda: lib.sn_lstm_seq_fwd(zx.data_ptr(), wh.data_ptr(), None, None, acts.data_ptr(), c_all.data_ptr(), h_all.data_ptr(), None, T, N, H, st))
    print('%-24s %.1f us = %.2f us / step' % (name, us, us / T), flush=True)
