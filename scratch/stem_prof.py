import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st = torch.cuda.current_stream().cuda_stream
N, C, H, W, F = 512, 3, 32, 32, 64
x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(F, 3, 3, C, device='cuda') * 0.1; b = torch.randn(F, device='cuda')
y = torch.empty(N, H, W, F, device='cuda'); dy = torch.randn(N, H, W, F, device='cuda')
dw = torch.zeros(F, 3, 3, C, device='cuda'); db = torch.zeros(F, device='cuda')
scr = torch.zeros(28 * F + 2, dtype=torch.float64, device='cuda')
for i in range(3):
    lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)
    scr.zero_()
    lib.sn_conv2d_fprop_stats(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, scr.data_ptr(), st)
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)
torch.cuda.synchronize()
# event timing
def t(fn, n=20):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print('fprop       us', t(lambda: lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)))
print('fprop_stats us', t(lambda: lib.sn_conv2d_fprop_stats(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, scr.data_ptr(), st)))
print('fprop lin   us', t(lambda: lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 0, 1, st)))
print('wgrad+db    us', t(lambda: lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)))
print('wgrad       us', t(lambda: lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)))
