import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=20):
    for i in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
def conv(N, C, H, W, F):
    x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(F, 3, 3, C, device='cuda') * 0.05; b = torch.randn(F, device='cuda')
    y = torch.empty(N, H, W, F, device='cuda')
    xb = x.to(torch.bfloat16); wb = w.to(torch.bfloat16)
    scr = torch.zeros(28 * F + 2, dtype=torch.float64, device='cuda')
    g = (N, C, H, W, F, 3, 3, 1, 1, 1)
    flops = 2.0 * N * H * W * F * C * 9
    a = t(lambda: lib.sn_conv2d_fprop_bf16(xb.data_ptr(), wb.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, st))
    c = t(lambda: lib.sn_conv2d_fprop_bf16_stats(xb.data_ptr(), wb.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, scr.data_ptr(), st))
    d = t(lambda: lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, 1, st))
    print('conv %s: bf16 fprop %.1f us (%.0f TF/s), +stats %.1f us, tf32 fprop %.1f us (%.0f TF/s)' % (str((N, C, H, W, F)), a, flops / a / 1e6, c, d, flops / d / 1e6))
conv(512, 64, 16, 16, 128)
conv(512, 128, 8, 8, 256)
conv(512, 256, 4, 4, 256)
