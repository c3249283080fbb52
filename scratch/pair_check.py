# Correctness + timing of the CTA-pair igemm tiles against a float64 torch reference.
# usage: SN_IGEMM_PAIR=0|1 python scratch/pair_check.py
import os, sys, torch
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
    torch.manual_seed(0)
    x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(F, 3, 3, C, device='cuda') * 0.05; b = torch.randn(F, device='cuda')
    y = torch.empty(N, H, W, F, device='cuda')
    xb = x.to(torch.bfloat16); wb = w.to(torch.bfloat16)
    g = (N, C, H, W, F, 3, 3, 1, 1, 1)
    flops = 2.0 * N * H * W * F * C * 9
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), b.double(), padding=1).permute(0, 2, 3, 1)
    refb = torch.nn.functional.conv2d(xb.permute(0, 3, 1, 2).double(), wb.permute(0, 3, 1, 2).double(), b.double(), padding=1).permute(0, 2, 3, 1)
    y.fill_(float('nan'))
    lib.sn_conv2d_fprop_bf16(xb.data_ptr(), wb.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 0, st); torch.cuda.synchronize()
    eb = ((y.double() - refb).abs().max() / refb.abs().max()).item()
    y.fill_(float('nan'))
    lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 0, 1, st); torch.cuda.synchronize()
    et = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    a = t(lambda: lib.sn_conv2d_fprop_bf16(xb.data_ptr(), wb.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, st))
    d = t(lambda: lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, 1, st))
    print('pair=%s conv %s: bf16 err %.2e %.1f us (%.0f TF/s), tf32 err %.2e %.1f us (%.0f TF/s)' % (
        os.environ.get('SN_IGEMM_PAIR', '1'), str((N, C, H, W, F)), eb, a, flops / a / 1e6, et, d, flops / d / 1e6), flush=True)
    assert eb < 1e-5 and et < 2e-3, 'mismatch'
conv(512, 64, 16, 16, 128)
conv(512, 128, 8, 8, 256)
conv(512, 256, 4, 4, 256)
conv(512, 128, 16, 16, 64)      # dgrad-like shapes of cfg3 (C and F swapped): BN = 64, no pairs
conv(512, 256, 8, 8, 128)
conv(510, 64, 16, 16, 128)      # odd image count -> still an even number of pixel tiles? (510*256/128 = 1020)
