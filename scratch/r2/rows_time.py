import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from shinnosuke_b200._lib import lib
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    torch.cuda._sleep(int(4e8))      # the host runs ahead of the device: intervals between events are kernel time, not call time
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)])) * 1e3
for (N, C, H, F) in [(512, 64, 16, 128), (128, 64, 32, 64), (512, 128, 8, 256)]:
    x = torch.randn(N * H * H * C, device='cuda').to(torch.bfloat16)
    w = (0.05 * torch.randn(F * 9 * C, device='cuda')).to(torch.bfloat16)
    b = torch.zeros(F, device='cuda')
    y = torch.empty(N * H * H * F, device='cuda')
    yb = torch.empty(N * H * H * F, device='cuda', dtype=torch.bfloat16)
    scr = torch.zeros(28 * F + 2, device='cuda', dtype=torch.float64)
    g = (N, C, H, H, F, 3, 3, 1, 1, 1)
    fl = 2.0 * N * H * H * F * C * 9
    us = t(lambda: lib.sn_conv2d_fprop_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, st))
    us2 = t(lambda: lib.sn_conv2d_fprop_bf16_stats_ybf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), yb.data_ptr(), *g, 1, scr.data_ptr(), st))
    # dgrad: dy (N,H,H,F) bf16, wt [C][9][F] bf16 -> dx (N,H,H,C) fp32
    dy = torch.randn(N * H * H * F, device='cuda').to(torch.bfloat16)
    wt = (0.05 * torch.randn(C * 9 * F, device='cuda')).to(torch.bfloat16)
    dx = torch.empty(N * H * H * C, device='cuda')
    us3 = t(lambda: lib.sn_conv2d_dgrad_bf16(dy.data_ptr(), wt.data_ptr(), dx.data_ptr(), *g, 0, st))
    print('N%d C%d %dx%d F%d rows=%s: fprop f32-out %.1f us (%.0f TF/s), fprop bf16-out+stats %.1f us (%.0f TF/s), dgrad %.1f us (%.0f TF/s)  last=%s'
          % (N, C, H, H, F, os.environ.get('SN_IGEMM_ROWS', '1'), us, fl / us / 1e6, us2, fl / us2 / 1e6, us3, fl / us3 / 1e6, lib.load().sn_last_kernel()), flush=True)
