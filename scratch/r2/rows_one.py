import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from shinnosuke_b200._lib import lib
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
N, C, H, F = 512, 64, 16, 128
x = torch.randn(N * H * H * C, device='cuda').to(torch.bfloat16)
w = (0.05 * torch.randn(F * 9 * C, device='cuda')).to(torch.bfloat16)
b = torch.zeros(F, device='cuda')
yb = torch.empty(N * H * H * F, device='cuda', dtype=torch.bfloat16)
scr = torch.zeros(28 * F + 2, device='cuda', dtype=torch.float64)
g = (N, C, H, H, F, 3, 3, 1, 1, 1)
for _ in range(4):
    lib.sn_conv2d_fprop_bf16_stats_ybf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), yb.data_ptr(), *g, 1, scr.data_ptr(), st)
torch.cuda.synchronize()
print('ok', lib.load().sn_last_kernel())
