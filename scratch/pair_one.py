# one tf32 + one bf16 fprop of the conv3 shape (for ncu): SN_IGEMM_PAIR=0|1
import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st = torch.cuda.current_stream().cuda_stream
import os
N, C, H, W, F = [int(v) for v in os.environ.get("SHAPE", "512,128,8,8,256").split(",")]
x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(F, 3, 3, C, device='cuda') * 0.05; b = torch.randn(F, device='cuda')
y = torch.empty(N, H, W, F, device='cuda'); xb = x.to(torch.bfloat16); wb = w.to(torch.bfloat16)
g = (N, C, H, W, F, 3, 3, 1, 1, 1)
for i in range(3):
    lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, 1, st)
    lib.sn_conv2d_fprop_bf16(xb.data_ptr(), wb.data_ptr(), b.data_ptr(), y.data_ptr(), *g, 1, st)
torch.cuda.synchronize()
