import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st = torch.cuda.current_stream().cuda_stream
N, C, H, W, F = 8, 3, 32, 32, 64
x = torch.randn(N, H, W, C, device='cuda'); w = torch.randn(F, 3, 3, C, device='cuda') * 0.1; b = torch.randn(F, device='cuda')
y = torch.empty(N, H, W, F, device='cuda'); dy = torch.randn(N, H, W, F, device='cuda')
dw = torch.zeros(F, 3, 3, C, device='cuda'); db = torch.zeros(F, device='cuda')
which = sys.argv[1]
if which == 'fprop':
    lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)
else:
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N, C, H, W, F, 3, 3, 1, 1, 1, 1, 1, st)
torch.cuda.synchronize()
ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
if which == 'fprop':
    print('fprop max err', (y - ref).abs().max().item())
print('ok', which)
