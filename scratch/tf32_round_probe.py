"""Does tcgen05 kind::tf32 truncate or round the fp32 operand to 10 mantissa bits?  (Matters for a
3xTF32 split: hi = what the hardware sees, lo = x - hi.)  1x1 convolution of one pixel."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st = torch.cuda.current_stream().cuda_stream
N, C, H, W, F = 64, 128, 8, 8, 64          # big enough for the tensor-core path (flops >= 2^22)
for frac, name in ((0.75, '1 + 0.75 ulp_tf32'), (0.25, '1 + 0.25 ulp_tf32'), (0.5, '1 + 0.5 ulp_tf32 (tie)'), (1.5, '1 + 1.5 ulp')):
    xv = np.float32(1.0 + frac * 2.0 ** -10)
    x = torch.zeros(N, H, W, C, device='cuda'); x[..., 0] = float(xv)
    w = torch.zeros(F, 1, 1, C, device='cuda'); w[:, 0, 0, 0] = 1.0
    y = torch.empty(N, H, W, F, device='cuda')
    lib.sn_conv2d_fprop(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), N, C, H, W, F, 1, 1, 1, 0, 0, 0, 1, st)
    torch.cuda.synchronize()
    got = float(y[0, 0, 0, 0])
    print('%-26s x = %.10f  ->  product with 1.0 = %.10f  (%s)' % (name, float(xv), got, 'truncated' if got < float(xv) else ('rounded up' if got > float(xv) else 'exact')))
