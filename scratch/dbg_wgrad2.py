import sys, numpy as np, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
torch.manual_seed(0)
def run(N,C,H,W,F,k,pad, kind):
    if kind=='int':
        x = torch.randint(-4,5,(N,H,W,C),device='cuda').float()/4
        dy = torch.randint(-4,5,(N,H,W,F),device='cuda').float()/4
    else:
        x = torch.randn(N,H,W,C,device='cuda'); dy = torch.randn(N,H,W,F,device='cuda')
    if kind=='trunc':
        x = (x.view(torch.int32) & ~0x1FFF).view(torch.float32); dy=(dy.view(torch.int32) & ~0x1FFF).view(torch.float32)
    dw = torch.zeros(F,k,k,C,device='cuda'); dw2 = torch.zeros_like(dw)
    st = torch.cuda.current_stream().cuda_stream
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, N,C,H,W,F,k,k,1,pad,pad,1,1,st)
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw2.data_ptr(), None, N,C,H,W,F,k,k,1,pad,pad,1,0,st)
    torch.cuda.synchronize()
    d=(dw-dw2).abs()
    print(kind,(N,C,H,W,F,k), 'ref absmax %.3f maxdiff %.4f  meandiff %.5f' % (dw2.abs().max().item(), d.max().item(), d.mean().item()))
    idx = torch.nonzero(d > 0.25*d.max())
    print('   n big', idx.shape[0], idx[:6].tolist())
for kind in ('int','trunc','rand'):
    run(4,64,16,16,128,3,1,kind)
run(64,64,16,16,128,3,1,'rand')
