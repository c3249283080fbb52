import sys, numpy as np, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
torch.manual_seed(0)
def run(N,C,H,W,F,k,pad):
    x = torch.randint(-4,5,(N,H,W,C),device='cuda').float()/4
    OH, OW = H+2*pad-k+1, W+2*pad-k+1
    dy = torch.randint(-4,5,(N,OH,OW,F),device='cuda').float()/4
    dw = torch.zeros(F,k,k,C,device='cuda'); dw2 = torch.zeros_like(dw)
    st = torch.cuda.current_stream().cuda_stream
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, N,C,H,W,F,k,k,1,pad,pad,1,1,st)
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw2.data_ptr(), None, N,C,H,W,F,k,k,1,pad,pad,1,0,st)
    torch.cuda.synchronize()
    d=(dw-dw2).abs()
    print((N,C,H,W,F,k,pad), 'ref absmax %.3f maxdiff %.4f' % (dw2.abs().max().item(), d.max().item()), 'per-tap', [round(v,2) for v in d.amax(dim=(0,3)).flatten().tolist()])
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, N,C,H,W,F,k,k,1,pad,pad,1,1,st)
    e1.record(); torch.cuda.synchronize()
    fl = 2.0*N*OH*OW*F*C*k*k
    print('   %.1f us  %.1f TF/s' % (e0.elapsed_time(e1)*200, fl/(e0.elapsed_time(e1)*0.2e-3)/1e12))
run(512,64,16,16,128,3,1)
run(512,128,8,8,256,3,1)
run(512,256,4,4,256,3,1)
run(64,64,32,32,64,3,1)
run(33,32,8,8,48,3,1)
run(16,64,18,18,64,3,0)
run(16,64,16,16,32,5,2)
run(40,128,4,4,128,3,1)
run(7,64,8,8,130,3,1)
