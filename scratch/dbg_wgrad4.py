import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
def run(N,C,H,W,F):
    x = torch.randn(N,H,W,C,device='cuda'); dy = torch.randn(N,H,W,F,device='cuda'); dw = torch.zeros(F,3,3,C,device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    for i in range(2): lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, N,C,H,W,F,3,3,1,1,1,1,1,st)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5): lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, N,C,H,W,F,3,3,1,1,1,1,1,st)
    e1.record(); torch.cuda.synchronize()
    print((N,C,H,W,F), '%.1f us' % (e0.elapsed_time(e1)*200))
run(512,64,16,16,128); run(512,128,8,8,256)
