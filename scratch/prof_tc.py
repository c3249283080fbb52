import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st=torch.cuda.current_stream().cuda_stream
def conv(N,C,H,W,F):
    x=torch.randn(N,H,W,C,device='cuda'); w=torch.randn(F,3,3,C,device='cuda')*0.05; b=torch.randn(F,device='cuda')
    y=torch.empty(N,H,W,F,device='cuda'); dy=torch.randn(N,H,W,F,device='cuda'); dw=torch.zeros(F,3,3,C,device='cuda'); db=torch.zeros(F,device='cuda')
    dx=torch.empty_like(x); ws=torch.empty(F*9*C,device='cuda')
    for i in range(2):
        lib.sn_conv2d_fprop(x.data_ptr(),w.data_ptr(),b.data_ptr(),y.data_ptr(),N,C,H,W,F,3,3,1,1,1,1,1,st)
        lib.sn_conv2d_wgrad(x.data_ptr(),dy.data_ptr(),dw.data_ptr(),db.data_ptr(),N,C,H,W,F,3,3,1,1,1,1,1,st)
        lib.sn_conv2d_dgrad(dy.data_ptr(),w.data_ptr(),dx.data_ptr(),N,C,H,W,F,3,3,1,1,1,0,1,ws.data_ptr(),st)
    torch.cuda.synchronize()
conv(512,64,16,16,128)
conv(512,128,8,8,256)
