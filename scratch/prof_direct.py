import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
N,C,H,W,F=512,3,32,32,64
x=torch.randn(N,H,W,C,device='cuda'); w=torch.randn(F,3,3,C,device='cuda')*0.1; b=torch.randn(F,device='cuda')
y=torch.empty(N,H,W,F,device='cuda'); dy=torch.randn(N,H,W,F,device='cuda'); dw=torch.zeros(F,3,3,C,device='cuda'); db=torch.zeros(F,device='cuda')
st=torch.cuda.current_stream().cuda_stream
for i in range(3):
    lib.sn_conv2d_fprop(x.data_ptr(),w.data_ptr(),b.data_ptr(),y.data_ptr(),N,C,H,W,F,3,3,1,1,1,1,1,st)
    lib.sn_conv2d_wgrad(x.data_ptr(),dy.data_ptr(),dw.data_ptr(),db.data_ptr(),N,C,H,W,F,3,3,1,1,1,1,1,st)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e2=torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): lib.sn_conv2d_fprop(x.data_ptr(),w.data_ptr(),b.data_ptr(),y.data_ptr(),N,C,H,W,F,3,3,1,1,1,1,1,st)
e1.record()
for i in range(10): lib.sn_conv2d_wgrad(x.data_ptr(),dy.data_ptr(),dw.data_ptr(),db.data_ptr(),N,C,H,W,F,3,3,1,1,1,1,1,st)
e2.record(); torch.cuda.synchronize()
print('fprop us', e0.elapsed_time(e1)*100, 'wgrad us', e1.elapsed_time(e2)*100)
