import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st=torch.cuda.current_stream().cuda_stream
def run(N,H,W,C):
    x=torch.randn(N,H,W,C,device='cuda'); g=torch.ones(C,device='cuda'); b=torch.zeros(C,device='cuda')
    y=torch.empty(N,H//2,W//2,C,device='cuda'); code=torch.empty(N*(H//2)*(W//2)*C,dtype=torch.uint8,device='cuda')
    sm=torch.empty(C,device='cuda'); si=torch.empty(C,device='cuda'); mm=torch.zeros(C,device='cuda'); mv=torch.ones(C,device='cuda')
    scr=torch.zeros(20*C+2,dtype=torch.float64,device='cuda'); dyp=torch.randn_like(y); dx=torch.empty_like(x); dg=torch.zeros(C,device='cuda'); db=torch.zeros(C,device='cuda')
    def f(): lib.sn_bn_pool_fwd_train(x.data_ptr(),g.data_ptr(),b.data_ptr(),y.data_ptr(),code.data_ptr(),sm.data_ptr(),si.data_ptr(),mm.data_ptr(),mv.data_ptr(),scr.data_ptr(),N,H,W,C,1e-6,0.99,0,st)
    def bw(): lib.sn_bn_pool_bwd(dyp.data_ptr(),code.data_ptr(),x.data_ptr(),g.data_ptr(),sm.data_ptr(),si.data_ptr(),dx.data_ptr(),dg.data_ptr(),db.data_ptr(),scr.data_ptr(),N,H,W,C,0,1,st)
    for fn,name in ((f,'fwd'),(bw,'bwd')):
        for i in range(3): fn()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print((N,H,W,C), name, '%.1f us' % (e0.elapsed_time(e1)*100))
for s in [(512,32,32,64),(512,16,16,128),(512,8,8,256),(512,4,4,256)]: run(*s)
