import sys, numpy as np, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
torch.manual_seed(0)
def run(N,C,H,W,F,k,pad):
    x = torch.randint(-4,5,(N,H,W,C),device='cuda').float()/4
    dy = torch.randint(-4,5,(N,H,W,F),device='cuda').float()/4
    dw = torch.zeros(F,k,k,C,device='cuda'); db = torch.zeros(F,device='cuda')
    dw2 = torch.zeros_like(dw); db2=torch.zeros_like(db)
    st = torch.cuda.current_stream().cuda_stream
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), N,C,H,W,F,k,k,1,pad,pad,1,1,st)
    lib.sn_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), dw2.data_ptr(), db2.data_ptr(), N,C,H,W,F,k,k,1,pad,pad,1,0,st)
    torch.cuda.synchronize()
    print((N,C,H,W,F,k), 'tc absmax', dw.abs().max().item(), 'ref absmax', dw2.abs().max().item(), 'maxdiff', (dw-dw2).abs().max().item(), 'db diff', (db-db2).abs().max().item())
    if (dw-dw2).abs().max().item() > 1e-3:
        d = (dw-dw2).abs()
        print('  nonzero frac tc', (dw!=0).float().mean().item(), ' per-tap maxdiff', d.amax(dim=(0,3)).flatten().tolist())
        print('  per f-quarter maxdiff', [d[i*32:(i+1)*32].max().item() for i in range(F//32)])
        print('  per c-chunk maxdiff', [d[..., i*32:(i+1)*32].max().item() for i in range(C//32)])
run(4,64,16,16,128,3,1)
run(8,64,16,16,128,1,0)
run(64,32,1,1,128,1,0)
