import sys, torch
sys.path.insert(0, '.')
from shinnosuke_b200._lib import lib
st = torch.cuda.current_stream().cuda_stream
def t(fn, n=20):
    for i in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
flush = torch.empty(160 << 20, dtype=torch.uint8, device='cuda')
def run(N, H, W, C):
    x = torch.randn(N, H, W, C, device='cuda'); yp = torch.randn(N, H // 2, W // 2, C, device='cuda'); dyp = torch.randn_like(yp)
    code = torch.randint(1, 15, (N * (H // 2) * (W // 2) * C,), dtype=torch.uint8, device='cuda')
    g = torch.ones(C, device='cuda'); b = torch.zeros(C, device='cuda'); mean = torch.zeros(C, device='cuda'); inv = torch.ones(C, device='cuda')
    dg = torch.zeros(C, device='cuda'); db = torch.zeros(C, device='cuda'); dx = torch.empty_like(x); cs = torch.zeros(C, device='cuda')
    scr = torch.zeros(52 * C + 2, dtype=torch.float64, device='cuda')
    a = (dyp.data_ptr(), code.data_ptr(), x.data_ptr(), yp.data_ptr(), g.data_ptr(), b.data_ptr(), mean.data_ptr(), inv.data_ptr())
    def local():
        lib.sn_bn_pool_bwd_local(*a, dg.data_ptr(), db.data_ptr(), scr.data_ptr(), N, H, W, C, 0, torch.cuda.current_stream().cuda_stream)
    def apply():
        lib.sn_bn_pool_bwd_apply(dyp.data_ptr(), code.data_ptr(), x.data_ptr(), dx.data_ptr(), None, cs.data_ptr(), scr.data_ptr(), N, H, W, C, 0, 1, torch.cuda.current_stream().cuda_stream)
    def apply_nocs():
        lib.sn_bn_pool_bwd_apply(dyp.data_ptr(), code.data_ptr(), x.data_ptr(), dx.data_ptr(), None, None, scr.data_ptr(), N, H, W, C, 0, 1, torch.cuda.current_stream().cuda_stream)
    def memset_only():
        lib.sn_memset(scr.data_ptr(), 0, 28 * C * 8, torch.cuda.current_stream().cuda_stream)
    def fin():
        lib.sn_batchnorm_finalize_bwd(g.data_ptr(), mean.data_ptr(), inv.data_ptr(), scr.data_ptr(), N * H * W, C, torch.cuda.current_stream().cuda_stream)
    def graphed(fn, reps=20):
        fn(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(reps): fn()
        for _ in range(3): gr.replay()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1000
    print('bn_pool_bwd %s in-graph (warm L2): memset+reduce %.1f us, memset alone %.1f, finalize kernel %.1f us, apply %.1f us, apply without colsum %.1f us' % (str((N, H, W, C)), graphed(local), graphed(memset_only), graphed(fin), graphed(apply), graphed(apply_nocs)))
run(512, 32, 32, 64)
run(512, 16, 16, 128)
run(512, 8, 8, 256)
run(512, 4, 4, 256)
