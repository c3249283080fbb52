import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from shinnosuke_b200._lib import lib
from oracle import shinnosuke_oracle as O
from conftest import rel_err
torch.cuda.set_device(0)
st = torch.cuda.current_stream().cuda_stream
def run(N, C, H, W, F, k, pad, act, ybf16, stats=True, time_it=False):
    r = np.random.default_rng(N + C + F)
    x = r.standard_normal((N, C, H, W)).astype(np.float32)
    w = (0.3 * r.standard_normal((F, C, k, k))).astype(np.float32)
    b = r.standard_normal(F).astype(np.float32)
    assert lib.load().sn_conv2d_stem_supported(N, C, H, W, F, k, k, 1, pad, pad), 'unsupported'
    xd = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1))).cuda()
    wd = torch.from_numpy(np.ascontiguousarray(w.transpose(0, 2, 3, 1))).cuda()
    bd = torch.from_numpy(b).cuda()
    x4 = torch.empty(N * H * W * 4, device='cuda')
    oh, ow = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    y = torch.full((N * oh * ow * F,), float('nan'), device='cuda', dtype=torch.bfloat16 if ybf16 else torch.float32)
    scr = torch.zeros(28 * F + 2, device='cuda', dtype=torch.float64)
    lib.sn_conv2d_stem_prepare(xd.data_ptr(), x4.data_ptr(), N, C, H, W, st)
    lib.sn_conv2d_stem_fprop(x4.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), N, C, H, W, F, k, k, 1, pad, pad, act, int(ybf16),
                             scr.data_ptr() if stats else None, st)
    torch.cuda.synchronize()
    yh = y.float().cpu().numpy().reshape(N, oh, ow, F).transpose(0, 3, 1, 2)
    nn = min(N, 16)
    ref, _ = O.conv2d_forward(x[:nn].astype(np.float64), w.astype(np.float64), b.reshape(1, F).astype(np.float64), 1, (pad, pad), 'relu' if act else 'linear')
    e = rel_err(yh[:nn], ref)
    e2 = rel_err(yh[-2:], O.conv2d_forward(x[-2:].astype(np.float64), w.astype(np.float64), b.reshape(1, F).astype(np.float64), 1, (pad, pad), 'relu' if act else 'linear')[0])
    msg = 'N%d C%d %dx%d F%d k%d pad%d act%d bf16=%d: rel err %.2e (last images %.2e) finite=%s' % (N, C, H, W, F, k, pad, act, ybf16, e, e2, np.isfinite(yh).all())
    if stats:
        s = scr.cpu().numpy()[:16 * F].reshape(8, 2, F).sum(0)
        yy = yh.astype(np.float64)
        msg += ' stats err %.1e %.1e' % (rel_err(s[0], yy.sum(axis=(0, 2, 3))), rel_err(s[1], (yy * yy).sum(axis=(0, 2, 3))))
    if time_it:
        scr.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(12)]
        for i in range(11):
            ev[i].record()
            lib.sn_conv2d_stem_fprop(x4.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), N, C, H, W, F, k, k, 1, pad, pad, act, int(ybf16),
                                     scr.data_ptr() if stats else None, st)
        ev[11].record(); torch.cuda.synchronize()
        msg += ' | %.1f us' % (np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(1, 10)]) * 1e3)
    print(msg, flush=True)
def run_wgrad(N, C, H, W, F, k, pad, time_it=False):
    r = np.random.default_rng(N + C + F + 1)
    x = r.standard_normal((N, C, H, W)).astype(np.float32)
    oh, ow = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    dy = r.standard_normal((N, F, oh, ow)).astype(np.float32)
    xd = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1))).cuda()
    dyd = torch.from_numpy(np.ascontiguousarray(dy.transpose(0, 2, 3, 1))).cuda()
    x4 = torch.empty(N * H * W * 4, device='cuda')
    dw = torch.zeros(F * k * k * C, device='cuda'); db = torch.zeros(F, device='cuda')
    lib.sn_conv2d_stem_prepare(xd.data_ptr(), x4.data_ptr(), N, C, H, W, st)
    lib.sn_conv2d_stem_wgrad(x4.data_ptr(), dyd.data_ptr(), dw.data_ptr(), db.data_ptr(), N, C, H, W, F, k, k, 1, pad, pad, st)
    torch.cuda.synchronize()
    dwh = dw.cpu().numpy().reshape(F, k, k, C).transpose(0, 3, 1, 2)
    w0 = np.zeros((F, C, k, k)); b0 = np.zeros((1, F))
    dw_ref = np.zeros((F, C, k, k)); db_ref = np.zeros(F)
    for a in range(0, N, 64):
        _, cache = O.conv2d_forward(x[a:a + 64].astype(np.float64), w0, b0, 1, (pad, pad), 'linear')
        _, dwc, dbc = O.conv2d_backward(dy[a:a + 64].astype(np.float64), cache, need_dx=False)
        dw_ref += dwc; db_ref += np.asarray(dbc).reshape(-1)
    msg = 'wgrad N%d C%d %dx%d F%d k%d pad%d: dw err %.2e db err %.2e' % (N, C, H, W, F, k, pad, rel_err(dwh, dw_ref), rel_err(db.cpu().numpy(), db_ref))
    if time_it:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(12)]
        for i in range(11):
            ev[i].record()
            lib.sn_conv2d_stem_wgrad(x4.data_ptr(), dyd.data_ptr(), dw.data_ptr(), db.data_ptr(), N, C, H, W, F, k, k, 1, pad, pad, st)
        ev[11].record(); torch.cuda.synchronize()
        msg += ' | %.1f us' % (np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(1, 10)]) * 1e3)
    print(msg, flush=True)

if os.environ.get('SN_STEM_DBG'):
    run(512, 3, 32, 32, 64, 3, 1, 1, True, stats=False, time_it=True)
    run(512, 3, 32, 32, 64, 3, 1, 1, True, stats=True, time_it=True)
    sys.exit(0)
mode = sys.argv[1] if len(sys.argv) > 1 else 'fprop'
if mode == 'fprop':
    run(8, 3, 32, 32, 64, 3, 1, 1, False)
    run(8, 3, 32, 32, 64, 3, 1, 1, True)
    run(6, 3, 30, 32, 32, 3, 0, 0, False)
    run(5, 1, 28, 28, 32, 2, 0, 1, False)
    run(3, 4, 16, 16, 128, 3, 1, 0, True)
    run(7, 3, 16, 16, 64, 3, 1, 1, True)
    run(4, 3, 64, 64, 32, 3, 1, 0, False)
    run(512, 3, 32, 32, 64, 3, 1, 1, True, time_it=True)
    run(512, 3, 32, 32, 64, 3, 1, 1, True, stats=False, time_it=True)
    run(512, 3, 32, 32, 64, 3, 1, 1, False, time_it=True)
elif mode == 'wgrad_small':
    run_wgrad(2, 3, 32, 32, 64, 3, 1)
else:
    run_wgrad(8, 3, 32, 32, 64, 3, 1)
    run_wgrad(6, 3, 30, 32, 32, 3, 0)
    run_wgrad(5, 1, 28, 28, 32, 2, 0)
    run_wgrad(3, 4, 16, 16, 128, 3, 1)
    run_wgrad(512, 3, 32, 32, 64, 3, 1, time_it=True)
