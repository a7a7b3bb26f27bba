"""Bring-up aid: per-quantity errors of the tensor-core conv evaluator against the fp64 torch module."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_convmma import _callbacks, _setup
from _problems import rel_err

for shape in [(8, 128, 8, 8), (2, 32, 4, 4)]:
    func, x, w, out_r, vu_r, gp_r, ref = _setup(shape, seed=shape[0])
    mine = copy.deepcopy(func)
    cb = _callbacks(mine, shape)
    cb.begin(True)
    out = cb.f(0.0, x.reshape(-1)).view(shape)
    vu, gp = cb.vjp(0.0, x.reshape(-1), w.reshape(-1))
    print(shape, "f %.2e  vu %.2e" % (rel_err(out, out_r), rel_err(vu.view(shape), vu_r)))
    for (n, _), a, b in zip(mine.named_parameters(), gp, gp_r):
        print("   %-14s %.2e  (max ref %.2e)" % (n, rel_err(a.view_as(b), b) if float(b.abs().max()) > 0 else -1, float(b.abs().max())))
    # hidden-layer check: the same block in fp32 torch (IEEE) -- how far is plain fp32 from the fp64 reference?
    torch.backends.cudnn.allow_tf32 = False
    f32 = copy.deepcopy(func)
    xr = x.clone().requires_grad_(True)
    o = f32(0.0, xr); o.backward(w)
    print("   torch fp32 itself: f %.2e vu %.2e" % (rel_err(o, out_r), rel_err(xr.grad, vu_r)))
    for (n, p), b in zip(f32.named_parameters(), gp_r):
        if float(b.abs().max()) > 0:
            print("      %-14s %.2e" % (n, rel_err(p.grad, b)))
