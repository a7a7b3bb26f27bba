"""Bring-up aid: fp32 dense-MLP forward at full size, layer by layer, repeated (determinism + accuracy)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from _workloads import KSExplicit
from pnode_b200 import sliced as sl
from pnode_b200.densemlp import DenseMlpCallbacks, recognise_relu_mlp

torch.manual_seed(0)
n, H, B = 1024, 3200, 256
for dtype in (torch.float32, torch.float64):
    func = KSExplicit(n, hidden=H, dtype=dtype).cuda()
    lins = [m for m in func.F if isinstance(m, torch.nn.Linear)]
    u = (0.5 * torch.randn(B, n, dtype=torch.float64)).to(dtype).cuda()
    x, xr = u, u.double()
    for l, lin in enumerate(lins):
        ref = torch.nn.functional.linear(xr, lin.weight.double(), lin.bias.double())
        outs = [sl.gemm(sl.slice_rows(x), sl.slice_rows(lin.weight.detach()), bias=lin.bias.detach(), relu=l < 4) for _ in range(3)]
        if l < 4:
            ref = ref.relu()
        print(dtype, "layer", l, tuple(lin.weight.shape), "err %.2e" % ((outs[0].double() - ref).abs().max() / ref.abs().max()).item(),
              "repeatable", all(torch.equal(outs[0], o) for o in outs[1:]), flush=True)
        x, xr = outs[0], ref
    meta = torch.empty(B, n, dtype=dtype, device="cuda")
    spec = recognise_relu_mlp(func, meta)
    cb = DenseMlpCallbacks(func, meta.shape, spec[0], spec[1])
    cb.begin(True, keep=True)
    ref = func(0.0, u).double()
    for keep in (False, True, True):
        o = cb.f(0.0, u.reshape(-1).clone(), keep=keep).view(B, n)
        print(dtype, "dmlp forward keep=%s err %.2e" % (keep, ((o.double() - ref).abs().max() / ref.abs().max()).item()), flush=True)
