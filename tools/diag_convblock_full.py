import sys, copy, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from _workloads import OdeConvBlock
from _problems import rel_err
from pnode_b200.convblock import ConvBlockCallbacks
from pnode_b200.options import Options
Options.insert_args(["-pnode_convblock_native", "1"])
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
for shape in [(8, 32, 32, 32), (64, 32, 32, 32), (256, 32, 32, 32), (256, 64, 16, 16)]:
    for dtype in (torch.float64,):
        func = OdeConvBlock(shape[1], dtype=dtype).cuda()
        g = torch.Generator().manual_seed(1)
        x = torch.randn(shape, generator=g, dtype=torch.float64).to(dtype).cuda()
        w = torch.randn(shape, generator=g, dtype=torch.float64).to(dtype).cuda()
        f = copy.deepcopy(func); xr = x.clone().requires_grad_(True); out = f(0.0, xr); out.backward(w)
        cb = ConvBlockCallbacks(copy.deepcopy(func), torch.Size(shape))
        o = cb.f(0.0, x.reshape(-1)).view(shape)
        vu, gp = cb.vjp(0.0, x.reshape(-1), w.reshape(-1))
        names = [n for n, _ in f.named_parameters()]
        errs = {n: rel_err(a.view_as(p.grad), p.grad) for n, a, p in zip(names, gp, f.parameters()) if "conv" not in n or "bias" not in n}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print(shape, dtype, "f %.2e  vu %.2e" % (rel_err(o, out.detach()), rel_err(vu.view(shape), xr.grad)), worst, flush=True)
