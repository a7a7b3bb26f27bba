#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ by importing the REFERENCE's own Python model code in this container
(/root/reference is not available on the GPU box, so the vectors are committed).  PETSc is absent, so what can be pinned this way
is the part of the path that is pure Python/torch in the reference: the right-hand-side modules themselves.

  convblock_ref_fp64.pt -- BASELINE config 4: `BasicBlock2` of /root/reference/examples-pnode/models/sqnxt_PETSc.py:70-121 (the ODE
      block's `func`), fp64, train mode: a seeded input x, cotangent w, the module's state_dict, and f(x), (df/dx)^T w,
      (df/dp)^T w, the BatchNorm buffers after ONE forward -- from the reference module's forward and torch autograd on the CPU.

  cnf_ref_fp64.pt -- BASELINE config 3: FlattenFunc(ODEfunc(ODEnet(hidden (60,), D=6, concatsquash, softplus))) of
      /root/reference/ffjord-pnode/lib/layers/{odefunc.py:97-385, cnf.py:123-150, diffeq_layers/basic.py:76-86} with a fixed
      Hutchinson probe: the flattened state y = cat(z, logp), f(t, y) at two times, and (df/dy)^T w, (df/dp)^T w (second-order
      autograd through the trace estimator), fp64 on the CPU.  The reference package imports petsc4py at module level; this
      repo's petsc4py shim (options database only) satisfies that import -- none of the code exercised here touches PETSc.

  ks_ref_fp64.pt -- BASELINE config 5: ODEFuncIM(fixed_linear=True, dx=22/64) and ODEFuncEX(64, 48) of
      /root/reference/examples-sinode/KS/models/imex.py:6-70 (the reference builds the stencil with device="cuda:0"; the
      generator redirects that one torch.tensor call to the CPU), fp64: f_IM, f_EX and their VJPs at a seeded state.

usage (in the build container):  python tests/golden/make_golden.py
"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/examples-pnode/models/sqnxt_PETSc.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_sqnxt", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(1234)
    torch.set_num_threads(1)
    cases = {}
    for name, (N, C, H, W) in {"block_c16_8x8": (4, 16, 8, 8), "block_c32_4x16": (3, 32, 4, 16)}.items():
        func = mod.BasicBlock2(C).double()
        func.train()
        g = torch.Generator().manual_seed(C)
        with torch.no_grad():
            for m in func.modules():  # non-trivial BatchNorm affine parameters (the reference initialises them to 1 / 0)
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.weight.copy_(torch.rand(m.num_features, generator=g, dtype=torch.float64) + 0.5)
                    m.bias.copy_(0.3 * torch.randn(m.num_features, generator=g, dtype=torch.float64))
        state = {k: v.clone() for k, v in func.state_dict().items()}
        x = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
        w = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
        xr = x.clone().requires_grad_(True)
        out = func(0.0, xr)
        out.backward(w)
        cases[name] = {"shape": (N, C, H, W), "state": state, "x": x, "w": w, "f": out.detach().clone(), "vjp_x": xr.grad.clone(),
                       "vjp_p": {n: p.grad.clone() for n, p in func.named_parameters()},
                       "buffers_after": {k: v.clone() for k, v in func.state_dict().items() if "running" in k or "tracked" in k}}
    torch.save({"source": REF + ":70-121 (BasicBlock2)", "torch": str(torch.__version__), "cases": cases},
               os.path.join(HERE, "convblock_ref_fp64.pt"))
    print("wrote", os.path.join(HERE, "convblock_ref_fp64.pt"))
    make_cnf()
    make_ks()


def make_ks():
    spec = importlib.util.spec_from_file_location("ref_imex", "/root/reference/examples-sinode/KS/models/imex.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(99)
    N, H, B = 64, 48, 4
    dx = 22.0 / N
    orig = torch.tensor
    try:  # ODEFuncIM places its fixed stencil on "cuda:0" (imex.py:33): same formula, on the CPU
        torch.tensor = lambda *a, **k: orig(*a, **{kk: v for kk, v in k.items() if kk != "device"})
        f_im = mod.ODEFuncIM(fixed_linear=True, dx=dx).double()
    finally:
        torch.tensor = orig
    f_ex = mod.ODEFuncEX(input_size=N, hidden=H).double()
    g = torch.Generator().manual_seed(5)
    y = 0.5 * torch.randn(B, N, generator=g, dtype=torch.float64)
    w = torch.randn(B, N, generator=g, dtype=torch.float64)
    out = {"dx": dx, "N": N, "H": H, "y": y, "w": w, "state_im": {k: v.clone() for k, v in f_im.state_dict().items()},
           "state_ex": {k: v.clone() for k, v in f_ex.state_dict().items()}}
    for name, f in (("im", f_im), ("ex", f_ex)):
        yr = y.clone().requires_grad_(True)
        val = f(0.1, yr)
        ps = [p for p in f.parameters() if p.requires_grad]
        grads = torch.autograd.grad(val, [yr] + ps, w, allow_unused=True)
        out[name] = {"f": val.detach().clone(), "vjp_y": grads[0].clone(),
                     "vjp_p": {n: gg.clone() for (n, p), gg in zip([(n, p) for n, p in f.named_parameters() if p.requires_grad],
                                                                  grads[1:])}}
    torch.save(out, os.path.join(HERE, "ks_ref_fp64.pt"))
    print("wrote", os.path.join(HERE, "ks_ref_fp64.pt"))


def make_cnf():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))  # petsc4py / pnode shims of this repo
    sys.path.insert(0, "/root/reference/ffjord-pnode")
    import lib.layers.cnf as ref_cnf
    import lib.layers.odefunc as ref_odefunc

    torch.manual_seed(4321)
    B, D, H = 5, 6, 60
    net = ref_odefunc.ODEnet(hidden_dims=(H,), input_shape=(D,), strides=None, conv=False, layer_type="concatsquash",
                             nonlinearity="softplus").double()
    odefunc = ref_odefunc.ODEfunc(diffeq=net, divergence_fn="approximate", residual=False, rademacher=False).double()
    g = torch.Generator().manual_seed(7)
    e = torch.randn(B, D, generator=g, dtype=torch.float64)
    odefunc.before_odeint(e=e)
    z = torch.randn(B, D, generator=g, dtype=torch.float64)
    logp = torch.zeros(B, 1, dtype=torch.float64)
    func = ref_cnf.FlattenFunc(odefunc, (z, logp))
    y = torch.cat((z.reshape(-1), logp.reshape(-1)))
    w = torch.randn(y.numel(), generator=g, dtype=torch.float64)
    out = {}
    for t in (0.0, 0.37):
        yr = y.clone().requires_grad_(True)
        f = func(t, yr)
        grads = torch.autograd.grad(f, [yr] + list(func.parameters()), w, allow_unused=True)
        out["t=%g" % t] = {"t": t, "f": f.detach().clone(), "vjp_y": grads[0].clone(),
                           "vjp_p": {n: (gg.clone() if gg is not None else None)
                                     for (n, _), gg in zip(func.named_parameters(), grads[1:])}}
    torch.save({"source": "/root/reference/ffjord-pnode/lib/layers (FlattenFunc(ODEfunc(ODEnet)))", "B": B, "D": D, "H": H,
                "state": {k: v.clone() for k, v in net.state_dict().items()}, "e": e, "y": y, "w": w, "evals": out,
                "param_names": [n for n, _ in func.named_parameters()]}, os.path.join(HERE, "cnf_ref_fp64.pt"))
    print("wrote", os.path.join(HERE, "cnf_ref_fp64.pt"))


if __name__ == "__main__":
    main()
