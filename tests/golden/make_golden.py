#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/ by importing the REFERENCE's own Python model code in this container
(/root/reference is not available on the GPU box, so the vectors are committed).  PETSc is absent, so what can be pinned this way
is the part of the path that is pure Python/torch in the reference: the right-hand-side modules themselves.

  convblock_ref_fp64.pt -- BASELINE config 4: `BasicBlock2` of /root/reference/examples-pnode/models/sqnxt_PETSc.py:70-121 (the ODE
      block's `func`), fp64, train mode: a seeded input x, cotangent w, the module's state_dict, and f(x), (df/dx)^T w,
      (df/dp)^T w, the BatchNorm buffers after ONE forward -- from the reference module's forward and torch autograd on the CPU.

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


if __name__ == "__main__":
    main()
