"""Generates tests/golden/cfg4_full_{fp32,fp64}.pt: BASELINE config 4 at FULL size (SqueezeNext ODE block 1, state
[256,32,32,32], RK4, t=[1.0], Nt=1 => ONE step of h=1, train-mode BatchNorm) through the CPU ORACLE on seeded inputs.
Run once in the build container:   python tests/golden/make_cfg4_full.py
Stored (the state has 8.4 M entries): 8192 sampled entries of the final state and of lambda + their norms and sums, mu in
full, and the BatchNorm running statistics of the first / last layer after the pass."""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import OracleODEPetsc  # noqa: E402
from _workloads import OdeConvBlock  # noqa: E402


def main():
    B, C, HW, seed = 256, 32, 32, 3
    torch.set_num_threads(os.cpu_count())
    # fp32_kinkfree: BatchNorm shifts of +8 sigma keep every ReLU input away from 0, so that the fp32 comparison involves no branch
    # decisions (a unit within fp32 rounding of the kink takes a different branch in any two fp32 implementations and moves
    # lambda / mu by O(1e-3); with 23 M units per evaluation a few are expected)
    for name, dtype in (("fp64", torch.float64), ("fp32", torch.float32), ("fp32_kinkfree", torch.float32)):
        g = torch.Generator().manual_seed(seed)
        u0 = torch.randn(B, C, HW, HW, generator=g, dtype=torch.float64).to(dtype)
        gout = torch.randn(1, B, C, HW, HW, generator=g, dtype=torch.float64).to(dtype)
        t = torch.tensor([1.0], dtype=torch.float64)
        func = OdeConvBlock(C, dtype=dtype)
        if name.endswith("kinkfree"):
            with torch.no_grad():
                for m in func.modules():
                    if isinstance(m, torch.nn.BatchNorm2d):
                        m.bias.copy_(8.0 * m.weight)
        ode = OracleODEPetsc(["-ts_adapt_type", "none"])
        ode.setupTS(u0, func, step_size=1.0, method="rk4")
        t0 = time.time()
        y0 = u0.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t)
        (out * gout).sum().backward()
        print(name, "oracle pass: %.1f s" % (time.time() - t0), flush=True)
        mu = torch.cat([p.grad.reshape(-1) for p in func.parameters()])
        gi = torch.Generator().manual_seed(99)
        idx = torch.randint(0, u0.numel(), (8192,), generator=gi)
        uf, lam = out[-1].detach().reshape(-1), y0.grad.reshape(-1)
        torch.save({"B": B, "C": C, "HW": HW, "seed": seed, "index": idx, "u_sample": uf[idx].clone(), "lam_sample": lam[idx].clone(),
                    "u_norm": float(uf.double().norm()), "lam_norm": float(lam.double().norm()), "u_sum": float(uf.double().sum()),
                    "lam_sum": float(lam.double().sum()), "u_absmax": float(uf.abs().max()), "lam_absmax": float(lam.abs().max()),
                    "mu": mu.clone(), "bn1_running_mean": func.bn1.running_mean.clone(), "bn5_running_var": func.bn5.running_var.clone(),
                    "nfe": func.nfe}, os.path.join(HERE, "cfg4_full_%s.pt" % name))


if __name__ == "__main__":
    main()
