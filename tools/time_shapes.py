#!/usr/bin/env python
"""Tiny-MLP right-hand sides beyond the spiral shape: fused sweeps (csrc/mlp_rk.cu PNODE_FOR_SHAPES, narrower layers
zero-padded) vs the generic path (-pnode_fused 0), RK4, 9 steps, 2^16 trajectories.  Run under gpurun."""
import copy
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from _problems import SpiralFunc, spiral_inputs  # noqa: E402

from pnode import petsc_adjoint  # noqa: E402
from pnode_b200.options import Options  # noqa: E402

B = 1 << 16
for dtype in (torch.float64, torch.float32):
    for dim, hidden in ((2, 50), (2, 100), (2, 64), (3, 50), (4, 50), (4, 32), (1, 50)):
        u0, t, gout = spiral_inputs(B, dtype=dtype, dim=dim)
        u0, t, gout = u0.cuda(), t.cuda(), gout.cuda()
        row = {"dtype": str(dtype).split(".")[1], "dim": dim, "hidden": hidden, "trajectories": B}
        for label, argv in (("fused", []), ("generic", ["-pnode_fused", "0"])):
            Options.clear_all()
            Options.insert_args(["-ts_adapt_type", "none"] + argv)
            func = copy.deepcopy(SpiralFunc(dtype=dtype, hidden=hidden, dim=dim, bias_std=0.1)).cuda()
            ode = petsc_adjoint.ODEPetsc()
            ode.setupTS(u0, func, step_size=0.025, method="rk4", enable_adjoint=True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            tot, reps = 0.0, 5
            for i in range(reps + 2):
                func.zero_grad(set_to_none=True)
                y0 = u0.clone().requires_grad_(True)
                torch.cuda.synchronize()
                ev[0].record()
                ode.odeint_adjoint(y0, t).backward(gout)
                ev[1].record()
                torch.cuda.synchronize()
                if i >= 2:
                    tot += ev[0].elapsed_time(ev[1])
            row[label + "_ms"] = tot / reps
            row[label + "_path"] = ode.path
        row["speedup"] = row["generic_ms"] / row["fused_ms"]
        print(json.dumps(row), flush=True)
