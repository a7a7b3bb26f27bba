#!/usr/bin/env python
"""Config 2 (2^20 trajectories, RK4, 9 steps) with stage checkpoints vs -ts_trajectory_solution_only 1 on the fused sweeps:
milliseconds per sweep and bytes of checkpoints kept in HBM.  Run under gpurun."""
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

for dtype in (torch.float64, torch.float32):
    u0, t, gout = spiral_inputs(1 << 20, dtype=dtype)
    u0, t, gout = u0.cuda(), t.cuda(), gout.cuda()
    for argv in ([], ["-ts_trajectory_solution_only", "1"]):
        Options.clear_all()
        Options.insert_args(["-ts_adapt_type", "none"] + argv)
        func = copy.deepcopy(SpiralFunc(dtype=dtype)).cuda()
        ode = petsc_adjoint.ODEPetsc()
        ode.setupTS(u0, func, step_size=0.025, method="rk4", enable_adjoint=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tf = ta = 0.0
        reps = 6
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        for i in range(reps + 2):
            func.zero_grad(set_to_none=True)
            y0 = u0.clone().requires_grad_(True)
            torch.cuda.synchronize()
            ev[0].record()
            pred = ode.odeint_adjoint(y0, t)
            ev[1].record()
            pred.backward(gout)
            ev[2].record()
            torch.cuda.synchronize()
            if i >= 2:
                tf += ev[0].elapsed_time(ev[1])
                ta += ev[1].elapsed_time(ev[2])
        print(json.dumps({"dtype": str(dtype), "argv": argv, "path": ode.path, "fwd_ms": tf / reps, "adj_ms": ta / reps,
                          "peak_extra_mb": (torch.cuda.max_memory_allocated() - base) / 1e6}))
