#!/usr/bin/env python
"""Can one fwd+adjoint pass of a bench_configs.py workload be captured in a CUDA graph, and what does a replay cost?
    python tools/graph_probe.py --configs 1 4 4c 5"""
import argparse
import copy
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench_configs as bc  # noqa: E402


def timed(fn, iters):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="*", default=["1", "4", "5"])
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    from pnode import petsc_adjoint
    from pnode_b200.options import Options

    table = bc.config_table()
    dev = torch.device("cuda:0")
    for key in args.configs:
        name, build = table[key]
        spec = build()
        Options.clear_all()
        Options.insert_args(spec["argv"])
        to_dev = spec.get("to_dev", lambda f, d: f.to(d))
        funcs = [to_dev(copy.deepcopy(f), dev) for f in spec["funcs"]]
        u0, t, target = spec["u0"].to(dev), spec["t"].to(dev), spec["target"].to(dev)
        step, ode = bc._make_step(lambda: petsc_adjoint.ODEPetsc(), funcs, u0, t, target, spec["kw"], spec["step"], dev,
                                  spec.get("each_call_setup", False))
        for _ in range(3):
            step()
        eager = timed(step, args.iters)
        print("%-12s eager %.4f ms/pass (%s)" % (name, eager, ode.path), flush=True)
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            rep = timed(g.replay, args.iters)
            print("%-12s graph %.4f ms/pass" % (name, rep), flush=True)
        except Exception:
            traceback.print_exc(limit=6)
            print("%-12s capture failed" % name, flush=True)
            torch.cuda.synchronize()


if __name__ == "__main__":
    main()
