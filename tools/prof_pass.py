#!/usr/bin/env python
"""One fwd+adjoint pass of a bench_configs.py workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`):
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python tools/prof_pass.py --config 5 [--opt -pnode_fused 0]"""
import argparse
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench_configs as bc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="5")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--opt", nargs="*", default=[])
    ap.add_argument("--cprofile", type=int, default=0, help="N > 0: cProfile N passes (host time) instead of the ncu window")
    args = ap.parse_args()
    from pnode import petsc_adjoint
    from pnode_b200.options import Options

    table = bc.config_table()
    name, build = table[args.config]
    spec = build()
    Options.clear_all()
    Options.insert_args(spec["argv"] + list(args.opt))
    dev = torch.device("cuda:0")
    to_dev = spec.get("to_dev", lambda f, d: f.to(d))
    funcs = [to_dev(copy.deepcopy(f), dev) for f in spec["funcs"]]
    u0, t, target = spec["u0"].to(dev), spec["t"].to(dev), spec["target"].to(dev)
    step, ode = bc._make_step(lambda: petsc_adjoint.ODEPetsc(), funcs, u0, t, target, spec["kw"], spec["step"], dev,
                              spec.get("each_call_setup", False))
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if args.cprofile:
        import cProfile
        import pstats
        import time

        t0 = time.perf_counter()
        for _ in range(args.cprofile):
            step()
        torch.cuda.synchronize()
        print("wall ms/pass (unprofiled): %.4f" % ((time.perf_counter() - t0) * 1e3 / args.cprofile))
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(args.cprofile):
            step()
        torch.cuda.synchronize()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(45)
        st.sort_stats("tottime").print_stats(30)
        return
    torch.cuda.profiler.start()
    for _ in range(args.passes):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", name, ode.path)


if __name__ == "__main__":
    main()
