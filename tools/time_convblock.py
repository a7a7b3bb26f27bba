#!/usr/bin/env python
"""Time the conv-block right-hand side (pnode_convblock_forward) and its VJP (pnode_convblock_vjp) on the four CIFAR ODE-block
shapes of BASELINE config 4, CUDA events around `reps` back-to-back calls (GPU time) and wall clock (host launch time), against
the algorithmic traffic of SURVEY.md section 8d:  f = 5.5 C HW w bytes / sample, vjp = 3 x that (recompute + dgrad + wgrad).
usage: python tools/time_convblock.py [--reps 20] [--once] [--dtype f32]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from _workloads import OdeConvBlock  # noqa: E402
from pnode_b200.convblock import ConvBlockCallbacks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--once", action="store_true", help="one f and one vjp per shape, no timing (for ncu)")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--shapes", default="32x32,64x16,128x8,256x4")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--native", type=int, default=1, help="0: library convolutions + csrc/bn_relu.cu (the non-native evaluator)")
    args = ap.parse_args()
    dt = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if args.dtype == "f32" else 8
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6532.2}
    for sh in args.shapes.split(","):
        Cc, H = (int(v) for v in sh.split("x"))
        func = OdeConvBlock(Cc, dtype=dt).cuda()
        from pnode_b200.options import Options
        Options.insert_args(["-pnode_convblock_native", str(args.native)])
        cb = ConvBlockCallbacks(func, torch.Size((args.batch, Cc, H, H)))
        assert cb.native == bool(args.native)
        x = torch.randn(args.batch * Cc * H * H, dtype=dt, device="cuda")
        g = torch.randn_like(x)
        mu = torch.zeros(cb.nparams, dtype=dt, device="cuda")
        if not args.native:
            from pnode_b200.device import DeviceOps
            ops = DeviceOps(x.device, dt)

            def vjp_acc():
                vu, gp = cb.vjp(0.0, x, g)
                ops.multi_axpy(mu, gp, cb.sizes, 1.0)
        else:
            vjp_acc = lambda: cb.vjp_accumulate(0.0, x, g, mu, 1.0)
        if args.once:
            cb.f(0.0, x)
            vjp_acc()
            torch.cuda.synchronize()
            continue
        res = {"shape": [args.batch, Cc, H, H], "dtype": args.dtype}
        for name, call, mult in (("f", lambda: cb.f(0.0, x), 1.0), ("vjp", vjp_acc, 3.0)):
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.reps):
                call()
            e1.record()
            host = (time.perf_counter() - t0) / args.reps
            torch.cuda.synchronize()
            gpu = e0.elapsed_time(e1) * 1e-3 / args.reps
            bytes_alg = mult * 5.5 * Cc * H * H * w * args.batch
            flops = mult * 4.5 * Cc * Cc * H * H * args.batch
            res[name] = {"gpu_us": gpu * 1e6, "host_issue_us": host * 1e6, "algorithmic_GBps": bytes_alg / gpu / 1e9,
                         "frac_hbm": bytes_alg / gpu / 1e9 / peaks["hbm_gbs"], "tflops": flops / gpu / 1e12}
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
