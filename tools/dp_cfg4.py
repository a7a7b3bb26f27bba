#!/usr/bin/env python
"""Weak-scaling measurement of BASELINE config 4 (CIFAR conv ODE block 1, 256 images per GPU, RK4 + adjoint, fp32) under
torchrun: batch sharded over the ranks, BatchNorm statistics of the GLOBAL batch exchanged inside the kernels over NVLink
(csrc/conv_block.cu: stats_exchange), mu all-reduced with NCCL.  Prints one JSON line on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29621 tools/dp_cfg4.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from _workloads import OdeConvBlock
    from pnode import petsc_adjoint
    from pnode_b200.options import Options
    from pnode_b200.parallel import BatchComm

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    iters, warm = 20, 5
    C, HW, B = 32, 32, 256
    Options.insert_args(["-ts_adapt_type", "none"])
    g = torch.Generator().manual_seed(3 + rank)
    u0 = torch.randn(B, C, HW, HW, generator=g).to(dev)
    target = torch.randn(1, B, C, HW, HW, generator=g).to(dev)
    t = torch.tensor([1.0], dtype=torch.float64, device=dev)
    func = OdeConvBlock(C).to(dev)
    ode = petsc_adjoint.ODEPetsc()
    if world > 1:
        ode.comm = BatchComm()
    ode.setupTS(u0, func, step_size=1.0, method="rk4", enable_adjoint=True)

    def step():
        func.zero_grad(set_to_none=True)
        ode.setupTS(u0, func, step_size=1.0, method="rk4", enable_adjoint=True)
        loss = torch.mean(torch.abs(ode.odeint_adjoint(u0, t) - target))
        loss.backward()

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "cfg4 block 1 [256,32,32,32] per GPU, RK4 fwd+adjoint fp32, batch sharded, global BatchNorm "
                                    "statistics exchanged in-kernel over NVLink", "n_gpus": world, "ms_per_pass": float(ms),
                          "traj_steps_per_s": B * world / (float(ms) * 1e-3), "native": bool(ode._cb_im.native),
                          "collectives_per_pass": (ode.comm.collectives // (iters + warm)) if world > 1 else 0}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
