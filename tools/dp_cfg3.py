#!/usr/bin/env python
"""Weak-scaling measurement of BASELINE config 3 (FFJORD CNF, D=6, H=60, dopri5 adaptive, fp32) under torchrun: 2^20
trajectories per GPU, batch sharded; per step attempt ONE scalar all-reduce (NCCL) of the weighted squared error before
accept/reject so that all ranks share the reference's single step sequence; mu all-reduced inside the adjoint kernel over
NVLink peer memory.  Prints one JSON line on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29631 tools/dp_cfg3.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from _workloads import CNFFunc, cnf_to
    from pnode import petsc_adjoint
    from pnode_b200.options import Options
    from pnode_b200.parallel import BatchComm

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    iters, warm = 10, 3
    B = 1 << 20
    Options.insert_args(["-ts_trajectory_type", "memory"])
    g = torch.Generator().manual_seed(2 + rank)
    func = cnf_to(CNFFunc(B, 6, (60,), dtype=torch.float32, seed=rank), dev)
    u0 = torch.cat((torch.randn(B, 6, generator=g).view(-1), torch.zeros(B))).to(dev)
    target = torch.randn(2, B * 7, generator=g).to(dev)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64, device=dev)
    ode = petsc_adjoint.ODEPetsc()
    if world > 1:
        ode.comm = BatchComm()
        ode.comm.enable_peer_reduce()

    def step():
        func.zero_grad(set_to_none=True)
        ode.setupTS(u0, func, step_size=0.05, method="dopri5", enable_adjoint=True)
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
        loop = ode._loop
        print(json.dumps({"config": "cfg3 FFJORD CNF D=6 H=60, 2^20 trajectories per GPU, dopri5 adaptive, fp32, batch sharded",
                          "n_gpus": world, "path": ode.path, "ms_per_pass": float(ms), "accepted_steps": loop.steps,
                          "attempts": len(loop.attempts),
                          "traj_steps_per_s": B * world * loop.steps / (float(ms) * 1e-3)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
