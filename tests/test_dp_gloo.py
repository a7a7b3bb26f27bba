"""N>1 host logic on CPU: world_size-2 `gloo` run of the batch-sharded path (tests/_fake_ops.py stands in for the device
kernels).  Checks the two exchanges of SURVEY.md section 8e: (1) mu is all-reduced so every rank ends with the
full-batch parameter gradient, (2) adaptive runs all-reduce the squared weighted error per attempt so every rank takes
the single-process step sequence (global N in the norm)."""
import copy
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, ragged, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pnode_b200.petsc_adjoint as pa
        from _fake_ops import FakeOps
        from _problems import TimeMLP
        from pnode_b200.options import Options
        from pnode_b200.parallel import BatchComm, shard_batch

        pa.DeviceOps = FakeOps
        pa._check_device = lambda t, what: None
        Options.clear_all()
        Options.insert_args(["-ts_rtol", "1e-6", "-ts_atol", "1e-6"])
        torch.manual_seed(0)
        B = 51 if ragged else 50
        g = torch.Generator().manual_seed(5)
        u0 = torch.randn(B, 6, generator=g, dtype=torch.float64)
        gout = torch.randn(3, B, 6, generator=g, dtype=torch.float64)
        t = torch.tensor([0.0, 0.4, 1.0], dtype=torch.float64)
        func = TimeMLP(d=6, hidden=16)

        def run(u, go, comm):
            f = copy.deepcopy(func)
            ode = pa.ODEPetsc()
            ode.comm = comm
            ode.setupTS(u, f, step_size=0.3, method="dopri5", enable_adjoint=True)
            y0 = u.clone().requires_grad_(True)
            out = ode.odeint_adjoint(y0, t)
            (out * go).sum().backward()
            return out.detach(), y0.grad, [p.grad.clone() for p in f.parameters()], ode._loop.attempts

        full = run(u0, gout, None)
        comm = BatchComm()
        mine = run(shard_batch(u0, rank, world).contiguous(), shard_batch(gout, rank, world, dim=1).contiguous(), comm)
        ok = True
        msgs = []
        # identical step sequence (accept/reject pattern and step sizes)
        if [a[2] for a in full[3]] != [a[2] for a in mine[3]]:
            ok = False
            msgs.append("accept pattern differs")
        for a, b in zip(full[3], mine[3]):
            if abs(a[1] - b[1]) > 1e-12 * abs(a[1]) or abs(a[3] - b[3]) > 1e-9 * abs(a[3]):
                ok = False
                msgs.append("step %r vs %r" % (a, b))
        if not any(not a[2] for a in full[3]):
            ok = False
            msgs.append("case has no rejection")
        ref_out = shard_batch(full[0], rank, world, dim=1)
        ref_lam = shard_batch(full[1], rank, world)
        if not torch.allclose(mine[0], ref_out, rtol=1e-12, atol=1e-13):
            ok = False
            msgs.append("trajectory shard differs")
        if not torch.allclose(mine[1], ref_lam, rtol=1e-11, atol=1e-13):
            ok = False
            msgs.append("lambda shard differs")
        for a, b in zip(mine[2], full[2]):
            if not torch.allclose(a, b, rtol=1e-11, atol=1e-13):
                ok = False
                msgs.append("mu differs")
        # shard sizes: the ragged split (26 / 25 samples) is noticed by every rank, the even one passes on every rank
        local = shard_batch(u0, rank, world).shape[0]
        if comm.same_on_all_ranks(local) != (not ragged):
            ok = False
            msgs.append("same_on_all_ranks(%d) is wrong" % local)
        q.put((rank, ok, msgs[:3], comm.collectives))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_batch_sharded_world2_matches_single_process(ragged):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (1 if ragged else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msgs, ncoll in res:
        assert ok, (rank, msgs)
        assert ncoll >= 3  # one per step attempt + the mu all-reduce


def test_shard_batch_partitions_exactly():
    from pnode_b200.parallel import shard_batch

    x = torch.arange(23).reshape(23, 1)
    for world in (1, 2, 4, 8):
        parts = [shard_batch(x, r, world) for r in range(world)]
        assert torch.equal(torch.cat(parts), x)
        assert max(p.shape[0] for p in parts) - min(p.shape[0] for p in parts) <= 1
