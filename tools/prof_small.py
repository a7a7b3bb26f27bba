"""cProfile of the host side of the two latency-bound configurations (cfg1: spiral batch 20; cfg3: FFJORD B=1000), 50 passes each:
where the Python time of one fwd+adjoint pass goes (stream lookups, device->host reads, autograd plumbing)."""
import sys, os, cProfile, pstats, io, copy, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench_configs as bc
from pnode import petsc_adjoint
from pnode_b200.options import Options
for name, build in (("cfg3", bc._cnf(1000, "f32")), ("cfg1", bc.cfg1)):
    spec = build()
    Options.clear_all(); Options.insert_args(spec["argv"])
    dev = torch.device("cuda:0")
    to_dev = spec.get("to_dev", lambda f, d: f.to(d))
    funcs = [to_dev(copy.deepcopy(f), dev) for f in spec["funcs"]]
    u0, t, target = spec["u0"].to(dev), spec["t"].to(dev), spec["target"].to(dev)
    step, ode = bc._make_step(lambda: petsc_adjoint.ODEPetsc(), funcs, u0, t, target, spec["kw"], spec["step"], dev, spec.get("each_call_setup", False))
    for _ in range(5): step()
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(50): step()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
    print("=====", name); print("\n".join(s.getvalue().splitlines()[:60]))
