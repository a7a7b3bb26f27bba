import sys, os, copy, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import bench_configs as bc
from pnode import petsc_adjoint
from pnode_b200.options import Options
from pnode_b200 import _lib
lib = _lib.load()
def stats():
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    on = lib.pnode_graph_cache_stats(C.byref(a), C.byref(b), C.byref(c))
    return on, a.value, b.value, c.value
for code in sys.argv[1:]:
    name, build = bc.config_table()[code]
    spec = build()
    Options.clear_all(); Options.insert_args(spec["argv"])
    dev = torch.device("cuda:0")
    funcs = [copy.deepcopy(f).to(dev) for f in spec["funcs"]]
    u0, t, target = spec["u0"].to(dev), spec["t"].to(dev), spec["target"].to(dev)
    step, ode = bc._make_step(lambda: petsc_adjoint.ODEPetsc(), funcs, u0, t, target, spec["kw"], spec["step"], dev, spec.get("each_call_setup", False))
    prev = stats()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(40):
        torch.cuda.synchronize(); ev[0].record()
        step()
        ev[1].record(); torch.cuda.synchronize()
        s = stats()
        print(name, i, "on=%d replays+%d recorded+%d direct+%d  %.3f ms" % (s[0], s[1]-prev[1], s[2]-prev[2], s[3]-prev[3], ev[0].elapsed_time(ev[1])), flush=True)
        prev = s
