"""Noise floor of the REFERENCE algorithm for BASELINE config 5 at full size (KS N=1024, H=3200, B=256, fp64): the CPU oracle
against itself when (a) the Newton step with LU is replaced by the direct solve of the same system, (b) an explicit inverse is
used, (c) only the evaluation order inside f_I changes (dense product instead of conv1d).  Output in this container:
    direct traj 2.22e-08 lam 0.00e+00 mu 7.99e-12
    inv traj 2.73e-06 lam 0.00e+00 mu 1.59e-08
    perturb_guess traj 1.14e-08 lam 0.00e+00 mu 5.22e-12
The full-size parity test (tests/test_gpu_sinode.py) takes its tolerances from these numbers."""
import sys, os, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import OracleODEPetsc
import oracle.odepetsc as oo
from _workloads import KSExplicit, KSImplicit, ks_dx
torch.set_num_threads(8)
N,H,B,seed=1024,3200,256,4
def run(variant):
    g = torch.Generator().manual_seed(seed)
    u0 = 0.5 * torch.randn(B, N, generator=g, dtype=torch.float64)
    gout = torch.randn(2, B, N, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.2], dtype=torch.float64)
    argv = ["-ts_adapt_type", "none", "-snes_type", "ksponly", "-ts_arkimex_type", "3"]
    f_im, f_ex = KSImplicit(ks_dx(N)), KSExplicit(N, hidden=H)
    ode = OracleODEPetsc(argv)
    ode.setupTS(u0, f_im, step_size=0.2, method="imex", imex_form=True, func2=f_ex, batch_size=B, linear_solver="torch")
    cb=ode.cb
    if variant=="direct":
        def solve(t_, Z, shift, guess, aff=None):
            J=cb._jac(t_, guess); A=shift*torch.eye(N,dtype=torch.float64)-J
            return torch.linalg.solve(A,(shift*Z).reshape(-1,N).T).T.reshape(Z.shape)
        cb.implicit_solve=solve
    elif variant=="inv":
        def solve(t_, Z, shift, guess, aff=None):
            J=cb._jac(t_, guess); A=shift*torch.eye(N,dtype=torch.float64)-J
            return (shift*Z.reshape(-1,N))@torch.linalg.inv(A).T
        cb.implicit_solve=lambda *a,**k: solve(*a,**k).reshape(a[1].shape)
    elif variant=="perturb_guess":
        # same Newton-LU algorithm, but f_I evaluated through the dense matrix instead of conv1d (different rounding only)
        def f_im_(t_, u):
            J=cb._jac(t_, u); return (u.reshape(-1,N)@J.T).reshape(u.shape)
        cb.f_im=f_im_
    y0=u0.clone().requires_grad_(True)
    out=ode.odeint_adjoint(y0,t); (out*gout).sum().backward()
    mu=torch.cat([p.grad.reshape(-1) for p in f_ex.parameters()])
    return out[-1].detach(), y0.grad.clone(), mu
def rel(a,b): return float((a-b).abs().max()/b.abs().max())
base=run("oracle")
for v in ("direct","inv","perturb_guess"):
    r=run(v)
    print(v, "traj %.2e lam %.2e mu %.2e"%(rel(r[0],base[0]),rel(r[1],base[1]),rel(r[2],base[2])), flush=True)
