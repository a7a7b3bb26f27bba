import sys, copy, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import OracleODEPetsc
from pnode import petsc_adjoint
from pnode_b200.options import Options
from _problems import rel_err
from _workloads import OdeConvBlock
for dtype in (torch.float64, torch.float32):
  for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
    C,HW,B=32,8,16
    func=OdeConvBlock(C,dtype=dtype); g=torch.Generator().manual_seed(3)
    u0=torch.randn(B,C,HW,HW,generator=g,dtype=torch.float64).to(dtype); gout=torch.randn(1,B,C,HW,HW,generator=g,dtype=torch.float64).to(dtype)
    t=torch.tensor([1.0],dtype=torch.float64)
    res=[]
    for dev,make in (("cpu",lambda: OracleODEPetsc(["-ts_adapt_type","none"])),("cuda",lambda: petsc_adjoint.ODEPetsc()), ("cuda", lambda: OracleODEPetsc(["-ts_adapt_type","none"]))):
        Options.clear_all(); Options.insert_args(["-ts_adapt_type","none"])
        f=copy.deepcopy(func).to(dev); ode=make()
        if isinstance(ode, OracleODEPetsc) and dev=="cuda":
            import oracle.odepetsc as oo
        try:
            ode.setupTS(u0.to(dev), f, step_size=0.5, method="rk4", enable_adjoint=True)
            y0=u0.to(dev).clone().requires_grad_(True); out=ode.odeint_adjoint(y0,t.to(dev)); (out*gout.to(dev)).sum().backward()
            res.append((out.detach().cpu(), y0.grad.cpu(), [p.grad.cpu() for p in f.parameters()]))
        except Exception as e:
            print("fail", dev, type(ode).__name__, repr(e)[:200]); res.append(None)
    o,p,og=res
    print(dtype, "tf32",tf32, "traj",rel_err(p[0],o[0]),"lam",rel_err(p[1],o[1]),"mu",max(rel_err(a,b) for a,b in zip(p[2],o[2])))
    if og is not None:
        print("   oracle-on-gpu vs oracle-cpu: traj",rel_err(og[0],o[0]),"lam",rel_err(og[1],o[1]))
