"""ORACLE (test infrastructure, NOT product code): restatement of the Python half of pnode --
`ODEPetsc.setupTS / odeint / odeint_adjoint`, the PETSc callbacks and the autograd bridge
(/root/reference/pnode/petsc_adjoint.py:366-947, /root/reference/pnode/torch_linearsolve.py:7-35) -- on top of
`oracle.petsc_ts.OracleTS` (the restated PETSc half).  Same class surface as the reference so that parity tests read
like /root/reference/tests/test_pnode.py.  Runs on CPU tensors, fp64 or fp32.
"""
import torch
import torch.nn as nn

from .petsc_ts import OracleTS, parse_petsc_options


def _flat_params(mod):
    return [p for p in mod.parameters() if p.requires_grad] if isinstance(mod, nn.Module) else []


def _cat(ts, like):
    flat = [t.contiguous().view(-1) for t in ts]
    return torch.cat(flat) if flat else torch.zeros(0, dtype=like.dtype, device=like.device)


class _Callbacks:
    """evalRHSFunction / evalIFunction / RHSJacShell.multTranspose / IJacShell._vjp / PCShell, restated."""

    def __init__(self, ode):
        self.o = ode

    # Without imex_form the reference registers ONE function: as the IFunction when implicit_form=True, else as the
    # RHSFunction (petsc_adjoint.py:666-730).  An ARKIMEX scheme then integrates an empty other half ([PETSc]: a missing
    # RHSFunction contributes 0, a missing IFunction is F = udot).
    def _no_ex(self):
        o = self.o
        return o.ts.kind == "arkimex" and not o.imex and bool(getattr(o, "implicit_form", False))

    def _no_im(self):
        o = self.o
        return o.ts.kind == "arkimex" and not o.imex and not bool(getattr(o, "implicit_form", False))

    def f_ex(self, t, u):
        if self._no_ex():
            return torch.zeros_like(u)
        with torch.no_grad():
            return self.o.funcEX(t, u.view(self.o.tensor_size)).reshape(u.shape).clone()  # petsc_adjoint.py:405

    def f_im(self, t, u):
        if self._no_im():
            return torch.zeros_like(u)
        with torch.no_grad():
            return self.o.funcIM(t, u.view(self.o.tensor_size)).reshape(u.shape).clone()  # petsc_adjoint.py:427

    def _vjp(self, func, t, u, w):
        params = _flat_params(func)
        with torch.enable_grad():
            x = u.detach().clone().view(self.o.tensor_size).requires_grad_(True)
            out = func(t, x)  # one forward re-evaluation per adjoint stage (petsc_adjoint.py:68)
            g = torch.autograd.grad(out, [x] + params, w.view(out.shape), allow_unused=True)
        vu = g[0] if g[0] is not None else torch.zeros_like(x)
        vp = [gi if gi is not None else torch.zeros_like(p) for gi, p in zip(g[1:], params)]  # misc.py:9-14
        return vu.reshape(u.shape), _cat(vp, u)

    def vjp_ex(self, t, u, w):
        if self._no_ex():
            return torch.zeros_like(u), None
        return self._vjp(self.o.funcEX, t, u, w)

    def vjp_im(self, t, u, w):
        if self._no_im():
            return torch.zeros_like(u), None
        return self._vjp(self.o.funcIM, t, u, w)

    def pad_params(self, vp_im, vp_ex):
        """mu layout = [mu_IM (npIM), mu_EX (npEX)] for IMEX (petsc_adjoint.py:322-330, 351-359)."""
        o = self.o
        if not o.imex:
            if vp_im is None and vp_ex is None:
                return torch.zeros(o.np, dtype=o.tensor_dtype)
            return vp_im if vp_im is not None else vp_ex
        z = lambda n: torch.zeros(n, dtype=o.tensor_dtype)
        return torch.cat((vp_im if vp_im is not None else z(o.npIM), vp_ex if vp_ex is not None else z(o.npEX)))

    # -- implicit stage: solve shift*(Y - Z) - f_I(t, Y) = 0 ------------------------------------------------------
    def _jac(self, t, y):
        o = self.o
        if self._no_im():
            n = y.numel() // (o.batch_size if o.linear_solver == "torch" and hasattr(o, "batch_size") else 1)
            if o.linear_solver == "torch":
                n = y.view(o.tensor_size)[0:1].numel()
            else:
                n = y.numel()
            return torch.zeros(n, n, dtype=y.dtype)
        if o.linear_solver == "torch":
            # dense Jacobian of sample 0 only (petsc_adjoint.py:479), applied per sample (torch_linearsolve.py:25-29)
            if o._J0 is None:
                y0 = y.view(o.tensor_size)[0:1].detach().clone()
                J = torch.autograd.functional.jacobian(lambda v: o.funcIM(t, v), y0)
                N = y0.numel()
                o._J0 = J.reshape(N, N)
            return o._J0
        yy = y.detach().clone().view(o.tensor_size)
        J = torch.autograd.functional.jacobian(lambda v: o.funcIM(t, v), yy)
        n = yy.numel()
        return J.reshape(n, n)

    def _lin_solve(self, J, shift, rhs, transpose):
        o = self.o
        N = J.shape[0]
        M = self.mass if self.mass is not None else torch.eye(N, dtype=J.dtype)
        A = shift * M - J
        if transpose:
            A = A.T
        if o.linear_solver == "torch":
            R = rhs.reshape(-1, N)
            return torch.linalg.solve(A, R.T).T.reshape(rhs.shape)
        return torch.linalg.solve(A, rhs.reshape(-1)).reshape(rhs.shape)

    @property
    def mass(self):
        return self.o.mass

    def implicit_solve(self, t, Z, shift, guess, aff=None):
        o = self.o
        y = guess.clone()
        F0 = None
        for it in range(50):
            if self.mass is not None:
                F = shift * (self.mass @ (y - Z).reshape(-1)).reshape(y.shape) - self.f_im(t, y)
            else:
                F = shift * (y - Z) - self.f_im(t, y)
            if aff is not None:
                F = F - aff
            fn = float(F.norm())
            if F0 is None:
                F0 = fn
            elif fn <= 1e-8 * F0 or fn <= 1e-50:
                break
            d = self._lin_solve(self._jac(t, y), shift, F, transpose=False)
            y = y - d
            if o.ksponly:
                break
            if float(d.norm()) <= 1e-8 * float(y.norm()):
                break
        return y

    def implicit_solve_transpose(self, t, y, shift, rhs):
        return self._lin_solve(self._jac(t, y), shift, rhs, transpose=True)


class OracleODEPetsc:
    """Same public surface as pnode.petsc_adjoint.ODEPetsc (petsc_adjoint.py:366-900)."""

    def __init__(self, argv=None):
        self.options = parse_petsc_options(argv)
        self.ts = OracleTS(self.options)
        self.tensor_size = None
        self.tensor_dtype = None
        self.device = None
        self.funcIM = None
        self.funcEX = None
        self.flat_params = None
        self.np = self.npIM = self.npEX = None
        self.imex = None
        self.linear_solver = None
        self.ksponly = self.options.get("snes_type") == "ksponly"
        self.mass = None
        self._J0 = None
        self.cb = _Callbacks(self)

    def setupTS(self, u_tensor, func, step_size=0.01, enable_adjoint=True, implicit_form=False, use_dlpack=True,
                method="dopri5", mass=None, imex_form=False, func2=None, batch_size=1, linear_solver="petsc",
                fixed_jacobian=False, matrixfree_jacobian=True, fixed_jacobian_across_solves=None):
        if imex_form and func2 is None:
            raise ValueError("func2 must be provided to enable imex_form=True")  # petsc_adjoint.py:585-586
        self.mass = None if mass is None else mass.reshape(u_tensor.numel(), u_tensor.numel()).to(u_tensor.dtype)
        self.imex = imex_form
        self.linear_solver = linear_solver
        if self.funcIM is not func or self.funcEX is not (func2 if imex_form else func):
            if imex_form:
                self.funcIM, self.funcEX = func, func2
                pim, pex = _flat_params(func), _flat_params(func2)
                self.npIM = sum(p.numel() for p in pim)
                self.flat_params = _cat(pim + pex, u_tensor)
                self.np = self.flat_params.numel()
                self.npEX = self.np - self.npIM
            else:
                self.funcIM = self.funcEX = func
                self.flat_params = _cat(_flat_params(func), u_tensor)
                self.np = self.npIM = self.npEX = self.flat_params.numel()
        if (u_tensor.size() != self.tensor_size or u_tensor.dtype != self.tensor_dtype
                or u_tensor.device != self.device):
            # the method is (re)applied only when the state's meta-data changes (petsc_adjoint.py:627-656)
            self.tensor_size = u_tensor.size()
            self.tensor_dtype = u_tensor.dtype
            self.device = u_tensor.device
            self.ts.kind, self.ts.scheme = "rk", "3bs"
            self.ts.set_method(method)
            self.implicit_form = implicit_form
        self.step_size = step_size
        self.enable_adjoint = enable_adjoint
        self.ts.set_from_options()

    # petsc_adjoint.py:518-532
    def _poststep(self, stepno, t):
        h = None
        if self.cur_sol_index < len(self.sol_times):
            if isinstance(self.step_size, list) and stepno < len(self.step_size):
                h = self.step_size[stepno]
            self.cur_sol_steps[self.cur_sol_index] += 1
            delta = 1e-5 if self.tensor_dtype == torch.double else 1e-3
            if abs(t - self.sol_times[self.cur_sol_index]) < delta:
                self.cur_sol_index += 1
        return h

    def odeint(self, u0, t):
        self._J0 = None  # reset_jacobianIM / reset_factor once per odeint (petsc_adjoint.py:792-799)
        u = u0.detach().clone()
        self.sol_times = [float(x) for x in t.cpu().to(torch.float64)]
        h0 = self.step_size[0] if isinstance(self.step_size, list) else self.step_size
        if t.shape[0] == 1:
            uf, _ = self.ts.solve(self.cb, u, 0.0, None, self.sol_times[0], h0, None, self.enable_adjoint)
            return torch.stack([uf], dim=0)
        self.cur_sol_steps = [0] * t.shape[0]
        self.cur_sol_index = 1
        _, sols = self.ts.solve(self.cb, u, self.sol_times[0], self.sol_times, self.sol_times[-1], h0, self._poststep,
                                self.enable_adjoint)
        if self.cur_sol_index != len(self.sol_times):
            raise Exception("TSSolve fails to step on all the specified points")  # petsc_adjoint.py:867-868
        return torch.stack([s.view(self.tensor_size) for s in sols], dim=0)

    def petsc_adjointsolve(self, t, lam, mu, i=1):
        if t.shape[0] == 1:
            n = int(round(abs(float(t[0]) / self.ts.h_last)))  # petsc_adjoint.py:875
            n = min(n, len(self.ts.traj))
        else:
            n = self.cur_sol_steps[i]
        return self.ts.adjoint_steps(self.cb, n, lam, mu)

    def odeint_adjoint(self, y0, t):
        if not isinstance(self.funcIM, nn.Module):
            raise ValueError("func is required to be an instance of nn.Module.")  # petsc_adjoint.py:896-897
        # flat_params must be the live cat() node so that mu reaches every Parameter.grad (petsc_adjoint.py:899)
        params = _flat_params(self.funcIM) + (_flat_params(self.funcEX) if self.imex else [])
        flat = _cat(params, y0)
        return _OracleAdjoint.apply(y0, t, flat, self)


class _OracleAdjoint(torch.autograd.Function):
    """OdeintAdjointMethod (petsc_adjoint.py:903-947)."""

    @staticmethod
    def forward(ctx, y0, t, flat_params, ode):
        ctx.ode = ode
        with torch.no_grad():
            ans = ode.odeint(y0, t)
        ctx.save_for_backward(t, ans)
        return ans

    @staticmethod
    def backward(ctx, grad_output):
        t, ans = ctx.saved_tensors
        ode = ctx.ode
        T = ans.shape[0]
        with torch.no_grad():
            lam = grad_output[-1].clone()
            mu = torch.zeros(ode.np, dtype=ans.dtype)
        if T == 1:
            lam, mu = ode.petsc_adjointsolve(t, lam, mu)
        for i in range(T - 1, 0, -1):
            lam, mu = ode.petsc_adjointsolve(t, lam, mu, i)
            lam = lam + grad_output[i - 1]  # forcing added after each segment (petsc_adjoint.py:938)
        return lam.detach().clone(), None, mu.detach().clone(), None
