"""Generic-path time stepper: any `nn.Module` right-hand side, evaluated by torch ON THE GPU, with every vector
operation PETSc would do between callbacks (stage combination, completion + embedded error + weighted norm, adjoint
stage combination, lambda / mu accumulation) done by the fused kernels of csrc/vecops.cu -- one launch per stage
instead of one per AXPY term -- and stage checkpoints kept in HBM.

What it replaces: [PETSc] TSSolve / TSStep_RK / TSStep_ARKIMEX / TSStep_Theta / TSAdaptChoose /
TSTrajectory(memory) / TSAdjointStep_{RK,ARKIMEX,Theta}, and the arithmetic of the reference's callbacks
evalRHSFunction, evalIFunction, RHSJacShell.multTranspose, IJacShell._vjp, RHSJacPShell / IJacPShell.multTranspose
(pnode/petsc_adjoint.py:52-82, 179-196, 303-363, 393-441) and torch_linearsolve.PCShell (pnode/torch_linearsolve.py).
"""
import math

import torch

from .controller import TimeLoop
from .errors import Error


class Callbacks:
    """The two closures the engine needs from the user's modules: f(t,u) and vjp(t,u,w) -> (J^T w, Jp^T w).

    Stage evaluations of a solve that an adjoint sweep will follow keep their autograd graph (keep=True), keyed by the stage
    tensor: the adjoint stage at the same point then differentiates the kept graph instead of re-evaluating the module -- the
    reference re-evaluates (petsc_adjoint.py:64-70) because PETSc's trajectory only stores vectors; the arithmetic is the same,
    one forward evaluation per adjoint stage is saved (a third of the adjoint's cost).  It is switched off for modules whose
    forward has side effects the re-evaluation would repeat (train-mode BatchNorm statistics, active Dropout), by
    `-pnode_reuse_graph 0`, and beyond a memory budget (`-pnode_reuse_graph_mb`, default 8192)."""

    def __init__(self, func, tensor_size):
        self.func = func
        self.tensor_size = tensor_size
        self.params = [p for p in func.parameters() if p.requires_grad] if isinstance(func, torch.nn.Module) else []
        self.sizes = [p.numel() for p in self.params]
        self.nparams = sum(self.sizes)
        self.nfe = 0
        self.nvjp = 0
        self._graphs = {}
        self._graph_bytes = 0
        self._marks = []
        self.reused_graphs = 0
        from .options import Options

        self._graph_budget = int(float(Options().getString("pnode_reuse_graph_mb", "8192")) * (1 << 20))
        self._reuse = Options().getString("pnode_reuse_graph", "1") not in ("0", "false", "no") and self._side_effect_free()

    def _side_effect_free(self):
        if not isinstance(self.func, torch.nn.Module):
            return False
        for m in self.func.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm) and m.training and m.track_running_stats:
                return False
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout1d, torch.nn.Dropout2d, torch.nn.Dropout3d,
                              torch.nn.AlphaDropout)) and m.training and m.p > 0:
                return False
        return True

    def begin(self, forward, keep=True, comm=None):
        """Start of a forward solve (kept graphs of the previous one are dropped) / of an adjoint sweep."""
        if forward:
            self._graphs.clear()
            self._graph_bytes = 0
            self._marks = []

    def release(self):
        """End of the adjoint sweep: nothing kept by the forward is needed any more (frees the graphs' activations now instead
        of at the next forward solve)."""
        self._graphs.clear()
        self._graph_bytes = 0
        self._marks = []

    def mark(self):
        """Remember the kept graphs so far: a rejected step attempt rolls back to here."""
        self._marks = list(self._graphs.keys())

    def rollback(self):
        for k in [k for k in self._graphs if k not in self._marks]:
            self._graph_bytes -= self._graphs.pop(k)[5]

    def _pversions(self):
        return tuple(p._version for p in self.params)

    def f(self, t, u, keep=False):
        """evalRHSFunction (petsc_adjoint.py:393-405): t is handed over as a python float."""
        self.nfe += 1
        if keep and self._reuse and self._graph_bytes < self._graph_budget:
            before = torch.cuda.memory_allocated(u.device) if u.is_cuda else 0
            with torch.enable_grad():
                x = u.detach().view(self.tensor_size).requires_grad_(True)
                out = self.func(t, x)
            if isinstance(out, torch.Tensor) and out.requires_grad:
                # what the kept graph holds: measured on the device; on the host-logic test double a nominal estimate
                nbytes = max(torch.cuda.memory_allocated(u.device) - before, 0) if u.is_cuda else 4 * out.numel() * out.element_size()
                self._graphs[u.data_ptr()] = (u, u._version, self._pversions(), x, out, nbytes, float(t))
                self._graph_bytes += nbytes
            flat = out.detach().reshape(-1)
            if flat.dtype != u.dtype:
                flat = flat.to(u.dtype)
            return flat if flat.is_contiguous() else flat.contiguous()
        with torch.no_grad():
            out = self.func(t, u.view(self.tensor_size))
        out = out.detach().reshape(-1)
        if out.dtype != u.dtype:
            out = out.to(u.dtype)
        return out if out.is_contiguous() else out.contiguous()

    def jvp(self, t, u, v):
        """RHSJacShell.mult / IJacShell._jvp (petsc_adjoint.py:19-49, 129-144): J v, by forward-mode AD."""
        self.nfe += 1
        f = lambda x: self.func(t, x.view(self.tensor_size)).reshape(-1)
        try:
            with torch.no_grad():
                _, jv = torch.func.jvp(f, (u.detach().reshape(-1),), (v.detach().reshape(-1),))
        except Exception:  # op without a forward-mode rule: the reference's double-backward trick
            with torch.enable_grad():
                x = u.detach().reshape(-1).requires_grad_(True)
                out = f(x)
                dummy = torch.zeros_like(out, requires_grad=True)
                (g,) = torch.autograd.grad(out, x, dummy, create_graph=True)
                (jv,) = torch.autograd.grad(g, dummy, v.detach().reshape(-1))
        return jv.detach().contiguous()

    def vjp(self, t, u, w, want_u=True, want_params=True):
        """RHSJacShell.multTranspose (petsc_adjoint.py:52-82): one autograd.grad gives J^T w and the per-parameter
        (df/dp)^T w (None for unused parameters, misc.py:9-14)."""
        self.nvjp += 1
        ent = self._graphs.get(u.data_ptr())
        if ent is not None and ent[1] == u._version and ent[2] == self._pversions() and ent[6] == float(t) and \
                ent[0].numel() == u.numel():
            # the graph of the forward evaluation at this very stage is still alive: differentiate it (no re-evaluation)
            x, out = ent[3], ent[4]
            self.reused_graphs += 1
            if isinstance(getattr(self.func, "nfe", None), int):
                self.func.nfe += 1  # observable count of the reference, whose adjoint calls func once per stage
            inputs = ([x] if want_u else []) + (self.params if want_params else [])
            if not inputs:
                return None, []
            g = torch.autograd.grad(out, inputs, w.view(out.shape).to(out.dtype), allow_unused=True, retain_graph=True)
        else:
            with torch.enable_grad():
                x = u.detach().view(self.tensor_size).requires_grad_(True)
                out = self.func(t, x)
                inputs = ([x] if want_u else []) + (self.params if want_params else [])
                if not inputs:
                    return None, []
                g = torch.autograd.grad(out, inputs, w.view(out.shape).to(out.dtype), allow_unused=True)
        if want_u:
            vu, gp = g[0], list(g[1:])
            vu = torch.zeros_like(u) if vu is None else vu.reshape(-1).contiguous()
        else:
            vu, gp = None, list(g)
        return vu, gp


def gmres(ops, apply_op, b, rtol=1e-5, atol=1e-50, restart=30, max_it=10000):
    """Restarted GMRES with classical Gram-Schmidt, zero initial guess, no preconditioner -- [PETSc] KSPGMRES as the reference
    configures it by default for its matrix-free shells.  The Krylov basis lives in HBM; per inner iteration:
    one operator application, one fused multi-dot (pnode_mdot: all Gram-Schmidt coefficients + ||w||^2 in one pass, one
    8(k+2)-byte host read), one pnode_lincomb for the orthogonalisation, one multi-dot for the new norm.  The small
    Hessenberg least-squares problem is solved on the host with Givens rotations.  Returns (x, iterations, residual)."""
    x = torch.zeros_like(b)
    _, bb = ops.mdot([], b)
    bnorm = math.sqrt(bb)
    if bnorm == 0.0:
        return x, 0, 0.0
    tol = max(rtol * bnorm, atol)
    r, rnorm, its = b, bnorm, 0
    while its < max_it:
        V = [torch.empty_like(b)]
        ops.lincomb(V[0], None, 0.0, [r], [1.0 / rnorm])
        H = [[0.0] * restart for _ in range(restart + 1)]
        cs, sn, g = [0.0] * restart, [0.0] * restart, [rnorm] + [0.0] * restart
        k_used = 0
        for k in range(restart):
            w = apply_op(V[k])
            h, _ = ops.mdot(V[:k + 1], w)
            wn = torch.empty_like(w)
            ops.lincomb(wn, w, 1.0, V[:k + 1], [-c for c in h])
            _, nn = ops.mdot([], wn)
            hk1 = math.sqrt(max(nn, 0.0))
            col = list(h) + [hk1]
            for i in range(k):  # apply the previous rotations to the new column
                a, c = col[i], col[i + 1]
                col[i], col[i + 1] = cs[i] * a + sn[i] * c, -sn[i] * a + cs[i] * c
            den = math.hypot(col[k], col[k + 1])
            cs[k], sn[k] = (col[k] / den, col[k + 1] / den) if den != 0.0 else (1.0, 0.0)
            col[k], col[k + 1] = den, 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            for i in range(k + 1):
                H[i][k] = col[i]
            its += 1
            k_used = k + 1
            rnorm = abs(g[k + 1])
            if rnorm <= tol or hk1 == 0.0 or its >= max_it:
                break
            V.append(torch.empty_like(b))
            ops.lincomb(V[k + 1], None, 0.0, [wn], [1.0 / hk1])
        y = [0.0] * k_used
        for i in range(k_used - 1, -1, -1):
            acc = g[i] - sum(H[i][j] * y[j] for j in range(i + 1, k_used))
            y[i] = acc / H[i][i] if H[i][i] != 0.0 else 0.0
        ops.lincomb(x, x, 1.0, V[:k_used], y)
        if rnorm <= tol or its >= max_it:
            break
        ax = apply_op(x)  # restart: true residual
        r = torch.empty_like(b)
        ops.lincomb(r, b, 1.0, [ax], [-1.0])
        _, rr = ops.mdot([], r)
        rnorm = math.sqrt(rr)
        if rnorm <= tol:
            break
    return x, its, rnorm


def block_gmres(ops, apply_op, b, nseg, rtol=1e-5, atol=1e-50, restart=30, max_it=10000):
    """GMRES on `nseg` right-hand sides at once -- the reference's linear_solver="hpddm" (hpddm_linearsolve.py:13-49: KSPHPDDM,
    type BGMRES, no preconditioner, the state seen as a dense [n/batch x batch] matrix and handed to KSPMatSolve).  Every
    segment (sample) of the flattened state carries its own Krylov space and its own relative tolerance, as HPDDM's
    per-column convergence test does; the segments advance in lockstep, so one iteration is ONE operator application on the
    whole batch, one pnode_mdot_seg (all Gram-Schmidt coefficients of all segments, left on the device), one
    pnode_lincomb_seg for the orthogonalisation, one more of each for the new norms and the normalisation, and one
    (k+2) x nseg host read for the nseg small Hessenberg problems (Givens rotations, vectorised over the segments in numpy).
    A segment that converges exactly (zero new basis vector) drops out by itself.  Zero initial guess (the reference
    randomises the guess only when -ksp_initial_guess_nonzero is set).  Returns (x, iterations, max residual norm)."""
    import numpy as np

    n = b.numel()
    x = torch.zeros_like(b)
    bb = ops.mdot_seg([], b, nseg)
    bnorm = np.sqrt(np.maximum(bb[0].cpu().numpy(), 0.0))
    if not (bnorm > 0.0).any():
        return x, 0, 0.0
    tol = np.maximum(rtol * bnorm, atol)
    r, rr_dev, rnorm, its = b, bb, bnorm.copy(), 0
    while its < max_it:
        V = [torch.empty_like(b)]
        ops.lincomb_seg(V[0], None, 0.0, [r], rr_dev[-1:], 2, nseg)  # r / ||r|| per segment
        H = np.zeros((nseg, restart + 1, restart))
        cs, sn = np.zeros((nseg, restart)), np.zeros((nseg, restart))
        g = np.zeros((nseg, restart + 1))
        g[:, 0] = rnorm
        k_used = 0
        for k in range(restart):
            w = apply_op(V[k])
            hd = ops.mdot_seg(V[:k + 1], w, nseg)
            wn = torch.empty_like(w)
            ops.lincomb_seg(wn, w, 1.0, V[:k + 1], hd, 1, nseg)
            nn = ops.mdot_seg([], wn, nseg)
            host = torch.cat((hd[:k + 1], nn)).cpu().numpy()  # the one host read of the iteration
            col = np.empty((nseg, k + 2))
            col[:, :k + 1] = host[:k + 1].T
            hk1 = np.sqrt(np.maximum(host[k + 1], 0.0))
            col[:, k + 1] = hk1
            for i in range(k):  # previous rotations on the new column
                a, c = col[:, i].copy(), col[:, i + 1].copy()
                col[:, i], col[:, i + 1] = cs[:, i] * a + sn[:, i] * c, -sn[:, i] * a + cs[:, i] * c
            den = np.hypot(col[:, k], col[:, k + 1])
            ok = den != 0.0
            safe = np.where(ok, den, 1.0)
            cs[:, k] = np.where(ok, col[:, k] / safe, 1.0)
            sn[:, k] = np.where(ok, col[:, k + 1] / safe, 0.0)
            col[:, k], col[:, k + 1] = den, 0.0
            g[:, k + 1] = -sn[:, k] * g[:, k]
            g[:, k] = cs[:, k] * g[:, k]
            H[:, :k + 1, k] = col[:, :k + 1]
            its += 1
            k_used = k + 1
            rnorm = np.abs(g[:, k + 1])
            if ((rnorm <= tol) | (hk1 == 0.0)).all() or its >= max_it:
                break
            V.append(torch.empty_like(b))
            ops.lincomb_seg(V[k + 1], None, 0.0, [wn], nn, 2, nseg)
        y = np.zeros((nseg, k_used))
        for i in range(k_used - 1, -1, -1):
            acc = g[:, i] - (H[:, i, i + 1:k_used] * y[:, i + 1:]).sum(1)
            d = H[:, i, i]
            y[:, i] = np.where(d != 0.0, acc / np.where(d != 0.0, d, 1.0), 0.0)
        yd = torch.from_numpy(np.ascontiguousarray(y.T)).to(b.device)
        ops.lincomb_seg(x, x, 1.0, V[:k_used], yd, 0, nseg)
        if (rnorm <= tol).all() or its >= max_it:
            break
        ax = apply_op(x)  # restart: true residuals
        r = torch.empty_like(b)
        ops.lincomb(r, b, 1.0, [ax], [-1.0])
        rr_dev = ops.mdot_seg([], r, nseg)
        rnorm = np.sqrt(np.maximum(rr_dev[0].cpu().numpy(), 0.0))
        if (rnorm <= tol).all():
            break
    return x, its, float(rnorm.max())


class ImplicitSolver:
    """Solve shift*(Y - Z) - f_I(t, Y) = 0 for one implicit stage, and the transposed linearised system for the adjoint.

    linear_solver == "torch" (torch_linearsolve.PCShell): f_I is sample-independent; the dense [N,N] Jacobian of sample 0
        (petsc_adjoint.py:479) is inverted once per shift and applied to all samples as ONE GEMM  X <- R @ inv(A)^T
        ("factor once per step size, inverse-apply as tensor-core GEMM").
    "hpddm" (hpddm_linearsolve.PCShell): Newton on the full state with the block solver `block_gmres` above, one right-hand
        side per sample (batch_size segments).
    otherwise ("petsc"): Newton on the full state.  Small systems (n <= DENSE_SMALL, ROBER-like) assemble the
        dense Jacobian and solve by LU; larger ones are matrix-free Newton-GMRES like the reference's IJacShell
        (petsc_adjoint.py:98-196): J x by forward-mode AD (torch.func.jvp; the reference uses the double-backward trick),
        J^T x by one reverse-mode pass, `gmres` above as the Krylov solver (-ksp_rtol, -ksp_max_it).
    """

    DENSE_LIMIT = 8192
    DENSE_SMALL = 256

    def __init__(self, ops, cb_im, linear_solver, batch_size, ksponly, rtol=1e-8, max_it=50, ksp_rtol=1e-5,
                 ksp_max_it=10000):
        self.ksp_rtol = ksp_rtol
        self.ksp_max_it = ksp_max_it
        self.krylov_iterations = 0
        self.ops = ops
        self.cb = cb_im
        self.linear_solver = linear_solver
        self.batch = max(int(batch_size), 1)
        self.ksponly = ksponly
        self.rtol = rtol
        self.max_it = max_it
        self.reset()

    mass = None  # dense [n, n] mass matrix of M u' = f(t, u) (petsc_adjoint.py:426-431), theta methods only

    # fixed_jacobian=True ("the Jacobian is constant across ODE solves", petsc_adjoint.py:582): the block Jacobian and its
    # inverses survive from one odeint to the next -- and across setupTS calls, through `cache`, a dict the ODEPetsc object owns
    # -- as long as no parameter or buffer of the implicit function has been modified.  The reference documents the flag but
    # rebuilds anyway (its reset condition at 792-795 is always true); honouring it removes an O(N^3) factorisation per solve.
    fixed_jacobian = False
    cache = None

    def _signature(self):
        f = self.cb.func
        if not isinstance(f, torch.nn.Module):
            return None
        return tuple((id(q), q._version) for q in list(f.parameters()) + list(f.buffers())) + (self.batch, self.linear_solver)

    def reset(self):
        """Once per odeint: parameters may have changed (petsc_adjoint.py:792-799)."""
        if self.fixed_jacobian and self.cache is not None:
            sig = self._signature()
            if sig is not None and self.cache.get("sig") == sig and self.cache.get("func") is self.cb.func:
                self._J0, self._inv = self.cache["J0"], self.cache["inv"]
                return
            self._J0, self._inv = None, {}
            self.cache.update(sig=sig, func=self.cb.func, J0=None, inv=self._inv)
            return
        self._J0 = None
        self._inv = {}

    def _mass_times(self, v, transpose=False):
        if self.mass is None:
            return v
        return torch.mv(self.mass.T if transpose else self.mass, v.reshape(-1))

    def _block_jacobian(self, t, y):
        if self._J0 is None:
            N = y.numel() // self.batch
            y0 = y.view(self.cb.tensor_size)[0:1].detach().clone()
            # vmap'd reverse mode like the reference (torch.func.jacrev, petsc_adjoint.py:475-480): one batched backward
            # instead of N sequential ones
            try:
                J = torch.func.jacrev(lambda v: self.cb.func(t, v))(y0)
            except Exception:
                J = torch.autograd.functional.jacobian(lambda v: self.cb.func(t, v), y0)
            self._J0 = J.reshape(N, N)
            if self.fixed_jacobian and self.cache is not None:
                self.cache["J0"] = self._J0
        return self._J0

    def _block_inverse(self, t, y, shift):
        key = float(shift)
        if key not in self._inv:
            J = self._block_jacobian(t, y)
            N = J.shape[0]
            if self.mass is not None:
                # reference petsc_adjoint.py:491-507 forms shift * mass - J; per-sample blocks need a block-diagonal mass
                # matrix with identical blocks
                M = self.mass
                Mb = M[:N, :N]
                if not torch.equal(M, torch.block_diag(*([Mb] * (M.shape[0] // N)))):
                    raise Error(-21, "linear_solver='torch' with mass= needs a block-diagonal mass matrix with identical "
                                     "[n/batch, n/batch] blocks")
                A = Mb * shift - J
            else:
                A = torch.eye(N, dtype=J.dtype, device=J.device).mul_(shift) - J
            self._inv[key] = torch.linalg.inv(A)
        return self._inv[key]

    def _dense_matrix(self, t, y, shift):
        n = y.numel()
        if n > self.DENSE_LIMIT:
            raise Error(-20, "implicit stage with a full-state dense Jacobian needs n <= %d (got %d); use "
                             "linear_solver='torch' for batched sample-independent operators" % (self.DENSE_LIMIT, n))
        yy = y.detach().clone().view(self.cb.tensor_size)
        J = torch.autograd.functional.jacobian(lambda v: self.cb.func(t, v), yy).reshape(n, n)
        if self.mass is not None:
            return self.mass * shift - J
        return torch.eye(n, dtype=J.dtype, device=J.device).mul_(shift) - J

    def _apply(self, t, y, shift, rhs, transpose):
        if self.linear_solver == "torch":
            Ainv = self._block_inverse(t, y, shift)
            N = Ainv.shape[0]
            R = rhs.view(-1, N)
            # per sample x = A^{-1} r  <=>  X = R A^{-T};   transposed solve: X = R A^{-1}
            return (R @ (Ainv if transpose else Ainv.T)).reshape(-1)
        if y.numel() > self.DENSE_SMALL or self.linear_solver == "hpddm":
            return self._krylov(t, y, shift, rhs.reshape(-1), transpose)
        A = self._dense_matrix(t, y, shift)
        return torch.linalg.solve(A.T if transpose else A, rhs.reshape(-1))

    def _krylov(self, t, y, shift, rhs, transpose):
        cb, ops = self.cb, self.ops
        y0 = y.detach()

        def op(v):  # (shift M - J) v   or its transpose
            if transpose:
                jv, _ = cb.vjp(t, y0, v, want_params=False)
            else:
                jv = cb.jvp(t, y0, v)
            out = torch.empty_like(v)
            ops.lincomb(out, self._mass_times(v, transpose), shift, [jv], [-1.0])
            return out

        if self.linear_solver == "hpddm" and rhs.numel() % self.batch == 0:
            x, its, _ = block_gmres(ops, op, rhs.contiguous(), self.batch, rtol=self.ksp_rtol, max_it=self.ksp_max_it)
        else:
            x, its, _ = gmres(ops, op, rhs.contiguous(), rtol=self.ksp_rtol, max_it=self.ksp_max_it)
        self.krylov_iterations += its
        return x

    def solve(self, t, Z, shift, guess, aff=None):
        """Newton on  shift * M (Y - Z) - f(t, Y) - aff = 0   (M = I, aff = 0 unless a mass matrix is set)."""
        y = guess
        f0 = None
        for _ in range(self.max_it):
            F = torch.empty_like(Z)
            if self.mass is None and aff is None:
                self.ops.lincomb(F, None, 0.0, [y, Z, self.cb.f(t, y)], [shift, -shift, -1.0])
            else:
                d0 = torch.empty_like(Z)
                self.ops.lincomb(d0, None, 0.0, [y, Z], [1.0, -1.0])
                terms, coefs = [self._mass_times(d0), self.cb.f(t, y)], [shift, -1.0]
                if aff is not None:
                    terms.append(aff)
                    coefs.append(-1.0)
                self.ops.lincomb(F, None, 0.0, terms, coefs)
            if not self.ksponly:
                fn = float(torch.linalg.vector_norm(F))
                if f0 is None:
                    f0 = fn
                elif fn <= self.rtol * f0 or fn <= 1e-50:
                    break
            d = self._apply(t, y, shift, F, transpose=False)
            ynew = torch.empty_like(Z)
            self.ops.lincomb(ynew, y, 1.0, [d], [-1.0])
            y = ynew
            if self.ksponly:  # -snes_type ksponly: exactly one Newton step
                break
            if float(torch.linalg.vector_norm(d)) <= 1e-8 * float(torch.linalg.vector_norm(y)):
                break
        return y

    def solve_transpose(self, t, y, shift, rhs):
        return self._apply(t, y, shift, rhs, transpose=True)


def _beta(c, r):
    """Binomial number beta(c, r) = C(c + r, c): the longest step sequence that c checkpoints (the one holding the start state
    included) reverse with no step advanced more than r times (Griewank & Walther, "revolve"); 0 for negative arguments."""
    if c < 0 or r < 0:
        return 0
    return math.comb(c + r, c)


def revolve_split(n, free):
    """Where to put the next checkpoint when n >= 2 steps starting at a stored state are to be reversed and `free` >= 1 more
    checkpoint slots are available: advance m steps, store, reverse the n - m steps behind it, then the first m.  The choice of
    Griewank & Walther (Algorithm 799) with c = free + 1 checkpoints: it minimises the number of recomputed steps."""
    c = free + 1
    r = 1
    while _beta(c, r) < n:
        r += 1
    if n <= _beta(c, r - 1) + _beta(c - 2, r - 1):
        m = _beta(c, r - 2)
    elif n >= _beta(c, r) - _beta(c - 3, r):
        m = _beta(c, r - 1)
    else:
        m = n - _beta(c - 1, r - 1) - _beta(c - 2, r - 1)
    return max(1, min(m, n - 1))


def revolve_forward_positions(n, c):
    """Step indices whose START state the forward sweep of n steps keeps when c solution checkpoints fit (index 0 always)."""
    pos, i0, nn, cc = [0], 0, n, c - 1  # cc: slots still free once the start state of the current segment is held
    while cc > 0 and nn > 1:
        m = revolve_split(nn, cc)
        i0 += m
        pos.append(i0)
        nn -= m
        cc -= 1
    return pos


class GenericTS:
    """Host-driven stage loop over device kernels.  `ops` is a pnode_b200.device.DeviceOps."""

    def __init__(self, ops, scheme, kind, atol, rtol, comm=None, solution_only=False, max_cps=None):
        self.ops = ops
        self.scheme = scheme
        self.kind = kind
        self.atol = atol
        self.rtol = rtol
        self.comm = comm  # optional data-parallel communicator (pnode_b200.parallel.BatchComm)
        # [PETSc] TSTrajectory memory: `-ts_trajectory_solution_only 1` keeps u_n only and recomputes the
        # stages in the adjoint; `-ts_trajectory_max_cps_ram N` keeps at most N solution checkpoints in HBM at any time, placed
        # by the binomial (revolve) schedule: fixed-step runs know their step count, so the forward sweep stores exactly the
        # states of the offline schedule; adaptive runs thin a uniform grid on the fly (half of the slots) and the reverse sweep
        # re-checkpoints inside each gap with the slots that have become free.  Same kernels, same arithmetic: identical results.
        self.solution_only = solution_only or (max_cps is not None)
        self.max_cps = max_cps
        self.traj = []
        self.recomputed_steps = 0
        self._keep = False  # stage evaluations keep what the adjoint stage at the same point can reuse

    # -- forward ----------------------------------------------------------------------------------------------------
    def solve(self, cb_ex, cb_im, imp, u0, loop: TimeLoop, save_trajectory):
        ops = self.ops
        self.traj = []
        self._stride = 1
        self.recomputed_steps = 0
        self.peak_checkpoints = 0
        self._keep_at = None
        if save_trajectory and self.max_cps is not None and not loop.adaptive:
            import copy as _copy
            dry, n = _copy.deepcopy(loop), 0  # the step schedule of a fixed-step run is known before it starts
            while not dry.done:
                dry.report(None)
                n += 1
            self._keep_at = set(revolve_forward_positions(n, max(self.max_cps, 1)))
        for cb in (cb_ex, cb_im):
            if cb is not None and hasattr(cb, "begin"):
                cb.begin(True, keep=save_trajectory, comm=self.comm)
        self._keep = bool(save_trajectory)
        u = u0.reshape(-1).clone()
        n_local = u.numel()
        n_global = n_local if self.comm is None else self.comm.global_count(n_local)
        sols = {0: u} if loop.span is not None else {}
        k_fsal = None
        while not loop.done:
            t, h = loop.t, loop.h
            if self.kind == "rk":
                unew, stages, sumsq = self._rk_attempt(cb_ex, t, h, u, k_fsal, loop.adaptive)
            elif self.kind == "arkimex":
                unew, stages, sumsq = self._ark_attempt(cb_ex, cb_im, imp, t, h, u, loop.adaptive)
            else:
                unew, stages, sumsq = self._theta_attempt(cb_im, imp, t, h, u)
            enorm = None
            if loop.adaptive:
                if self.comm is not None:
                    self.comm.allreduce_scalar(sumsq)  # one scalar per attempt: identical decision on every rank
                enorm = math.sqrt(float(sumsq.item()) / n_global)
            if not loop.report(enorm):
                k_fsal = None
                for cb in (cb_ex, cb_im):
                    if cb is not None and hasattr(cb, "rollback"):
                        cb.rollback()  # graphs / activation sets of the rejected attempt
                continue
            for cb in (cb_ex, cb_im):
                if cb is not None and hasattr(cb, "mark"):
                    cb.mark()
            if save_trajectory:
                self._record(t, h, u, stages)
            if self.kind == "rk" and self.scheme.fsal:
                k_fsal = stages[1][-1]
            u = unew
            if loop.last_out_slot >= 0:
                sols[loop.last_out_slot] = u
        loop.check_complete()
        return u, sols

    def _stored(self):
        return sum(1 for e in self.traj if e[3] is not None)

    def _record(self, t, h, u, stages):
        if not self.solution_only:
            self.traj.append((t, h, stages, None))
            return
        if self.max_cps is None:
            self.traj.append((t, h, None, u))
            return
        idx = len(self.traj)
        if self._keep_at is not None:  # offline binomial schedule
            self.traj.append((t, h, None, u if idx in self._keep_at else None))
        else:
            # step count unknown (adaptive): uniform grid thinned by stride doubling within half of the slots; the other half
            # is left to the reverse sweep's re-checkpointing
            self.traj.append((t, h, None, u))
            budget = max((max(self.max_cps, 1) + 1) // 2, 1)
            held = [i for i, e in enumerate(self.traj) if e[3] is not None]
            while len(held) > budget:
                self._stride *= 2
                for i in held:
                    if i % self._stride != 0:
                        tt, hh, _, _ = self.traj[i]
                        self.traj[i] = (tt, hh, None, None)
                held = [i for i, e in enumerate(self.traj) if e[3] is not None]
        self.peak_checkpoints = max(self.peak_checkpoints, self._stored())

    def _advance(self, cb_ex, cb_im, imp, i, u):
        t, h = self.traj[i][0], self.traj[i][1]
        if self.kind == "rk":
            unew, st, _ = self._rk_attempt(cb_ex, t, h, u, None, False)
        elif self.kind == "arkimex":
            unew, st, _ = self._ark_attempt(cb_ex, cb_im, imp, t, h, u, False)
        else:
            unew, st, _ = self._theta_attempt(cb_im, imp, t, h, u)
        self.recomputed_steps += 1
        return unew, st

    def _restore(self, cb_ex, cb_im, imp, idx):
        """Make step `idx` (the last one still in the trajectory) hold its stage values again: advance from the nearest stored
        solution at or before it (same kernels, same arithmetic => identical stages).  With a checkpoint budget the advance
        drops new checkpoints at the binomial split points while slots are free (revolve); a popped step frees its slot."""
        j = idx
        while self.traj[j][3] is None:
            j -= 1
        u = self.traj[j][3]
        if self.max_cps is None:
            # every start state is kept: one step to recompute
            for i in range(j, idx + 1):
                unew, st = self._advance(cb_ex, cb_im, imp, i, u)
                t, h, _, u_keep = self.traj[i]
                self.traj[i] = (t, h, st, u_keep)
                u = unew
            return
        pos = j
        while pos < idx:
            free = max(self.max_cps, 1) - self._stored()
            n = idx + 1 - pos
            m = revolve_split(n, free) if free >= 1 else n - 1
            for i in range(pos, pos + m):
                u, _ = self._advance(cb_ex, cb_im, imp, i, u)
            pos += m
            if free >= 1 and pos < idx:
                t, h, st, _ = self.traj[pos]
                self.traj[pos] = (t, h, st, u)
                self.peak_checkpoints = max(self.peak_checkpoints, self._stored())
        _, st = self._advance(cb_ex, cb_im, imp, idx, u)
        t, h, _, u_keep = self.traj[idx]
        self.traj[idx] = (t, h, st, u_keep)

    def _rk_attempt(self, cb, t, h, u, k_fsal, adaptive):
        sc, ops = self.scheme, self.ops
        s = sc.s
        Y, K = [], []
        y_next = None  # Y_{i+1} already formed by a right-hand-side evaluator that fuses the stage combination into its output pass
        for i in range(s):
            if i == 0:
                y = u
            elif y_next is not None:
                y = y_next
            else:
                idx = [j for j in range(i) if sc.A[i][j] != 0.0]
                y = torch.empty_like(u)
                ops.lincomb(y, u, 1.0, [K[j] for j in idx], [h * sc.A[i][j] for j in idx])
            Y.append(y)
            y_next = None
            if i == 0 and k_fsal is not None:
                K.append(k_fsal)
            elif (hasattr(cb, "f_and_combine") and i + 1 < s
                  and [j for j in range(i + 1) if sc.A[i + 1][j] != 0.0] == [i]):
                # Y_{i+1} = u + h a_{i+1,i} k_i depends on k_i alone (euler / midpoint / rk4): one pass writes k_i and Y_{i+1}
                k, y_next = cb.f_and_combine(t + sc.c[i] * h, y, u, h * sc.A[i + 1][i])
                K.append(k)
            else:
                K.append(cb.f(t + sc.c[i] * h, y, keep=self._keep))
        idx = [j for j in range(s) if sc.b[j] != 0.0 or (adaptive and sc.bembed[j] != 0.0)]
        unew = torch.empty_like(u)
        ew = [h * (sc.bembed[j] - sc.b[j]) for j in idx] if adaptive else None
        sumsq = ops.complete(unew, u, [K[j] for j in idx], [h * sc.b[j] for j in idx], ew, self.atol, self.rtol)
        return unew, (Y, K), sumsq

    def _ark_attempt(self, cb_ex, cb_im, imp, t, h, u, adaptive):
        sc, ops = self.scheme, self.ops
        s = sc.s
        Y, KI, KE = [], [], []
        for i in range(s):
            vecs, coefs = [], []
            for j in range(i):
                if sc.At[i][j] != 0.0 and KI[j] is not None:
                    vecs.append(KI[j])
                    coefs.append(h * sc.At[i][j])
                if sc.A[i][j] != 0.0 and KE[j] is not None:
                    vecs.append(KE[j])
                    coefs.append(h * sc.A[i][j])
            if vecs:
                Z = torch.empty_like(u)
                ops.lincomb(Z, u, 1.0, vecs, coefs)
            else:
                Z = u
            if cb_im is None:
                # only an RHSFunction is registered (reference petsc_adjoint.py:716-730): F = udot, the stage equation
                # shift (Y - Z) = 0 gives Y = Z and K^I = 0
                y, ki = Z, None
            elif sc.At[i][i] == 0.0:
                y = Z
                ki = cb_im.f(t + sc.ct[i] * h, y, keep=self._keep)
            else:
                shift = 1.0 / (h * sc.At[i][i])
                y = imp.solve(t + sc.ct[i] * h, Z, shift, Y[i - 1] if i > 0 else u)
                ki = torch.empty_like(u)
                ops.lincomb(ki, None, 0.0, [y, Z], [shift, -shift])  # K^I_i = shift (Y_i - Z), not re-evaluated
            Y.append(y)
            KI.append(ki)
            KE.append(None if cb_ex is None else cb_ex.f(t + sc.c[i] * h, y, keep=self._keep))
        vecs, bw, ew = [], [], []
        for j in range(s):
            be = sc.bembed[j] if (adaptive and sc.bembed is not None) else sc.b[j]
            if (sc.bt[j] != 0.0 or be != sc.bt[j]) and KI[j] is not None:
                vecs.append(KI[j])
                bw.append(h * sc.bt[j])
                ew.append(h * (be - sc.bt[j]))
            if (sc.b[j] != 0.0 or be != sc.b[j]) and KE[j] is not None:
                vecs.append(KE[j])
                bw.append(h * sc.b[j])
                ew.append(h * (be - sc.b[j]))
        unew = torch.empty_like(u)
        sumsq = ops.complete(unew, u, vecs, bw, ew if adaptive else None, self.atol, self.rtol)
        return unew, (Y,), sumsq

    def _theta_attempt(self, cb, imp, t, h, u):
        """theta-method, endpoint form: M (X - u) / h = (1 - theta) f(t, u) + theta f(t + h, X)."""
        theta = 0.5 if self.kind == "cn" else 1.0
        shift = 1.0 / (h * theta)
        if imp.mass is not None:
            # a (possibly singular) mass matrix cannot be folded into Z: keep the explicit half as an affine term
            aff = None
            if theta < 1.0:
                aff = torch.empty_like(u)
                self.ops.lincomb(aff, None, 0.0, [cb.f(t, u)], [(1.0 - theta) / theta])
            unew = imp.solve(t + h, u, shift, u, aff=aff)
            return unew, ([u, unew],), None
        if theta < 1.0:
            Z = torch.empty_like(u)
            self.ops.lincomb(Z, u, 1.0, [cb.f(t, u)], [h * (1.0 - theta)])
        else:
            Z = u
        unew = imp.solve(t + h, Z, shift, u)
        return unew, ([u, unew],), None

    # -- adjoint ----------------------------------------------------------------------------------------------------
    def adjoint_steps(self, cb_ex, cb_im, imp, nsteps, lam, mu, np_im):
        for cb in (cb_ex, cb_im):
            if cb is not None and hasattr(cb, "begin"):
                cb.begin(False, comm=self.comm)
        for _ in range(nsteps):
            if not self.traj:
                raise Error(-30, "adjoint requested more steps than the trajectory holds")
            if self.traj[-1][2] is None:
                self._restore(cb_ex, cb_im, imp, len(self.traj) - 1)
            t, h, stages, _ = self.traj.pop()
            if self.kind == "rk":
                lam = self._rk_adjoint(cb_ex, t, h, stages[0], lam, mu)
            elif self.kind == "arkimex":
                lam = self._ark_adjoint(cb_ex, cb_im, imp, t, h, stages[0], lam, mu, np_im)
            else:
                lam = self._theta_adjoint(cb_im, imp, t, h, stages[0], lam, mu)
        return lam

    def _rk_adjoint(self, cb, t, h, Y, lam, mu):
        sc, ops = self.scheme, self.ops
        s = sc.s
        vu = [None] * s   # J^T w of each stage;  lambda_{s,i} = coef[i] * vu[i] (the scaling is folded into later sums)
        coef = [0.0] * s
        for i in range(s - 1, -1, -1):
            if sc.fsal and i == s - 1:
                continue
            later = [j for j in range(i + 1, s) if sc.A[j][i] != 0.0 and vu[j] is not None]
            if sc.b[i] != 0.0 and not later:
                w = lam  # last stage: w = lambda itself (the VJP only reads it): no copy
                coef[i] = h * sc.b[i]
            elif sc.b[i] != 0.0:
                w = torch.empty_like(lam)
                ops.lincomb(w, lam, 1.0, [vu[j] for j in later], [sc.A[j][i] / sc.b[i] * coef[j] for j in later])
                coef[i] = h * sc.b[i]
            else:
                w = torch.empty_like(lam)
                ops.lincomb(w, None, 0.0, [vu[j] for j in later], [sc.A[j][i] * coef[j] for j in later])
                coef[i] = h
            if hasattr(cb, "vjp_accumulate"):  # RHS evaluator that adds coef * Jp^T w into mu itself
                vu[i], gp = cb.vjp_accumulate(t + sc.c[i] * h, Y[i], w, mu, coef[i])
            else:
                vu[i], gp = cb.vjp(t + sc.c[i] * h, Y[i], w)
            if gp is not None:
                ops.multi_axpy(mu, gp, cb.sizes, coef[i])
        live = [i for i in range(s) if vu[i] is not None]
        lam_n = torch.empty_like(lam)
        ops.lincomb(lam_n, lam, 1.0, [vu[i] for i in live], [coef[i] for i in live])
        return lam_n

    def _ark_adjoint(self, cb_ex, cb_im, imp, t, h, Y, lam, mu, np_im):
        sc, ops = self.scheme, self.ops
        s = sc.s
        ls = [None] * s
        mu_im = mu[:np_im] if np_im > 0 else None
        mu_ex = mu[np_im:]
        for i in range(s - 1, -1, -1):
            later_t = [j for j in range(i + 1, s) if sc.At[j][i] != 0.0]
            later_e = [j for j in range(i + 1, s) if sc.A[j][i] != 0.0]
            om = torch.empty_like(lam)
            ep = torch.empty_like(lam)
            ops.lincomb(om, lam, sc.bt[i], [ls[j] for j in later_t], [sc.At[j][i] for j in later_t])
            ops.lincomb(ep, lam, sc.b[i], [ls[j] for j in later_e], [sc.A[j][i] for j in later_e])
            parts = []
            if cb_ex is not None:
                if hasattr(cb_ex, "vjp_accumulate"):  # evaluator that adds h * Jp^T eps into mu itself
                    vu_e, gp_e = cb_ex.vjp_accumulate(t + sc.c[i] * h, Y[i], ep, mu_ex, h)
                else:
                    vu_e, gp_e = cb_ex.vjp(t + sc.c[i] * h, Y[i], ep)
                if gp_e is not None:
                    ops.multi_axpy(mu_ex, gp_e, cb_ex.sizes, h)
                parts.append(vu_e)
            gp_i0 = None
            if cb_im is not None:
                vu_i, gp_i0 = cb_im.vjp(t + sc.ct[i] * h, Y[i], om, want_params=cb_im.nparams > 0)
                parts.append(vu_i)
            r = torch.empty_like(lam)
            if cb_im is None or sc.At[i][i] == 0.0:
                # no implicit function registered: the stage equation is shift (Y - Z) = 0, whose transposed solve is h r
                ops.lincomb(r, None, 0.0, parts, [h] * len(parts))
                ls[i] = r
                w_im = om
            else:
                ops.lincomb(r, None, 0.0, parts, [1.0 / sc.At[i][i]] * len(parts))
                ls[i] = imp.solve_transpose(t + sc.ct[i] * h, Y[i], 1.0 / (h * sc.At[i][i]), r)
                w_im = torch.empty_like(lam)
                ops.lincomb(w_im, om, 1.0, [ls[i]], [sc.At[i][i]])
            if cb_im is not None and cb_im.nparams > 0:
                if w_im is om:
                    gp_i = gp_i0
                else:  # the j = i term needs the post-solve lambda_{s,i}
                    _, gp_i = cb_im.vjp(t + sc.ct[i] * h, Y[i], w_im, want_u=False)
                ops.multi_axpy(mu_im, gp_i, cb_im.sizes, h)
        lam_n = torch.empty_like(lam)
        ops.lincomb(lam_n, lam, 1.0, ls, [1.0] * s)
        return lam_n

    def _theta_adjoint(self, cb, imp, t, h, Y, lam, mu):
        ops = self.ops
        theta = 0.5 if self.kind == "cn" else 1.0
        u0, u1 = Y
        shift = 1.0 / (h * theta)
        c = (1.0 - theta) / theta
        # w = (shift M - J(u1))^-T lambda ;  lambda_n = shift M^T w + c J(u0)^T w ;  mu += Jp(u1)^T w + c Jp(u0)^T w
        w = imp.solve_transpose(t + h, u1, shift, lam)
        _, gp1 = cb.vjp(t + h, u1, w, want_u=False)
        ops.multi_axpy(mu, gp1, cb.sizes, 1.0)
        lam_n = torch.empty_like(lam)
        mw = imp._mass_times(w, transpose=True)
        if theta < 1.0:
            vu0, gp0 = cb.vjp(t, u0, w)
            ops.multi_axpy(mu, gp0, cb.sizes, c)
            ops.lincomb(lam_n, mw, shift, [vu0], [c])
        else:
            ops.lincomb(lam_n, mw, shift, [], [])
        return lam_n
