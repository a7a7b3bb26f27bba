"""ORACLE (test infrastructure, NOT product code; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this).

CPU restatement, in plain PyTorch, of the arithmetic that pnode delegates to PETSc TS / TSAdapt / TSTrajectory /
TSAdjoint, organised exactly like the reference executes it: a host-side stage loop, one `func(t, u)` call per stage
(/root/reference/pnode/petsc_adjoint.py:393-412), one `torch.autograd.grad` re-evaluating `func` per adjoint stage
(petsc_adjoint.py:52-82), one AXPY per tableau term, stage checkpoints kept in host memory.

PETSc itself is NOT under /root/reference and is unpinned (setup.py:9; .github/workflows/build.sh:4 clones `main`), so
the PETSc-internal parts follow the published algorithms as restated in SURVEY.md appendix A:
  A.1 TSStep_RK            -> `_rk_attempt`
  A.2 TSSolve / MATCHSTEP  -> `OracleTS.solve`, `_matchstep`
  A.3 TSAdaptChoose_Basic  -> `_adapt_basic`, `wrms_norm`
  A.4 TSAdjointStep_RK     -> `_rk_adjoint_step`
  A.5 TSTheta (+adjoint)   -> `_theta_attempt`, `_theta_adjoint_step`
  A.6 TSARKIMEX (+adjoint) -> `_ark_attempt`, `_ark_adjoint_step`
  A.7 TSTrajectory memory  -> `self.traj`

PARITY PIN: the forward path is pinned on the three known-answer values of /root/reference/tests/test_pnode.py:151-152,
179-180, 200-201 (tests/test_oracle_golden.py).  Gradients, adaptive step sequences and fp32 are UNPINNED by any
reference test (the reference prints gradients but asserts nothing, test_pnode.py:149-150); for those this oracle is
checked against `torch.autograd` through the unrolled scheme (tests/test_oracle_adjoint.py), which is the exact
reverse-mode derivative that a discrete adjoint must reproduce.
"""
import math

import torch

from . import tableaux as tbx

# ----------------------------------------------------------------------------------------------------------------------
# options database: the subset of PETSc's that the reference's scripts ever set ("Config / flags")


def parse_petsc_options(argv):
    """`petsc4py.init(sys.argv)` hands every `-key [value]` pair to PETSc's options DB (ode_demo_petsc.py:65-66)."""
    opts = {}
    i = 0
    argv = list(argv or [])
    while i < len(argv):
        a = argv[i]
        if isinstance(a, str) and a.startswith("-") and len(a) > 1 and not a[1].isdigit() and not a.startswith("--"):
            key = a[1:]
            if i + 1 < len(argv) and not (str(argv[i + 1]).startswith("-") and not _is_number(argv[i + 1])):
                opts[key] = str(argv[i + 1])
                i += 2
            else:
                opts[key] = None
                i += 1
        else:
            i += 1
    return opts


def _is_number(s):
    try:
        float(s)
        return True
    except (TypeError, ValueError):
        return False


# ----------------------------------------------------------------------------------------------------------------------
# TSAdapt pieces (appendix A.2 / A.3)

ADAPT_SAFETY = 0.9
ADAPT_REJECT_SAFETY = 0.5
ADAPT_CLIP = (0.1, 10.0)
ADAPT_DT_MIN = 1e-20
ADAPT_DT_MAX = 1e50
MATCHSTEP_FAC = (0.01, 2.0)
MAX_REJECT = 10
TSPAN_RELTOL = 1e-6
TSPAN_ABSTOL = 10 * 2.220446049250313e-16


def wrms_norm(u, y, atol, rtol):
    """TSErrorWeightedNorm, NORM_2: sqrt(mean(((u-y)/(atol+rtol*max(|u|,|y|)))^2)) over the WHOLE flattened state."""
    u = u.reshape(-1).double()
    y = y.reshape(-1).double()
    tol = atol + rtol * torch.maximum(u.abs(), y.abs())
    return math.sqrt(float((((u - y) / tol) ** 2).sum()) / u.numel())


def _adapt_basic(h, enorm, order, prev_accepted):
    """TSAdaptChoose_Basic: returns (accept, h_next) before the MATCHSTEP clamp."""
    safety = ADAPT_SAFETY
    if enorm > 1.0:
        if not prev_accepted:
            safety *= ADAPT_REJECT_SAFETY
        accept = h < (1 + 1.4901161193847656e-08) * ADAPT_DT_MIN
    else:
        accept = True
    hfac = safety * enorm ** (-1.0 / order) if enorm > 0 else math.inf
    hfac = min(max(hfac, ADAPT_CLIP[0]), ADAPT_CLIP[1])
    h_next = min(max(h * hfac, ADAPT_DT_MIN), ADAPT_DT_MAX)
    return accept, h_next


def _close(a, b, tol):
    return abs(a - b) <= tol


# ----------------------------------------------------------------------------------------------------------------------


class OracleTS:
    """One PETSc `TS` object: scheme, step controller, trajectory, cost gradients."""

    def __init__(self, options=None):
        self.opt = dict(options or {})
        self.kind = "rk"
        self.scheme = "3bs"
        self.atol = float(self.opt.get("ts_atol", 1e-4))
        self.rtol = float(self.opt.get("ts_rtol", 1e-4))
        self.max_reject = int(self.opt.get("ts_max_reject", MAX_REJECT))
        self.traj = []
        self.log = []  # (t, h, accepted, enorm) per attempt
        self.dt_span_cached = 0.0
        self.nfe = 0

    # -- scheme selection -------------------------------------------------------------------------------------------
    def set_method(self, method):
        self.kind, self.scheme = tbx.METHOD_TO_SCHEME.get(method, tbx.DEFAULT_SCHEME)

    def set_from_options(self):
        """TSSetFromOptions: command-line -ts_* beats the method= argument (petsc_adjoint.py:775)."""
        o = self.opt
        if "ts_type" in o:
            kind = o["ts_type"]
            if kind == "euler":
                self.kind, self.scheme = "rk", "1fe"
            elif kind == "rk":
                if self.kind != "rk":
                    self.kind, self.scheme = "rk", "3bs"
            elif kind == "arkimex":
                if self.kind != "arkimex":
                    self.kind, self.scheme = "arkimex", "3"
            elif kind in ("cn", "beuler"):
                self.kind, self.scheme = kind, None
            else:
                raise ValueError("oracle: unsupported -ts_type %s" % kind)
        if self.kind == "rk" and "ts_rk_type" in o:
            self.scheme = o["ts_rk_type"]
        if self.kind == "arkimex" and "ts_arkimex_type" in o:
            self.scheme = o["ts_arkimex_type"]
        self.atol = float(o.get("ts_atol", self.atol))
        self.rtol = float(o.get("ts_rtol", self.rtol))

    @property
    def tableau(self):
        if self.kind == "rk":
            return tbx.RK[self.scheme]
        if self.kind == "arkimex":
            return tbx.ARK[self.scheme]
        return None

    def adaptive(self):
        if self.opt.get("ts_adapt_type", "basic") == "none":
            return False
        if self.kind in ("cn", "beuler"):
            return False  # TSTheta switches its TSAdapt to none unless -ts_theta_adapt
        return self.tableau.bembed is not None

    # -- MATCHSTEP (A.2) ---------------------------------------------------------------------------------------------
    def _matchstep(self, t_new, h, h_next, span, ctr, max_time):
        """Clamp h_next so that the next step lands on the next target.  `t_new` is the time after the step just
        accepted, `ctr` the index of the next un-hit span point."""
        a = 1.0 + MATCHSTEP_FAC[0]
        b = MATCHSTEP_FAC[1]
        hit = False
        if span is not None:
            if _close(t_new, span[ctr], TSPAN_RELTOL * h + TSPAN_ABSTOL):
                hit = True
                tmax = span[ctr + 1] if ctr + 1 < len(span) else max_time
            else:
                tmax = span[ctr]
        else:
            tmax = max_time
        tend = t_new + h_next
        hmax = tmax - t_new
        out = h_next
        if t_new < tmax and tend > tmax:
            out = hmax
        if t_new < tmax and tend < tmax and out * b > hmax:
            out = hmax / 2
        if t_new < tmax and tend < tmax and h_next * a > hmax:
            out = hmax
        if span is not None:
            if h != out and not self.dt_span_cached:
                self.dt_span_cached = h
            if h == out and self.dt_span_cached and hit:
                out = self.dt_span_cached
                self.dt_span_cached = 0.0
        return out

    # -- forward ----------------------------------------------------------------------------------------------------
    def solve(self, ctx, u0, t0, span, max_time, h0, poststep=None, save_trajectory=True):
        """TSSolve.  ctx supplies f_ex / f_im / implicit_solve.  Returns (u_final, span_solutions)."""
        self.traj = []
        self.log = []
        self.dt_span_cached = 0.0
        u = u0.clone()
        t = float(t0)
        h = float(h0)
        step = 0
        ctr = 0
        sols = []
        if span is not None:
            sols.append(u.clone())
            ctr = 1
        adaptive = self.adaptive()
        k_fsal = None
        tend_all = span[-1] if span is not None else max_time
        # TSSolve, before the first step: with MATCHSTEP the initial step may not overshoot the first target
        first = span[1] if span is not None and len(span) > 1 else max_time
        if h >= first - t or _close(h, first - t, TSPAN_ABSTOL):
            h = first - t
        while t < tend_all and not _close(t, tend_all, TSPAN_ABSTOL):
            prev_accepted = True
            rejections = 0
            while True:
                if self.kind == "rk":
                    u_new, stages, err = self._rk_attempt(ctx, t, u, h, k_fsal, adaptive)
                elif self.kind == "arkimex":
                    u_new, stages, err = self._ark_attempt(ctx, t, u, h, adaptive)
                else:
                    u_new, stages, err = self._theta_attempt(ctx, t, u, h)
                if adaptive:
                    enorm = wrms_norm(u_new, u_new + err, self.atol, self.rtol)
                    accept, h_next = _adapt_basic(h, enorm, self.tableau.order, prev_accepted)
                else:
                    enorm, accept, h_next = -1.0, True, h
                self.log.append((t, h, accept, enorm))
                if accept:
                    h_next = self._matchstep(t + h, h, h_next, span, min(ctr, len(span) - 1) if span is not None else 0,
                                             max_time)
                    break
                h = h_next
                prev_accepted = False
                k_fsal = None  # a rolled-back step disables FSAL reuse; k_0 is recomputed (same value)
                rejections += 1
                if rejections > self.max_reject:
                    raise RuntimeError("TS_DIVERGED_STEP_REJECTED")
            if save_trajectory:
                self.traj.append({"t": t, "h": h, "u": u, "stages": stages})
            if self.kind == "rk" and self.tableau.fsal:
                k_fsal = stages["K"][-1]
            t_old = t
            t = t + h
            u = u_new
            step += 1
            h_used = h
            h = h_next
            if poststep is not None:
                h_override = poststep(step, t)
                if h_override is not None:
                    h = h_override
            if span is not None and ctr < len(span) and _close(t, span[ctr], TSPAN_RELTOL * h_used + TSPAN_ABSTOL):
                sols.append(u.clone())
                ctr += 1
            del t_old
        self.h_last = h
        self.steps = step
        return u, sols

    def _rk_attempt(self, ctx, t, u, h, k_fsal, want_err):
        tab = self.tableau
        A, b, be, c = tab.floats()
        s = tab.s
        Y, K = [], []
        for i in range(s):
            y = u.clone()
            for j in range(i):
                if A[i][j] != 0.0:
                    y = y + (h * A[i][j]) * K[j]  # one AXPY per tableau term (VecMAXPY)
            Y.append(y)
            if i == 0 and k_fsal is not None:
                K.append(k_fsal)
            else:
                K.append(ctx.f_ex(t + c[i] * h, y))
                self.nfe += 1
        u_new = u.clone()
        for j in range(s):
            if b[j] != 0.0:
                u_new = u_new + (h * b[j]) * K[j]
        err = None
        if want_err:
            err = torch.zeros_like(u)
            for j in range(s):
                w = be[j] - b[j]
                if w != 0.0:
                    err = err + (h * w) * K[j]
        return u_new, {"Y": Y, "K": K}, err

    def _ark_attempt(self, ctx, t, u, h, want_err):
        tab = self.tableau
        At, A, bt, b, be, ct, c = tab.floats()
        s = tab.s
        Y, KI, KE = [], [], []
        for i in range(s):
            Z = u.clone()
            for j in range(i):
                if At[i][j] != 0.0:
                    Z = Z + (h * At[i][j]) * KI[j]
                if A[i][j] != 0.0:
                    Z = Z + (h * A[i][j]) * KE[j]
            if At[i][i] == 0.0:
                y = Z
                ki = ctx.f_im(t + ct[i] * h, y)
            else:
                shift = 1.0 / (h * At[i][i])
                guess = Y[i - 1] if i > 0 else u
                y = ctx.implicit_solve(t + ct[i] * h, Z, shift, guess)
                ki = shift * (y - Z)
            Y.append(y)
            KI.append(ki)
            KE.append(ctx.f_ex(t + c[i] * h, y))
        u_new = u.clone()
        for j in range(s):
            if bt[j] != 0.0:
                u_new = u_new + (h * bt[j]) * KI[j]
            if b[j] != 0.0:
                u_new = u_new + (h * b[j]) * KE[j]
        err = None
        if want_err:
            err = torch.zeros_like(u)
            for j in range(s):
                w = be[j] - b[j]
                if w != 0.0:
                    err = err + (h * w) * (KI[j] + KE[j])
        return u_new, {"Y": Y}, err

    def _theta_attempt(self, ctx, t, u, h):
        theta = 0.5 if self.kind == "cn" else 1.0
        shift = 1.0 / (h * theta)
        if getattr(ctx, "mass", None) is not None:
            # M (u1 - u)/h = (1-theta) f(t,u) + theta f(t+h,u1), M possibly singular (evalIFunction, petsc_adjoint.py:426-431)
            aff = ((1.0 - theta) / theta) * ctx.f_im(t, u) if theta < 1.0 else None
            u_new = ctx.implicit_solve(t + h, u, shift, u, aff=aff)
            return u_new, {"Y": [u, u_new]}, None
        # endpoint form: u1 = u + h[(1-theta) f(t,u) + theta f(t+h,u1)]
        if theta < 1.0:
            Z = u + (h * (1.0 - theta)) * ctx.f_im(t, u)
        else:
            Z = u.clone()
        u_new = ctx.implicit_solve(t + h, Z, shift, u)
        return u_new, {"Y": [u, u_new]}, None

    # -- adjoint ----------------------------------------------------------------------------------------------------
    def adjoint_steps(self, ctx, nsteps, lam, mu):
        """TSAdjointSolve after TSAdjointSetSteps(nsteps): pops `nsteps` checkpoints off the trajectory."""
        for _ in range(nsteps):
            cp = self.traj.pop()
            if self.kind == "rk":
                lam, mu = self._rk_adjoint_step(ctx, cp, lam, mu)
            elif self.kind == "arkimex":
                lam, mu = self._ark_adjoint_step(ctx, cp, lam, mu)
            else:
                lam, mu = self._theta_adjoint_step(ctx, cp, lam, mu)
        return lam, mu

    def _rk_adjoint_step(self, ctx, cp, lam, mu):
        tab = self.tableau
        A, b, _, c = tab.floats()
        s = tab.s
        t, h, Y = cp["t"], cp["h"], cp["stages"]["Y"]
        ls = [None] * s
        for i in range(s - 1, -1, -1):
            if tab.fsal and i == s - 1:
                ls[i] = torch.zeros_like(lam)
                continue
            if b[i] != 0.0:
                w = lam.clone()
                for j in range(i + 1, s):
                    if A[j][i] != 0.0:
                        w = w + (A[j][i] / b[i]) * ls[j]
                coef = h * b[i]
            else:
                w = torch.zeros_like(lam)
                for j in range(i + 1, s):
                    if A[j][i] != 0.0:
                        w = w + A[j][i] * ls[j]
                coef = h
            vu, vp = ctx.vjp_ex(t + c[i] * h, Y[i], w)
            ls[i] = coef * vu
            mu = mu + coef * vp
        lam_n = lam.clone()
        for i in range(s):
            lam_n = lam_n + ls[i]
        return lam_n, mu

    def _ark_adjoint_step(self, ctx, cp, lam, mu):
        tab = self.tableau
        At, A, bt, b, _, ct, c = tab.floats()
        s = tab.s
        t, h, Y = cp["t"], cp["h"], cp["stages"]["Y"]
        ls = [None] * s
        for i in range(s - 1, -1, -1):
            om = bt[i] * lam
            ep = b[i] * lam
            for j in range(i + 1, s):
                if At[j][i] != 0.0:
                    om = om + At[j][i] * ls[j]
                if A[j][i] != 0.0:
                    ep = ep + A[j][i] * ls[j]
            vu_e, vp_e = ctx.vjp_ex(t + c[i] * h, Y[i], ep)
            vu_i, _ = ctx.vjp_im(t + ct[i] * h, Y[i], om)
            r = vu_i + vu_e
            if At[i][i] == 0.0:
                ls[i] = h * r
                w_im = om
            else:
                shift = 1.0 / (h * At[i][i])
                ls[i] = ctx.implicit_solve_transpose(t + ct[i] * h, Y[i], shift, r / At[i][i])
                w_im = om + At[i][i] * ls[i]
            _, vp_i = ctx.vjp_im(t + ct[i] * h, Y[i], w_im)
            mu = mu + h * ctx.pad_params(vp_i, vp_e)
        lam_n = lam.clone()
        for i in range(s):
            lam_n = lam_n + ls[i]
        return lam_n, mu

    def _theta_adjoint_step(self, ctx, cp, lam, mu):
        theta = 0.5 if self.kind == "cn" else 1.0
        t, h = cp["t"], cp["h"]
        u0, u1 = cp["stages"]["Y"]
        shift = 1.0 / (h * theta)
        if getattr(ctx, "mass", None) is not None:
            c = (1.0 - theta) / theta
            w = ctx.implicit_solve_transpose(t + h, u1, shift, lam)
            _, vp1 = ctx.vjp_im(t + h, u1, w)
            mu = mu + ctx.pad_params(vp1, None)
            lam_n = shift * (ctx.mass.T @ w.reshape(-1)).reshape(w.shape)
            if theta < 1.0:
                vu0, vp0 = ctx.vjp_im(t, u0, w)
                lam_n = lam_n + c * vu0
                mu = mu + c * ctx.pad_params(vp0, None)
            return lam_n, mu
        ls = ctx.implicit_solve_transpose(t + h, u1, shift, shift * lam)
        _, vp1 = ctx.vjp_im(t + h, u1, ls)
        mu = mu + (h * theta) * ctx.pad_params(vp1, None)
        lam_n = ls.clone()
        if theta < 1.0:
            vu0, vp0 = ctx.vjp_im(t, u0, ls)
            lam_n = lam_n + (h * (1.0 - theta)) * vu0
            mu = mu + (h * (1.0 - theta)) * ctx.pad_params(vp0, None)
        return lam_n, mu
