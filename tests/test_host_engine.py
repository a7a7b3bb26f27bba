"""HOST logic of the product (ODEPetsc option handling, stage loops, step controller, adjoint recurrences) against the
oracle, with the device kernels replaced by the torch-CPU test double of tests/_fake_ops.py.  The GPU parity tests
(test_gpu_*.py) run the same comparisons through the real C-ABI kernels."""
import copy

import pytest
import torch

from oracle import OracleODEPetsc
from pnode_b200.options import Options
from _fake_ops import patch_cpu
from _problems import (PETSC_ARGS, ROBER_STEPS, ROBER_T, Rober, RoberEX, RoberIM, SpiralFunc, TimeMLP, rel_err,
                       rober_truth, spiral_inputs)


def _both(monkeypatch, argv, setup_kw, funcs, u0, t, gout, step):
    pa = patch_cpu(monkeypatch)
    Options.insert_args(argv)
    res = []
    for make in (lambda: OracleODEPetsc(argv), lambda: pa.ODEPetsc()):
        fs = [copy.deepcopy(f) for f in funcs]
        kw = dict(setup_kw)
        if len(fs) == 2:
            kw["func2"] = fs[1]
        ode = make()
        ode.setupTS(u0, fs[0], step_size=step, enable_adjoint=True, **kw)
        y0 = u0.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t)
        (out * gout).sum().backward()
        grads = [p.grad.clone() for f in fs for p in f.parameters()]
        res.append((out.detach(), y0.grad.clone(), grads, ode))
    return res


def _assert_close(a, b, tol):
    assert rel_err(a[0], b[0]) < tol
    assert rel_err(a[1], b[1]) < tol
    assert len(a[2]) == len(b[2])
    for x, y in zip(a[2], b[2]):
        assert rel_err(x, y) < tol


@pytest.mark.parametrize("scheme", ["5bs", "5f", "3", "2a"])
def test_fixed_step_rk_types_from_the_option(monkeypatch, scheme):
    u0, t, gout = spiral_inputs(20)
    o, p = _both(monkeypatch, ["-ts_adapt_type", "none", "-ts_rk_type", scheme], dict(method="rk4"),
                 [SpiralFunc(bias_std=0.1)], u0, t, gout, 0.025)
    _assert_close(p, o, 1e-13)


def test_adaptive_bogacki_shampine_54(monkeypatch):
    u0, t, gout = spiral_inputs(20)
    o, p = _both(monkeypatch, ["-ts_rk_type", "5bs", "-ts_rtol", "1e-7", "-ts_atol", "1e-7"], dict(method="rk4"),
                 [SpiralFunc(bias_std=0.1)], u0, t, gout, 0.025)
    _assert_close(p, o, 1e-12)
    assert [a[2] for a in p[3]._loop.attempts] == [a[2] for a in o[3].ts.log]


@pytest.mark.parametrize("method", ["euler", "rk2", "bosh3", "rk4", "dopri5", "midpoint"])
def test_fixed_step_rk(monkeypatch, method):
    u0, t, gout = spiral_inputs(20)
    o, p = _both(monkeypatch, ["-ts_adapt_type", "none"], dict(method=method), [SpiralFunc(bias_std=0.1)], u0, t, gout, 0.025)
    _assert_close(p, o, 1e-13)
    assert [a[:3] for a in p[3]._loop.attempts] == [a[:3] for a in o[3].ts.log]


def test_adaptive_dopri5_identical_step_sequence(monkeypatch):
    func = TimeMLP(d=6, hidden=16)
    g = torch.Generator().manual_seed(5)
    u0 = torch.randn(50, 6, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.4, 1.0], dtype=torch.float64)
    gout = torch.randn(3, 50, 6, generator=g, dtype=torch.float64)
    o, p = _both(monkeypatch, ["-ts_rtol", "1e-6", "-ts_atol", "1e-6"], dict(method="dopri5"), [func], u0, t, gout, 0.3)
    log_o, log_p = o[3].ts.log, p[3]._loop.attempts
    assert len(log_o) == len(log_p) and len(log_o) > 4
    assert any(not a[2] for a in log_o), "the case must contain at least one rejected attempt"
    for a, b in zip(log_o, log_p):
        assert a[2] == b[2]
        assert a[0] == pytest.approx(b[0], rel=1e-12, abs=1e-15) and a[1] == pytest.approx(b[1], rel=1e-12)
        assert a[3] == pytest.approx(b[3], rel=1e-9)
    _assert_close(p, o, 1e-12)


def test_adaptive_bosh3_and_step_restore_after_span_point(monkeypatch):
    func = TimeMLP(d=4, hidden=8)
    g = torch.Generator().manual_seed(6)
    u0 = torch.randn(10, 4, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.25, 0.5, 0.75], dtype=torch.float64)
    gout = torch.randn(4, 10, 4, generator=g, dtype=torch.float64)
    o, p = _both(monkeypatch, [], dict(method="bosh3"), [func], u0, t, gout, 0.1)
    assert [(a[2]) for a in o[3].ts.log] == [(a[2]) for a in p[3]._loop.attempts]
    _assert_close(p, o, 1e-12)
    # fixed step that does not divide the output spacing: 0.1, then 0.15 left < 2h => halved to 0.075 + 0.075
    # (MATCHSTEP), then the un-shortened 0.1 is restored after the span point
    o, p = _both(monkeypatch, ["-ts_adapt_type", "none"], dict(method="rk4"), [func], u0, t, gout, 0.1)
    hs = [a[1] for a in p[3]._loop.attempts]
    assert hs[:4] == pytest.approx([0.1, 0.075, 0.075, 0.1])
    assert [a[:2] for a in o[3].ts.log] == pytest.approx([a[:2] for a in p[3]._loop.attempts])
    _assert_close(p, o, 1e-13)


def test_rober_goldens_through_product_host_logic(monkeypatch):
    true_y = rober_truth()
    cases = [(dict(method="cn", implicit_form=True), [Rober()], 1.8492e-6),
             (dict(method="imex", implicit_form=True, imex_form=True), [RoberIM(), RoberEX()], 3.1138e-6),
             (dict(method="rk3"), [Rober()], 1.8495e-6)]
    for kw, funcs, golden in cases:
        gout = torch.sign(torch.ones_like(true_y))
        o, p = _both(monkeypatch, PETSC_ARGS, kw, funcs, true_y[0], ROBER_T, gout, ROBER_STEPS)
        loss = torch.mean(torch.abs(p[0] - true_y)).item()
        assert loss == pytest.approx(golden, rel=2e-4)
        _assert_close(p, o, 1e-9)
        Options.clear_all()


@pytest.mark.parametrize("name", ["l2", "3", "4", "5", "ars122", "a2", "1bee", "2c", "2d", "2e", "prssp2", "bpr3", "ars443"])
def test_imex_batched_linear_solver(monkeypatch, name):
    from test_oracle_adjoint import LinearIM

    N, B = 8, 6
    g = torch.Generator().manual_seed(7)
    u0 = torch.randn(B, N, generator=g, dtype=torch.float64) * 0.5
    t = torch.tensor([0.0, 0.2, 0.4], dtype=torch.float64)
    gout = torch.randn(3, B, N, generator=g, dtype=torch.float64)
    argv = ["-ts_adapt_type", "none", "-snes_type", "ksponly", "-ts_arkimex_type", name]
    o, p = _both(monkeypatch, argv, dict(method="imex", imex_form=True, batch_size=B, linear_solver="torch"),
                 [LinearIM(N), TimeMLP(d=N, hidden=12)], u0, t, gout, 0.1)
    _assert_close(p, o, 1e-11)


def test_options_override_method_and_resetup_semantics(monkeypatch):
    pa = patch_cpu(monkeypatch)
    u0, t, gout = spiral_inputs(5)
    f = SpiralFunc()
    Options.insert_args(["-ts_adapt_type", "none", "-ts_type", "rk", "-ts_rk_type", "4"])
    ode = pa.ODEPetsc()
    ode.setupTS(u0, f, step_size=0.025, method="euler")  # command line wins (petsc_adjoint.py:775)
    assert ode._scheme.name == "4"
    Options.clear_all()
    Options.insert_args(["-ts_adapt_type", "none"])
    ode = pa.ODEPetsc()
    ode.setupTS(u0, f, step_size=0.025, method="rk4")
    assert ode._scheme.name == "4"
    ode.setupTS(u0, f, step_size=0.0125, method="euler", enable_adjoint=False)  # same meta: method NOT re-applied (C.3)
    assert ode._scheme.name == "4" and ode.step_size == 0.0125 and not ode.enable_adjoint
    ode.setupTS(u0[:3], f, step_size=0.025, method="euler")  # shape changed: method applied
    assert ode._scheme.name == "1fe"
    ode.setupTS(u0, f, method="no_such_method")  # silently the TS default (C.1)
    assert ode._scheme.name == "3bs"
    with pytest.raises(ValueError):
        ode.setupTS(u0, f, imex_form=True)
    ode2 = pa.ODEPetsc()
    ode2.setupTS(u0, lambda t, y: y, method="rk4")
    with pytest.raises(ValueError):
        ode2.odeint_adjoint(u0, t)  # func must be an nn.Module (petsc_adjoint.py:896-897)


def test_missed_output_point_raises(monkeypatch):
    pa = patch_cpu(monkeypatch)
    Options.insert_args(["-ts_adapt_type", "none"])
    u0, _, _ = spiral_inputs(3)
    ode = pa.ODEPetsc()
    # per-step list that walks past the second output time without landing near it
    ode.setupTS(u0, SpiralFunc(), step_size=[0.1, 0.1, 0.1], method="euler")
    t = torch.tensor([0.0, 0.1, 0.25 + 1e-4, 0.3], dtype=torch.float64)
    # the list entry set in PostStep overrides the MATCHSTEP clamp (petsc_adjoint.py:523-525), the run steps over
    # t=0.2501 and the reference's sanity check fires (petsc_adjoint.py:867-868)
    with pytest.raises(Exception, match="fails to step on all the specified points"):
        ode.odeint(u0, t)


def test_single_time_returns_leading_axis_one(monkeypatch):
    pa = patch_cpu(monkeypatch)
    Options.insert_args(["-ts_adapt_type", "none"])
    u0, _, _ = spiral_inputs(4)
    ode = pa.ODEPetsc()
    ode.setupTS(u0, SpiralFunc(), step_size=0.05, method="rk4", enable_adjoint=False)
    out = ode.odeint(u0, torch.tensor([0.2], dtype=torch.float64))
    assert out.shape == (1, 4, 1, 2)
    assert len(ode._loop.attempts) == 4 and ode._loop.attempts[0][0] == 0.0


@pytest.mark.parametrize("extra,expect_recompute", [(["-ts_trajectory_solution_only", "1"], 12),
                                                    (["-ts_trajectory_max_cps_ram", "3"], None),
                                                    (["-ts_trajectory_max_cps_ram", "1"], None)])
def test_bounded_checkpoint_storage_gives_identical_gradients(monkeypatch, extra, expect_recompute):
    """-ts_trajectory_solution_only / -ts_trajectory_max_cps_ram trade HBM for recomputation; the
    arithmetic is unchanged, so trajectories and gradients must be IDENTICAL to the store-everything run."""
    pa = patch_cpu(monkeypatch)
    func = TimeMLP(d=4, hidden=8)
    g = torch.Generator().manual_seed(6)
    u0 = torch.randn(10, 4, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.25, 0.5, 0.75], dtype=torch.float64)
    gout = torch.randn(4, 10, 4, generator=g, dtype=torch.float64)

    def run(argv):
        Options.clear_all()
        Options.insert_args(["-ts_adapt_type", "none"] + argv)
        f = copy.deepcopy(func)
        ode = pa.ODEPetsc()
        ode.setupTS(u0, f, step_size=0.0625, method="rk4", enable_adjoint=True)
        y0 = u0.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t)
        (out * gout).sum().backward()
        return out.detach(), y0.grad, [p.grad for p in f.parameters()], ode

    full = run([])
    lean = run(extra)
    assert torch.equal(full[0], lean[0]) and torch.equal(full[1], lean[1])
    assert all(torch.equal(a, b) for a, b in zip(full[2], lean[2]))
    assert lean[3]._engine.recomputed_steps >= 12 and full[3]._engine.recomputed_steps == 0
    if expect_recompute is not None:
        assert lean[3]._engine.recomputed_steps == expect_recompute


@pytest.mark.parametrize("method,ksp_rtol,tol", [("cn", "1e-10", 1e-7), ("beuler", "1e-10", 1e-7), ("cn", None, 1e-3)])
def test_matrix_free_newton_gmres_matches_dense_oracle(monkeypatch, method, ksp_rtol, tol):
    """implicit cn / beuler on a state too large for a dense Jacobian (n = 600): matrix-free Newton-GMRES
    (the reference's IJacShell + KSPGMRES) against the oracle's dense Newton; parity only to the Krylov tolerance."""
    func = SpiralFunc(bias_std=0.1)
    u0, _, gout = spiral_inputs(300)
    t = torch.tensor([0.0, 0.1, 0.2], dtype=torch.float64)
    argv = ["-ts_adapt_type", "none"] + (["-ksp_rtol", ksp_rtol] if ksp_rtol else [])
    o, p = _both(monkeypatch, argv, dict(method=method, implicit_form=True), [func], u0, t, gout[:3], 0.1)
    assert p[3]._imp.krylov_iterations > 0
    _assert_close(p, o, tol)


class _Pendulum(torch.nn.Module):
    """Index-1 pendulum DAE of examples-pnode/pendulum_DAE.py:108-117 with a trainable gravity constant."""

    def __init__(self):
        super().__init__()
        self.g = torch.nn.Parameter(torch.tensor(9.81, dtype=torch.float64))

    def forward(self, t, y):
        g = self.g
        return torch.stack((y[2], y[3], -y[0] * y[4], -y[1] * y[4] - g,
                            y[4] * (y[0] ** 2 + y[1] ** 2) + g * y[1] - (y[2] ** 2 + y[3] ** 2)))


@pytest.mark.parametrize("method", ["cn", "beuler"])
def test_mass_matrix_dae_theta_methods(monkeypatch, method):
    """mass= (M u' = f with singular M), evalIFunction petsc_adjoint.py:426-431, as used by pendulum_DAE.py:119-139."""
    M = torch.eye(5, dtype=torch.float64)
    M[-1, -1] = 0.0
    u0 = torch.tensor([1.0, 0.0, 0.0, 1.0, 1.0 - 0.0], dtype=torch.float64)  # consistent: lambda = |v|^2 - g y = 1
    t = torch.tensor([0.0, 0.02, 0.05], dtype=torch.float64)
    gout = torch.tensor([[0.3, -0.2, 0.5, 0.1, 0.0], [1.0, 0.5, -0.3, 0.2, 0.1], [0.7, -1.0, 0.4, 0.3, -0.2]],
                        dtype=torch.float64)
    o, p = _both(monkeypatch, ["-ts_adapt_type", "none"], dict(method=method, implicit_form=True, mass=M), [_Pendulum()],
                 u0, t, gout, 0.01)
    _assert_close(p, o, 1e-9)
    # the discrete adjoint of the DAE scheme is the derivative of the discrete map: finite-difference check on g
    base = (p[0] * gout).sum().item()
    import pnode_b200.petsc_adjoint as pa
    f2 = _Pendulum()
    with torch.no_grad():
        f2.g.add_(1e-6)
    ode = pa.ODEPetsc()
    ode.setupTS(u0, f2, step_size=0.01, method=method, implicit_form=True, mass=M, enable_adjoint=False)
    pert = (ode.odeint(u0, t) * gout).sum().item()
    assert (pert - base) / 1e-6 == pytest.approx(p[2][0].item(), rel=1e-4)


def _run_product(monkeypatch, argv, setup_kw, funcs, u0, t, gout, step, twice=False):
    pa = patch_cpu(monkeypatch)
    Options.clear_all()
    Options.insert_args(argv)
    fs = [copy.deepcopy(f) for f in funcs]
    kw = dict(setup_kw)
    if len(fs) == 2:
        kw["func2"] = fs[1]
    ode = pa.ODEPetsc()
    ode.setupTS(u0, fs[0], step_size=step, enable_adjoint=True, **kw)
    outs = []
    for _ in range(2 if twice else 1):
        for f in fs:
            f.zero_grad(set_to_none=True)
        y0 = u0.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t)
        (out * gout).sum().backward()
        outs.append((out.detach().clone(), y0.grad.clone(), [p.grad.clone() for f in fs for p in f.parameters()], ode, fs))
    return outs


def test_adjoint_reuses_the_forward_stage_graphs(monkeypatch):
    """Stage evaluations keep their autograd graph; the adjoint stage at the same point differentiates it instead of
    re-evaluating the module: same bits as the re-evaluating path (-pnode_reuse_graph 0), func.nfe still counts the reference's
    calls, graphs of rejected adaptive attempts are dropped, and a touched parameter invalidates what was kept."""
    func = TimeMLP(d=6, hidden=16)
    g = torch.Generator().manual_seed(5)
    u0 = torch.randn(50, 6, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.4, 1.0], dtype=torch.float64)
    gout = torch.randn(3, 50, 6, generator=g, dtype=torch.float64)
    argv = ["-ts_rtol", "1e-6", "-ts_atol", "1e-6"]
    (a,) = _run_product(monkeypatch, argv, dict(method="dopri5"), [func], u0, t, gout, 0.3)
    (b,) = _run_product(monkeypatch, argv + ["-pnode_reuse_graph", "0"], dict(method="dopri5"), [func], u0, t, gout, 0.3)
    cb_a, cb_b = a[3]._cb_ex, b[3]._cb_ex
    loop = a[3]._loop
    assert any(not x[2] for x in loop.attempts), "needs a rejected attempt"
    # dopri5: 6 adjoint stages per accepted step; the first stage of a step that inherited its slope from the previous step (FSAL)
    # was never evaluated at that step's own state tensor, so its adjoint stage re-evaluates
    carried = sum(1 for k in range(1, len(loop.attempts)) if loop.attempts[k][2] and loop.attempts[k - 1][2])
    assert cb_a.reused_graphs == 6 * loop.steps - carried and cb_b.reused_graphs == 0
    assert len(cb_a._graphs) <= 7 * loop.steps  # nothing kept from the rejected attempts
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and all(torch.equal(x, y) for x, y in zip(a[2], b[2]))
    assert cb_a.nvjp == cb_b.nvjp and cb_b.nfe == cb_a.nfe  # the engine's own counters do not depend on the reuse
    # a parameter modified between the forward and the backward invalidates the kept graphs (version counters)
    pa = patch_cpu(monkeypatch)
    Options.clear_all()
    Options.insert_args(["-ts_adapt_type", "none"])
    f = copy.deepcopy(func)
    ode = pa.ODEPetsc()
    ode.setupTS(u0, f, step_size=0.2, method="rk4", enable_adjoint=True)
    y0 = u0.clone().requires_grad_(True)
    out = ode.odeint_adjoint(y0, t)
    with torch.no_grad():
        next(f.parameters()).mul_(1.0)
    (out * gout).sum().backward()
    assert ode._cb_ex.reused_graphs == 0


def test_module_with_forward_side_effects_is_re_evaluated(monkeypatch):
    class Noisy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(4, 4).double()
            self.bn = torch.nn.BatchNorm1d(4).double()

        def forward(self, t, y):
            return self.bn(self.lin(y))

    g = torch.Generator().manual_seed(1)
    u0 = torch.randn(8, 4, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.2], dtype=torch.float64)
    gout = torch.randn(2, 8, 4, generator=g, dtype=torch.float64)
    (a,) = _run_product(monkeypatch, ["-ts_adapt_type", "none"], dict(method="rk4"), [Noisy()], u0, t, gout, 0.1)
    assert a[3]._cb_ex.reused_graphs == 0
    assert int(a[4][0].bn.num_batches_tracked) == 16  # 2 steps x 4 stages, forward + the adjoint's re-evaluation


def test_fixed_jacobian_keeps_the_factorisation_across_solves(monkeypatch):
    """fixed_jacobian=True (documented: constant across ODE solves): the block inverse survives odeint calls and setupTS calls
    until a parameter of the implicit function is touched; without the flag it is rebuilt per solve like the reference."""
    class Lin(torch.nn.Module):
        def __init__(self, n):
            super().__init__()
            gg = torch.Generator().manual_seed(2)
            self.A = torch.nn.Parameter(-torch.eye(n, dtype=torch.float64) + 0.1 * torch.randn(n, n, generator=gg, dtype=torch.float64))

        def forward(self, t, y):
            return y @ self.A.T

    n, B = 5, 6
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(B, n, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.2], dtype=torch.float64)
    gout = torch.randn(2, B, n, generator=g, dtype=torch.float64)
    f_im, f_ex = Lin(n), TimeMLP(d=n, hidden=8)
    argv = ["-ts_adapt_type", "none", "-snes_type", "ksponly"]
    kw = dict(method="imex", imex_form=True, batch_size=B, linear_solver="torch")
    a1, a2 = _run_product(monkeypatch, argv, dict(kw, fixed_jacobian=True), [f_im, f_ex], u0, t, gout, 0.1, twice=True)
    ode = a2[3]
    inv = dict(ode._imp._inv)
    assert len(inv) >= 1 and all(torch.equal(x, y) for x, y in zip(a1[2], a2[2])) and torch.equal(a1[0], a2[0])
    ode.setupTS(u0, a2[4][0], step_size=0.1, enable_adjoint=True, func2=a2[4][1], **dict(kw, fixed_jacobian=True))
    ode.odeint_adjoint(u0.clone().requires_grad_(True), t)
    assert all(ode._imp._inv[k] is v for k, v in inv.items()), "rebuilt although nothing changed"
    with torch.no_grad():
        a2[4][0].A.mul_(1.0)
    ode.odeint_adjoint(u0.clone().requires_grad_(True), t)
    assert all(ode._imp._inv[k] is not v for k, v in inv.items()), "kept although a parameter was touched"
    b1, b2 = _run_product(monkeypatch, argv, kw, [f_im, f_ex], u0, t, gout, 0.1, twice=True)
    assert torch.equal(b1[0], a1[0]) and all(torch.equal(x, y) for x, y in zip(b1[2], a1[2]))


class _Decay(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.k = torch.nn.Parameter(torch.tensor([1.0], dtype=torch.float64))

    def forward(self, t, u):
        return -self.k * u


@pytest.mark.parametrize("implicit_form", [False, True])
@pytest.mark.parametrize("how", ["method", "option"])
def test_arkimex_with_a_single_function_integrates_it_once(monkeypatch, implicit_form, how):
    """Oracle-independent: u' = -k u, u(0) = 1 must give exp(-t) (an ARKIMEX scheme handed ONE function -- method='imex'
    without imex_form, or -ts_type arkimex on an ordinary model -- registers it as the RHSFunction, or as the IFunction when
    implicit_form=True, reference petsc_adjoint.py:666-730; treating it as both halves integrates u' = 2 f -> exp(-2t))."""
    import math

    argv = ["-ts_adapt_type", "none"] + (["-ts_type", "arkimex"] if how == "option" else [])
    u0 = torch.ones(3, dtype=torch.float64)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    kw = dict(method="imex" if how == "method" else "rk4", implicit_form=implicit_form)
    gout = torch.stack([torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)])
    o, p = _both(monkeypatch, argv, kw, [_Decay()], u0, t, gout, 0.05)
    for res in (o, p):
        assert abs(res[0][-1][0].item() - math.exp(-1.0)) < 5e-6
        assert abs(res[1][0].item() - math.exp(-1.0)) < 5e-6          # d u(1) / d u0
        assert abs(res[2][0].item() + 3 * math.exp(-1.0)) < 5e-5      # d sum u(1) / d k = -t exp(-k t) per component
    _assert_close(p, o, 1e-12)


@pytest.mark.parametrize("n,c", [(12, 3), (48, 4), (100, 2), (7, 7), (300, 5), (5, 1)])
def test_revolve_schedule_is_binomial(n, c):
    """The checkpoint placement of -ts_trajectory_max_cps_ram: simulate the forward placement + the reverse sweep's
    re-checkpointing for n steps and c slots; never more than c states held, every step's stages produced exactly once in reverse
    order, and no step advanced more often than the optimal repetition number r (beta(c, r) >= n > beta(c, r - 1)) + 2."""
    import math

    from pnode_b200.engine import revolve_forward_positions, revolve_split

    stored = set(revolve_forward_positions(n, c))
    assert 0 in stored and len(stored) <= c
    advanced = [1] * n  # the forward sweep itself
    peak = len(stored)
    for idx in range(n - 1, -1, -1):
        pos = max(j for j in stored if j <= idx)
        while pos < idx:
            free = c - len(stored)
            m = revolve_split(idx + 1 - pos, free) if free >= 1 else idx - pos
            for i in range(pos, pos + m):
                advanced[i] += 1
            pos += m
            if free >= 1 and pos < idx:
                stored.add(pos)
                peak = max(peak, len(stored))
        advanced[idx] += 1  # its stages
        stored.discard(idx)
    assert peak <= c
    r = 1
    while math.comb(c + r, c) < n:
        r += 1
    if c > 1:
        # forward sweep + stage evaluation + at most r re-advances per step; total of the order of r sweeps, far below the
        # quadratic cost of a single checkpoint
        assert max(advanced) <= r + 2, (max(advanced), r)
        assert sum(advanced) <= (r + 2) * n


def test_checkpoint_budget_is_respected_and_results_identical(monkeypatch):
    pa = patch_cpu(monkeypatch)
    func = TimeMLP(d=4, hidden=8)
    g = torch.Generator().manual_seed(6)
    u0 = torch.randn(6, 4, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 1.0, 2.0, 3.0], dtype=torch.float64)
    gout = torch.randn(4, 6, 4, generator=g, dtype=torch.float64)

    def run(argv, method="rk4"):
        Options.clear_all()
        Options.insert_args(argv)
        f = copy.deepcopy(func)
        ode = pa.ODEPetsc()
        ode.setupTS(u0, f, step_size=0.0625, method=method, enable_adjoint=True)
        y0 = u0.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t)
        (out * gout).sum().backward()
        return out.detach(), y0.grad, [p.grad for p in f.parameters()], ode

    full = run(["-ts_adapt_type", "none"])
    for cps in (2, 4, 9):
        lean = run(["-ts_adapt_type", "none", "-ts_trajectory_max_cps_ram", str(cps)])
        eng = lean[3]._engine
        assert torch.equal(full[0], lean[0]) and torch.equal(full[1], lean[1])
        assert all(torch.equal(a, b) for a, b in zip(full[2], lean[2]))
        assert eng.peak_checkpoints <= cps
        # 48 steps: the binomial schedule needs about r sweeps (beta(4, 4) = 70 >= 48: r = 4), not 48^2 / 2 steps
        assert 48 <= eng.recomputed_steps <= {2: 48 * 10, 4: 48 * 6, 9: 48 * 4}[cps]
    # adaptive run (step count unknown in advance): budget still respected, results identical
    fulla = run(["-ts_rtol", "1e-6", "-ts_atol", "1e-6"], method="dopri5")
    leana = run(["-ts_rtol", "1e-6", "-ts_atol", "1e-6", "-ts_trajectory_max_cps_ram", "4"], method="dopri5")
    assert torch.equal(fulla[0], leana[0]) and torch.equal(fulla[1], leana[1])
    assert leana[3]._engine.peak_checkpoints <= 4


def test_timeloop_follow_mirrors_report():
    """TimeLoop.follow (bookkeeping behind the device controller) reaches the same state as report() when it is fed the
    verdicts and step sizes report() produced."""
    import random

    from pnode_b200.controller import TimeLoop

    rng = random.Random(4)
    for times in ([0.0, 0.3, 0.35, 1.0], [0.7], [0.0, 1.0]):
        a = TimeLoop(times, 0.2, True, 5, True)
        b = TimeLoop(times, 0.2, True, 5, True)
        slots_a, slots_b = [], []
        while not a.done:
            enorm = rng.choice([0.01, 0.3, 0.9, 1.7, 3.0])
            ok = a.report(enorm)
            assert b.follow(ok, enorm, a.t, a.h) == ok
            if ok:
                slots_a.append(a.last_out_slot), slots_b.append(b.last_out_slot)
        assert b.done and slots_a == slots_b
        for k in ("t", "h", "steps", "ctr", "cur_sol_index", "cur_sol_steps", "attempts", "last_h"):
            assert getattr(a, k) == getattr(b, k), k
        b.check_complete()


def test_block_gmres_solves_every_column_to_its_own_tolerance():
    """engine.block_gmres (linear_solver="hpddm"): one right-hand side per segment, per-segment Krylov spaces advanced in
    lockstep; against a dense solve.  Restart shorter than the iteration count, one zero right-hand side, one segment whose
    operator is the identity (converges in one iteration and drops out while the others go on)."""
    from _fake_ops import FakeOps
    from pnode_b200.engine import block_gmres

    nseg, n = 7, 12
    g = torch.Generator().manual_seed(11)
    A = torch.eye(n, dtype=torch.float64).repeat(nseg, 1, 1) + 0.4 * torch.randn(nseg, n, n, generator=g,
                                                                                 dtype=torch.float64) / n ** 0.5
    A[3] = torch.eye(n, dtype=torch.float64)
    B = torch.randn(nseg, n, generator=g, dtype=torch.float64)
    B[5] = 0.0
    B[2] *= 1e6  # per-column relative tolerance: a large column must not hide the small ones
    ops = FakeOps()
    calls = []

    def op(v):
        calls.append(1)
        return torch.einsum("sij,sj->si", A, v.view(nseg, n)).reshape(-1)

    x, its, res = block_gmres(ops, op, B.reshape(-1).clone(), nseg, rtol=1e-12, restart=5, max_it=200)
    want = torch.linalg.solve(A, B.unsqueeze(-1)).squeeze(-1)
    for s in range(nseg):
        err = (x.view(nseg, n)[s] - want[s]).norm() / max(float(want[s].norm()), 1e-300)
        assert err < 1e-10, (s, float(err))
    assert torch.equal(x.view(nseg, n)[5], torch.zeros(n, dtype=torch.float64))
    assert 12 <= its <= 40 and len(calls) >= its  # n = 12 unknowns per column; restarts add true-residual applications


@pytest.mark.parametrize("method", ["cn", "beuler"])
def test_hpddm_block_solver_matches_dense_oracle(monkeypatch, method):
    """linear_solver="hpddm" (pnode/hpddm_linearsolve.py): Newton with the block Krylov solver, batch_size right-hand sides."""
    func = SpiralFunc(bias_std=0.1)
    u0, _, gout = spiral_inputs(40)
    t = torch.tensor([0.0, 0.1, 0.2], dtype=torch.float64)
    argv = ["-ts_adapt_type", "none", "-ksp_rtol", "1e-12"]
    o, p = _both(monkeypatch, argv, dict(method=method, implicit_form=True, linear_solver="hpddm", batch_size=40), [func],
                 u0, t, gout[:3], 0.1)
    assert p[3]._imp.krylov_iterations > 0
    _assert_close(p, o, 1e-8)


@pytest.mark.parametrize("dim,hidden,hpad", [(2, 16, 50), (3, 50, 50), (4, 33, 50), (2, 64, 100)])
def test_fused_mlp_padding_round_trip(dim, hidden, hpad):
    """Host side of the widened tiny-MLP shapes (fused.FusedMlpRK): a layer narrower than the compiled width runs zero-padded;
    mu comes back in the padded layout and is cut to the module's parameter order and sizes."""
    from pnode_b200.fused import FusedMlpRK, MlpSpec

    lin1, lin2 = torch.nn.Linear(dim, hidden).double(), torch.nn.Linear(hidden, dim).double()
    f = FusedMlpRK.__new__(FusedMlpRK)
    f.spec, f.hidden = MlpSpec(lin1, lin2, 1, dim, hidden), hpad
    f._pad = None if hpad == hidden else (torch.zeros(hpad, dim, dtype=torch.float64), torch.zeros(hpad, dtype=torch.float64),
                                          torch.zeros(dim, hpad, dtype=torch.float64))
    g = torch.Generator().manual_seed(0)
    want = [torch.randn(p.shape, generator=g, dtype=torch.float64) for p in (lin1.weight, lin1.bias, lin2.weight, lin2.bias)]
    w1p, b1p, w2p = torch.full((hpad, dim), 7.0).double(), torch.full((hpad,), 7.0).double(), torch.full((dim, hpad), 7.0).double()
    w1p[:hidden], b1p[:hidden], w2p[:, :hidden] = want[0], want[1], want[2]
    mu = torch.cat([w1p.reshape(-1), b1p, w2p.reshape(-1), want[3]])
    got = f._unpad_mu(mu)
    assert got.numel() == sum(w.numel() for w in want)
    assert torch.equal(got, torch.cat([w.reshape(-1) for w in want]))
    if f._pad is not None:
        f.code = 1
        d = f._desc()  # copies the live parameters into the padded buffers
        assert d.hidden == hpad and torch.equal(f._pad[0][:hidden], lin1.weight.detach())
        assert float(f._pad[0][hidden:].abs().sum() + f._pad[1][hidden:].abs().sum() + f._pad[2][:, hidden:].abs().sum()) == 0.0
        assert torch.equal(f._pad[2][:, :hidden], lin2.weight.detach())
