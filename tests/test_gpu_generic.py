"""GPU parity of the generic path (torch evaluates f / VJPs on the GPU, csrc/vecops.cu does the TS arithmetic) against
the oracle: the reference's own ROBER known answers, adaptive dopri5 step sequences, IMEX with the batched solver."""
import copy

import pytest
import torch

from oracle import OracleODEPetsc
from pnode_b200.options import Options
from _problems import (PETSC_ARGS, ROBER_STEPS, ROBER_T, Rober, RoberEX, RoberIM, TimeMLP, rel_err, rober_truth)

pytestmark = pytest.mark.gpu


def _run(make_ode, funcs, kw, u0, t, gout, step, dev):
    fs = [copy.deepcopy(f).to(dev) for f in funcs]
    kw = dict(kw)
    if len(fs) == 2:
        kw["func2"] = fs[1]
    ode = make_ode()
    ode.setupTS(u0.to(dev), fs[0], step_size=step, enable_adjoint=True, **kw)
    y0 = u0.to(dev).clone().requires_grad_(True)
    out = ode.odeint_adjoint(y0, t.to(dev))
    (out * gout.to(dev)).sum().backward()
    return out.detach(), y0.grad, [p.grad for f in fs for p in f.parameters()], ode


def _pair(argv, funcs, kw, u0, t, gout, step):
    from pnode import petsc_adjoint

    Options.clear_all()
    Options.insert_args(argv)
    o = _run(lambda: OracleODEPetsc(argv), funcs, kw, u0, t, gout, step, "cpu")
    p = _run(lambda: petsc_adjoint.ODEPetsc(), funcs, kw, u0, t, gout, step, "cuda")
    return o, p


def _compare(p, o, tol):
    assert rel_err(p[0], o[0]) < tol
    assert rel_err(p[1], o[1]) < tol
    assert len(p[2]) == len(o[2])
    for a, b in zip(p[2], o[2]):
        assert rel_err(a, b) < tol


def test_reference_rober_known_answers_on_gpu():
    """The reference's three tests (tests/test_pnode.py:133-201), run through the drop-in on the GPU."""
    true_y = rober_truth()
    gout = torch.ones_like(true_y)
    cases = [(dict(method="cn", implicit_form=True), [Rober()], 1.85e-6, 3.36e-6, 1e-6),
             (dict(method="imex", implicit_form=True, imex_form=True), [RoberIM(), RoberEX()], 3.11e-6, 5.65e-6, 3e-6),
             (dict(method="rk3"), [Rober()], 1.85e-6, 3.21e-6, 1e-6)]
    for kw, funcs, g_loss, g_std, tol in cases:
        o, p = _pair(PETSC_ARGS, funcs, kw, true_y[0], ROBER_T, gout, ROBER_STEPS)
        err = torch.abs(p[0].cpu() - true_y)
        assert err.mean().item() == pytest.approx(g_loss, abs=tol)
        assert err.std().item() == pytest.approx(g_std, abs=tol)
        _compare(p, o, 1e-8)


@pytest.mark.parametrize("dtype,tol,ts_tol", [(torch.float64, 1e-10, "1e-6"), (torch.float32, 1e-4, "1e-5")])
def test_adaptive_dopri5_identical_accepted_steps(dtype, tol, ts_tol):
    func = TimeMLP(d=6, hidden=16, dtype=dtype)
    g = torch.Generator().manual_seed(5)
    u0 = torch.randn(1000, 6, generator=g, dtype=torch.float64).to(dtype)
    t = torch.tensor([0.0, 0.4, 1.0], dtype=torch.float64)
    gout = torch.randn(3, 1000, 6, generator=g, dtype=torch.float64).to(dtype)
    o, p = _pair(["-ts_rtol", ts_tol, "-ts_atol", ts_tol], [func], dict(method="dopri5"), u0, t, gout, 0.3)
    lo, lp = o[3].ts.log, p[3]._loop.attempts
    assert [a[2] for a in lo] == [a[2] for a in lp]  # same accept / reject pattern
    if dtype == torch.float64:
        assert any(not a[2] for a in lo), "the fp64 case must exercise a rejected attempt"
    htol = 1e-9 if dtype == torch.float64 else 1e-2
    for a, b in zip(lo, lp):
        assert a[1] == pytest.approx(b[1], rel=htol)
    _compare(p, o, tol)


@pytest.mark.parametrize("name", ["3", "l2", "4"])
def test_imex_torch_solver_on_gpu(name):
    from test_oracle_adjoint import LinearIM

    N, B = 16, 32
    g = torch.Generator().manual_seed(7)
    u0 = torch.randn(B, N, generator=g, dtype=torch.float64) * 0.5
    t = torch.tensor([0.0, 0.2, 0.4], dtype=torch.float64)
    gout = torch.randn(3, B, N, generator=g, dtype=torch.float64)
    argv = ["-ts_adapt_type", "none", "-snes_type", "ksponly", "-ts_arkimex_type", name]
    o, p = _pair(argv, [LinearIM(N), TimeMLP(d=N, hidden=24)],
                 dict(method="imex", imex_form=True, batch_size=B, linear_solver="torch"), u0, t, gout, 0.1)
    _compare(p, o, 1e-10)


@pytest.mark.parametrize("extra", [["-ts_trajectory_solution_only", "1"], ["-ts_trajectory_max_cps_ram", "2"]])
def test_bounded_checkpoint_storage_on_gpu(extra):
    """-ts_trajectory_solution_only / -ts_trajectory_max_cps_ram: less HBM, recomputed stages, identical
    results (same kernels, same arithmetic)."""
    from pnode import petsc_adjoint

    func = TimeMLP(d=6, hidden=16)
    g = torch.Generator().manual_seed(8)
    u0 = torch.randn(300, 6, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64)
    gout = torch.randn(3, 300, 6, generator=g, dtype=torch.float64)

    def run(argv):
        Options.clear_all()
        Options.insert_args(["-ts_adapt_type", "none"] + argv)
        return _run(lambda: petsc_adjoint.ODEPetsc(), [func], dict(method="dopri5"), u0, t, gout, 0.05, "cuda")

    full, lean = run([]), run(extra)
    assert torch.equal(full[0], lean[0]) and torch.equal(full[1], lean[1])
    assert all(torch.equal(a, b) for a, b in zip(full[2], lean[2]))
    assert lean[3]._engine.recomputed_steps >= 20 and full[3]._engine.recomputed_steps == 0


@pytest.mark.parametrize("method", ["cn", "beuler"])
def test_matrix_free_newton_gmres_on_gpu(method):
    """Implicit theta methods with the reference's default linear_solver="petsc": matrix-free Newton-GMRES on the GPU
    (pnode_mdot / pnode_lincomb Krylov kernels, torch forward/reverse-mode J products) vs the oracle's dense Newton."""
    from _problems import SpiralFunc, spiral_inputs

    func = SpiralFunc(bias_std=0.1)
    u0, _, gout = spiral_inputs(300)
    t = torch.tensor([0.0, 0.1, 0.2], dtype=torch.float64)
    o, p = _pair(["-ts_adapt_type", "none", "-ksp_rtol", "1e-10", "-pnode_fused", "0"], [func],
                 dict(method=method, implicit_form=True), u0, t, gout[:3], 0.1)
    assert p[3]._imp.krylov_iterations > 0
    _compare(p, o, 1e-7)


@pytest.mark.parametrize("method", ["cn", "beuler"])
def test_hpddm_block_krylov_on_gpu(method):
    """linear_solver="hpddm" (pnode/hpddm_linearsolve.py:13-49, KSPHPDDM BGMRES on the batch as a block of right-hand sides):
    engine.block_gmres on the pnode_mdot_seg / pnode_lincomb_seg kernels, one Krylov space per sample, vs the oracle's dense
    Newton."""
    from _problems import SpiralFunc, spiral_inputs

    func = SpiralFunc(bias_std=0.1)
    u0, _, gout = spiral_inputs(300)
    t = torch.tensor([0.0, 0.1, 0.2], dtype=torch.float64)
    o, p = _pair(["-ts_adapt_type", "none", "-ksp_rtol", "1e-12", "-pnode_fused", "0"], [func],
                 dict(method=method, implicit_form=True, linear_solver="hpddm", batch_size=300), u0, t, gout[:3], 0.1)
    assert p[3]._imp.krylov_iterations > 0
    _compare(p, o, 1e-9)


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
def test_mass_matrix_dae_on_gpu(method):
    """mass= (M u' = f, singular M: the reference's evalIFunction, petsc_adjoint.py:426-431, as examples-pnode/
    pendulum_DAE.py:119-139 uses it) through the drop-in on the GPU: trajectory, lambda and mu against the oracle, and mu
    against a finite difference of the discrete map."""
    from pnode import petsc_adjoint

    M = torch.eye(5, dtype=torch.float64)
    M[-1, -1] = 0.0
    u0 = torch.tensor([1.0, 0.0, 0.0, 1.0, 1.0], dtype=torch.float64)  # consistent initial value: lambda = |v|^2 - g y
    t = torch.tensor([0.0, 0.02, 0.05], dtype=torch.float64)
    gout = torch.tensor([[0.3, -0.2, 0.5, 0.1, 0.0], [1.0, 0.5, -0.3, 0.2, 0.1], [0.7, -1.0, 0.4, 0.3, -0.2]],
                        dtype=torch.float64)
    kw = dict(method=method, implicit_form=True, mass=M)
    o, p = _pair(["-ts_adapt_type", "none"], [_Pendulum()], kw, u0, t, gout, 0.01)
    _compare(p, o, 1e-9)  # Newton iterations converge to the SNES tolerance on both sides: parity to solver tolerance
    base = (p[0].cpu() * gout).sum().item()
    f2 = _Pendulum().cuda()
    with torch.no_grad():
        f2.g.add_(1e-6)
    ode = petsc_adjoint.ODEPetsc()
    ode.setupTS(u0.cuda(), f2, step_size=0.01, enable_adjoint=False, **kw)
    pert = (ode.odeint(u0.cuda(), t.cuda()).cpu() * gout).sum().item()
    assert (pert - base) / 1e-6 == pytest.approx(p[2][0].item(), rel=1e-4)
