"""GPU parity of the fused spiral sweeps (csrc/mlp_rk.cu, reached through ODEPetsc -> ctypes -> C ABI) against the oracle
on the same seeded inputs.  Bar: 1e-10 relative in fp64, 1e-4 in fp32 (BASELINE.json north_star)."""
import copy

import pytest
import torch

from oracle import OracleODEPetsc
from pnode_b200.options import Options
from _problems import SpiralFunc, rel_err, spiral_inputs

pytestmark = pytest.mark.gpu
TOL = {torch.float64: 1e-10, torch.float32: 1e-4}


def _oracle(func, u0, t, gout, method, step, argv):
    f = copy.deepcopy(func).double()
    ode = OracleODEPetsc(argv)
    ode.setupTS(u0.double(), f, step_size=step, method=method, enable_adjoint=True)
    y0 = u0.double().clone().requires_grad_(True)
    out = ode.odeint_adjoint(y0, t)
    (out * gout.double()).sum().backward()
    return out.detach(), y0.grad, [p.grad for p in f.parameters()], ode


def _product(func, u0, t, gout, method, step, argv, fused=True):
    from pnode import petsc_adjoint  # the drop-in import path

    Options.clear_all()
    Options.insert_args(argv + ([] if fused else ["-pnode_fused", "0"]))
    f = copy.deepcopy(func).cuda()
    ode = petsc_adjoint.ODEPetsc()
    ode.setupTS(u0.cuda(), f, step_size=step, method=method, enable_adjoint=True)
    y0 = u0.cuda().clone().requires_grad_(True)
    out = ode.odeint_adjoint(y0, t.cuda())
    (out * gout.cuda()).sum().backward()
    torch.cuda.synchronize()
    return out.detach(), y0.grad, [p.grad for p in f.parameters()], ode


def _compare(p, o, tol):
    assert p[0].shape == o[0].shape
    assert rel_err(p[0], o[0]) < tol, "trajectory"
    assert rel_err(p[1], o[1]) < tol, "lambda"
    for a, b in zip(p[2], o[2]):
        assert rel_err(a, b) < tol, "mu"


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("batch", [20, 1, 127, 129, 1000])
def test_config1_rk4_fused_matches_oracle(dtype, batch):
    """BASELINE config 1: 2-D state, tanh MLP 2-50-2, RK4, batch_time 10 (9 steps of 0.025), incl. ragged tiles."""
    func = SpiralFunc(dtype=dtype)
    u0, t, gout = spiral_inputs(batch, dtype=dtype)
    argv = ["-ts_adapt_type", "none"]
    o = _oracle(func, u0, t, gout, "rk4", 0.025, argv)
    p = _product(func, u0, t, gout, "rk4", 0.025, argv)
    assert p[3].path == "fused-mlp-rk"
    _compare(p, o, TOL[dtype])
    assert p[3]._loop.cur_sol_steps == o[3].cur_sol_steps


@pytest.mark.parametrize("method", ["euler", "rk2", "bosh3", "dopri5", "midpoint"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_every_explicit_scheme_fused(method, dtype):
    func = SpiralFunc(dtype=dtype, bias_std=0.2, seed=3)
    u0, t, gout = spiral_inputs(300, T=5, dtype=dtype, h=0.05)
    argv = ["-ts_adapt_type", "none"]
    o = _oracle(func, u0, t, gout, method, 0.025, argv)  # two steps per output interval
    p = _product(func, u0, t, gout, method, 0.025, argv)
    assert p[3].path == "fused-mlp-rk"
    _compare(p, o, TOL[dtype])


def test_identity_phi_and_rk_type_option():
    func = SpiralFunc(cube=False, bias_std=0.3, seed=5)
    u0, t, gout = spiral_inputs(64, T=4)
    argv = ["-ts_adapt_type", "none", "-ts_type", "rk", "-ts_rk_type", "3"]
    o = _oracle(func, u0, t, gout, "dopri5", 0.0125, argv)
    p = _product(func, u0, t, gout, "dopri5", 0.0125, argv)
    assert p[3].path == "fused-mlp-rk" and p[3]._scheme.name == "3"
    _compare(p, o, 1e-10)


def test_single_time_point_and_uneven_steps():
    func = SpiralFunc(bias_std=0.1)
    u0, _, gout = spiral_inputs(33)
    argv = ["-ts_adapt_type", "none"]
    t = torch.tensor([0.11], dtype=torch.float64)  # 0.025 x4 then a clamped last step
    o = _oracle(func, u0, t, gout[:1], "rk4", 0.025, argv)
    p = _product(func, u0, t, gout[:1], "rk4", 0.025, argv)
    assert p[0].shape == (1, 33, 1, 2)
    assert rel_err(p[0], o[0]) < 1e-10 and rel_err(p[1], o[1]) < 1e-10
    t = torch.tensor([0.0, 0.03, 0.1, 0.1 + 1e-3, 0.2], dtype=torch.float64)
    o = _oracle(func, u0, t, gout[:5], "bosh3", 0.04, argv)
    p = _product(func, u0, t, gout[:5], "bosh3", 0.04, argv)
    _compare(p, o, 1e-10)
    assert [a[:2] for a in p[3]._loop.attempts] == pytest.approx([a[:2] for a in o[3].ts.log])


def test_fused_equals_generic_path_and_is_bit_reproducible():
    func = SpiralFunc()
    u0, t, gout = spiral_inputs(5000)
    argv = ["-ts_adapt_type", "none"]
    a = _product(func, u0, t, gout, "rk4", 0.025, argv, fused=True)
    b = _product(func, u0, t, gout, "rk4", 0.025, argv, fused=False)
    assert a[3].path == "fused-mlp-rk" and b[3].path == "generic"
    _compare(a, b, 1e-11)
    a2 = _product(func, u0, t, gout, "rk4", 0.025, argv, fused=True)
    assert torch.equal(a[0], a2[0]) and torch.equal(a[1], a2[1])
    assert all(torch.equal(x, y) for x, y in zip(a[2], a2[2]))  # fixed-order mu reduction


def test_parameters_are_borrowed_not_copied():
    """An optimiser step between calls must be visible without a new setupTS (SURVEY.md 8b 'Ownership')."""
    from pnode import petsc_adjoint

    Options.insert_args(["-ts_adapt_type", "none"])
    func = SpiralFunc().cuda()
    u0, t, gout = spiral_inputs(16)
    ode = petsc_adjoint.ODEPetsc()
    ode.setupTS(u0.cuda(), func, step_size=0.025, method="rk4")
    y1 = ode.odeint_adjoint(u0.cuda(), t.cuda()).detach().clone()
    with torch.no_grad():
        for p in func.parameters():
            p.mul_(1.5)
    y2 = ode.odeint_adjoint(u0.cuda(), t.cuda()).detach()
    ref = OracleODEPetsc(["-ts_adapt_type", "none"])
    fcpu = copy.deepcopy(func).cpu()
    ref.setupTS(u0, fcpu, step_size=0.025, method="rk4")
    assert rel_err(y2, ref.odeint(u0, t)) < 1e-10 and rel_err(y1, y2) > 1e-4


@pytest.mark.parametrize("lean", [False, True])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_full_size_2pow20_properties(dtype, lean):
    """BASELINE config 2 at full size through size-independent properties: the batch is the 20-trajectory base case
    tiled 2^20/20 times (+ remainder), so every replica must reproduce the oracle's base trajectories / lambda and mu must
    be the replica-count-weighted sum of the oracle's per-trajectory contributions.  lean: with solution checkpoints only
    (-ts_trajectory_solution_only 1, stages recomputed inside the adjoint kernel)."""
    B = 1 << 20
    func = SpiralFunc(dtype=dtype)
    u0b, t, goutb = spiral_inputs(32, dtype=dtype)
    reps = B // 32
    u0 = u0b.repeat(reps, 1, 1)
    gout = goutb.repeat(1, reps, 1, 1)
    argv = ["-ts_adapt_type", "none"]
    o = _oracle(func, u0b, t, goutb, "rk4", 0.025, argv)
    p = _product(func, u0, t, gout, "rk4", 0.025, argv + (["-ts_trajectory_solution_only", "1"] if lean else []))
    assert p[3].path == "fused-mlp-rk" and p[3]._fused.solution_only == lean
    tol = TOL[dtype]
    out = p[0].view(10, reps, 32, 1, 2)
    assert rel_err(out[:, 0], o[0]) < tol and rel_err(out[:, reps - 1], o[0]) < tol and rel_err(out[:, reps // 2], o[0]) < tol
    assert float((out - out[:, :1]).abs().max()) == 0.0  # identical inputs => bit-identical replicas
    lam = p[1].view(reps, 32, 1, 2)
    assert rel_err(lam[0], o[1]) < tol and float((lam - lam[:1]).abs().max()) == 0.0
    for a, b in zip(p[2], o[2]):
        assert rel_err(a, b * reps) < tol


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("method,batch", [("rk4", 5000), ("rk4", 37), ("fixed_dopri5", 4500), ("bosh3", 4200),
                                          ("rk2", 4300), ("euler", 4100)])
def test_solution_only_checkpoints_inside_the_fused_sweep(dtype, method, batch):
    """-ts_trajectory_solution_only 1 on the fused path: the forward sweep keeps u_n per step (1/s of the
    stage checkpoints), the adjoint kernel recomputes the stages with the forward sweep's arithmetic -- bit-identical to
    the stage-checkpoint sweeps, and to the oracle within the bar."""
    func = SpiralFunc(dtype=dtype)
    u0, t, gout = spiral_inputs(batch, dtype=dtype)
    argv = ["-ts_adapt_type", "none"]
    full = _product(func, u0, t, gout, method, 0.025, argv)
    lean = _product(func, u0, t, gout, method, 0.025, argv + ["-ts_trajectory_solution_only", "1"])
    assert full[3].path == "fused-mlp-rk" and lean[3].path == "fused-mlp-rk"
    assert lean[3]._fused.solution_only and not full[3]._fused.solution_only
    if batch > 4096:  # below, the stage-checkpoint run uses the small-batch kernels (different summation order)
        assert torch.equal(full[0], lean[0])
        assert torch.equal(full[1], lean[1])
        assert all(torch.equal(a, b) for a, b in zip(full[2], lean[2]))
    o = _oracle(func, u0, t, gout, method, 0.025, argv)
    _compare(lean, o, TOL[dtype])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("dim,hidden,cube,method,batch", [
    (2, 100, True, "rk4", 700), (2, 64, False, "fixed_dopri5", 300), (2, 16, True, "rk4", 5000), (2, 51, True, "bosh3", 130),
    (3, 50, True, "rk4", 600), (3, 20, False, "rk2", 4200), (4, 50, True, "rk4", 515), (4, 33, True, "fixed_dopri5", 100),
    (1, 50, False, "rk4", 333), (1, 8, True, "euler", 64)])
def test_other_state_sizes_and_widths_take_the_fused_sweeps(dtype, dim, hidden, cube, method, batch):
    """The recogniser beyond the spiral shape: 1- to 4-dimensional states, any hidden width up to the compiled ones
    (narrower layers are zero-padded on the host: csrc/mlp_rk.cu PNODE_FOR_SHAPES), with and without stage checkpoints."""
    func = SpiralFunc(dtype=dtype, hidden=hidden, cube=cube, bias_std=0.2, dim=dim, seed=3)
    u0, t, gout = spiral_inputs(batch, dtype=dtype, dim=dim)
    argv = ["-ts_adapt_type", "none"]
    o = _oracle(func, u0, t, gout, method, 0.025, argv)
    p = _product(func, u0, t, gout, method, 0.025, argv)
    assert p[3].path == "fused-mlp-rk"
    assert [g.shape for g in p[2]] == [g.shape for g in o[2]]
    _compare(p, o, TOL[dtype])
    lean = _product(func, u0, t, gout, method, 0.025, argv + ["-ts_trajectory_solution_only", "1"])
    assert lean[3].path == "fused-mlp-rk"
    assert torch.equal(lean[0], p[0]) and torch.equal(lean[1], p[1]) and all(torch.equal(a, b) for a, b in zip(lean[2], p[2]))


def test_layers_wider_than_the_compiled_shapes_take_the_generic_path():
    func = SpiralFunc(hidden=128)
    u0, t, gout = spiral_inputs(40)
    argv = ["-ts_adapt_type", "none"]
    p = _product(func, u0, t, gout, "rk4", 0.025, argv)
    assert p[3].path == "generic"
    _compare(p, _oracle(func, u0, t, gout, "rk4", 0.025, argv), 1e-10)
