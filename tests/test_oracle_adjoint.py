"""Gradients are 'parity unpinned' in the reference (it prints them, asserts nothing: tests/test_pnode.py:149-150), so
the oracle's discrete adjoints are checked against torch.autograd through the unrolled scheme."""
import pytest
import torch
import torch.nn as nn

from oracle import OracleODEPetsc
from oracle import tableaux as otab
from _problems import SpiralFunc, TimeMLP, rel_err, spiral_inputs
from _unrolled import ark_unrolled_linear_im, rk_unrolled, theta_unrolled

ARGS = ["-ts_adapt_type", "none"]


def _oracle_grads(ode_kw, funcs, u0, t, gout, argv=ARGS, step=0.025):
    for f in funcs:
        f.zero_grad()
    ode = OracleODEPetsc(argv)
    ode.setupTS(u0, funcs[0], step_size=step, enable_adjoint=True, **ode_kw)
    y0 = u0.clone().requires_grad_(True)
    out = ode.odeint_adjoint(y0, t)
    (out * gout).sum().backward()
    return out.detach(), y0.grad.clone(), [p.grad.clone() for f in funcs for p in f.parameters() if p.requires_grad], ode


@pytest.mark.parametrize("method,scheme", [("rk4", "4"), ("bosh3", "3bs"), ("dopri5", "5dp"), ("euler", "1fe"), ("rk2", "2b")])
def test_rk_adjoint_matches_autograd(method, scheme):
    func = SpiralFunc(bias_std=0.1)
    u0, t, gout = spiral_inputs(20)
    out, lam, mu, ode = _oracle_grads(dict(method=method), [func], u0, t, gout)
    A, b, _, c = otab.RK[scheme].floats()
    sched = [(tt, hh, k + 1) for k, (tt, hh, ok, _) in enumerate(ode.ts.log)]
    assert len(sched) == 9
    func.zero_grad()
    y0 = u0.clone().requires_grad_(True)
    outs, _ = rk_unrolled(lambda tt, y: func(tt, y), y0, sched, A, b, c)
    ref = torch.stack([outs[k] for k in range(len(t))])
    (ref * gout).sum().backward()
    assert rel_err(out, ref) < 1e-14
    assert rel_err(lam, y0.grad) < 1e-12
    for g, p in zip(mu, func.parameters()):
        assert rel_err(g, p.grad) < 1e-12


@pytest.mark.parametrize("scheme", ["5bs", "5f", "3", "2a"])
def test_rk_adjoint_of_option_selected_schemes_matches_autograd(scheme):
    """Schemes only reachable through -ts_rk_type (petsc_adjoint.py:775), incl. the 8-stage FSAL Bogacki-Shampine 5(4)."""
    func = SpiralFunc(bias_std=0.1)
    u0, t, gout = spiral_inputs(20)
    out, lam, mu, ode = _oracle_grads(dict(method="rk4"), [func], u0, t, gout, argv=ARGS + ["-ts_rk_type", scheme])
    A, b, _, c = otab.RK[scheme].floats()
    sched = [(tt, hh, k + 1) for k, (tt, hh, ok, _) in enumerate(ode.ts.log)]
    func.zero_grad()
    y0 = u0.clone().requires_grad_(True)
    outs, _ = rk_unrolled(lambda tt, y: func(tt, y), y0, sched, A, b, c)
    ref = torch.stack([outs[k] for k in range(len(t))])
    (ref * gout).sum().backward()
    assert rel_err(out, ref) < 1e-14
    assert rel_err(lam, y0.grad) < 1e-12
    for g, p in zip(mu, func.parameters()):
        assert rel_err(g, p.grad) < 1e-12


def test_rk_adjoint_single_time_point():
    func = SpiralFunc(bias_std=0.1)
    u0, _, gout = spiral_inputs(7)
    t = torch.tensor([0.1], dtype=torch.float64)
    out, lam, mu, ode = _oracle_grads(dict(method="rk4"), [func], u0, t, gout[:1])
    assert out.shape == (1, 7, 1, 2) and len(ode.ts.log) == 4  # integrates [0, 0.1] in 4 steps of 0.025
    A, b, _, c = otab.RK["4"].floats()
    sched = [(tt, hh, -1) for (tt, hh, ok, _) in ode.ts.log]
    func.zero_grad()
    y0 = u0.clone().requires_grad_(True)
    _, uf = rk_unrolled(lambda tt, y: func(tt, y), y0, sched, A, b, c)
    (uf * gout[0]).sum().backward()
    assert rel_err(lam, y0.grad) < 1e-12
    for g, p in zip(mu, func.parameters()):
        assert rel_err(g, p.grad) < 1e-12


class LinearIM(nn.Module):
    """Trainable circulant 3-tap stencil (KS/Burgers-like implicit operator)."""

    def __init__(self, N):
        super().__init__()
        self.w = nn.Parameter(torch.tensor([0.7, -1.5, 0.6], dtype=torch.float64))
        self.N = N

    def matrix(self):
        N = self.N
        eye = torch.eye(N, dtype=torch.float64, device=self.w.device)
        return self.w[0] * torch.roll(eye, -1, 1) + self.w[1] * eye + self.w[2] * torch.roll(eye, 1, 1)

    def forward(self, t, y):
        return y @ self.matrix().T


@pytest.mark.parametrize("name", ["l2", "ars122", "a2", "3", "4", "5", "1bee", "2c", "2d", "2e", "prssp2", "bpr3", "ars443"])
def test_arkimex_adjoint_matches_autograd(name):
    N, B = 8, 5
    f_im = LinearIM(N)
    f_ex = TimeMLP(d=N, hidden=12)
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(B, N, generator=g, dtype=torch.float64) * 0.5
    t = torch.tensor([0.0, 0.1, 0.2, 0.3], dtype=torch.float64)
    gout = torch.randn(4, B, N, generator=g, dtype=torch.float64)
    argv = ARGS + ["-snes_type", "ksponly", "-ts_arkimex_type", name]
    out, lam, mu, ode = _oracle_grads(dict(method="imex", imex_form=True, func2=f_ex, batch_size=B, linear_solver="torch"),
                                      [f_im, f_ex], u0, t, gout, argv=argv, step=0.1)
    At, A, bt, b, _, ct, c = otab.ARK[name].floats()
    sched = [(tt, hh, k + 1) for k, (tt, hh, ok, _) in enumerate(ode.ts.log)]
    f_im.zero_grad(), f_ex.zero_grad()
    y0 = u0.clone().requires_grad_(True)
    outs, _ = ark_unrolled_linear_im(f_im, f_ex, f_im.matrix, y0, sched, At, A, b, ct, c)
    ref = torch.stack([outs[k] for k in range(4)])
    (ref * gout).sum().backward()
    assert rel_err(out, ref) < 1e-12
    assert rel_err(lam, y0.grad) < 1e-11
    refs = [p.grad for p in list(f_im.parameters()) + list(f_ex.parameters())]
    assert len(mu) == len(refs)
    for gmu, gr in zip(mu, refs):
        assert rel_err(gmu, gr) < 1e-11


@pytest.mark.parametrize("method,theta", [("cn", 0.5), ("beuler", 1.0)])
def test_theta_adjoint_matches_implicit_differentiation(method, theta):
    func = SpiralFunc(bias_std=0.1)
    u0, _, gout = spiral_inputs(4)
    u0 = u0.reshape(-1)  # full-state dense Newton
    t = torch.tensor([0.0, 0.1, 0.2], dtype=torch.float64)
    gout = gout[:3].reshape(3, -1)

    class Flat(nn.Module):
        def __init__(self, f):
            super().__init__()
            self.f = f

        def forward(self, tt, y):
            return self.f(tt, y.view(-1, 2)).reshape(y.shape)

    ff = Flat(func)
    out, lam, mu, ode = _oracle_grads(dict(method=method, implicit_form=True), [ff], u0, t, gout, step=0.1)
    sched = [(tt, hh, k + 1) for k, (tt, hh, ok, _) in enumerate(ode.ts.log)]
    ff.zero_grad()
    y0 = u0.clone().requires_grad_(True)
    outs, _ = theta_unrolled(lambda tt, y: ff(tt, y), y0, sched, theta)
    ref = torch.stack([outs[k] for k in range(3)])
    (ref * gout).sum().backward()
    assert rel_err(out, ref) < 1e-9
    assert rel_err(lam, y0.grad) < 1e-8
    for gmu, p in zip(mu, ff.parameters()):
        assert rel_err(gmu, p.grad) < 1e-8
