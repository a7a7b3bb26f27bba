"""Problem definitions shared by the parity tests.

ROBER: the reference's own known-answer fixture (/root/reference/tests/test_pnode.py:14-124) -- same rate constants,
same output times, same per-step step_size list, same SciPy BDF ground truth.
Spiral: the model of /root/reference/examples-pnode/ode_demo_petsc.py:207-230 with the synthetic inputs of SURVEY.md 8d.
"""
import numpy as np
import torch
import torch.nn as nn
from scipy.integrate import solve_ivp

ROBER_T = torch.cat((torch.tensor([0], dtype=torch.float64), torch.logspace(start=-5, end=-3, steps=3, dtype=torch.float64)))
ROBER_STEPS = (ROBER_T[1:] - ROBER_T[:-1]).tolist()
PETSC_ARGS = ["-ts_adapt_type", "none", "-ts_trajectory_type", "memory"]


def _rober_rhs(t, s):
    k1, k2, k3 = 0.04, 3e7, 1e4
    return np.array([-k1 * s[0] + k3 * s[1] * s[2], k1 * s[0] - k3 * s[1] * s[2] - k2 * s[1] ** 2, k2 * s[1] ** 2])


def _rober_jac(t, s):
    k1, k2, k3 = 0.04, 3e7, 1e4
    return np.array([[-k1, k3 * s[2], k3 * s[1]], [k1, -2.0 * k2 * s[1] - k3 * s[2], -k3 * s[1]], [0, 2.0 * k2 * s[1], 0]])


_TRUE = None


def rober_truth():
    global _TRUE
    if _TRUE is None:
        sol = solve_ivp(_rober_rhs, [0, 1.1e-3], [1.0, 0.0, 0.0], t_eval=ROBER_T.numpy(), method="BDF", jac=_rober_jac,
                        rtol=1e-11, atol=1e-14)
        _TRUE = torch.from_numpy(sol.y.T)
    return _TRUE.clone()


class Rober(nn.Module):
    def __init__(self):
        super().__init__()
        self.k = nn.Parameter(torch.tensor([0.05, 4e7, 2e4], dtype=torch.float64))

    def forward(self, t, y):
        k1, k2, k3 = self.k[0], self.k[1], self.k[2]
        f1 = -k1 * y[0] + k3 * y[1] * y[2]
        f2 = k1 * y[0] - k3 * y[1] * y[2] - k2 * y[1] ** 2
        f3 = k2 * y[1] ** 2
        return torch.stack((f1, f2, f3), -1)


class RoberIM(nn.Module):
    def __init__(self):
        super().__init__()
        self.k1 = nn.Parameter(torch.tensor([0.05], dtype=torch.float64))
        self.k3 = nn.Parameter(torch.tensor([2e4], dtype=torch.float64))

    def forward(self, t, y):
        k1, k3 = self.k1[0], self.k3[0]
        f1 = -k1 * y[0] + k3 * y[1] * y[2]
        f2 = k1 * y[0] - k3 * y[1] * y[2]
        return torch.stack((f1, f2, torch.zeros_like(f1)), -1)


class RoberEX(nn.Module):
    def __init__(self):
        super().__init__()
        self.k2 = nn.Parameter(torch.tensor([4e7], dtype=torch.float64))

    def forward(self, t, y):
        k2 = self.k2[0]
        f3 = k2 * y[1] ** 2
        return torch.stack((torch.zeros_like(f3), -f3, f3), -1)


class SpiralFunc(nn.Module):
    """ODEFunc of ode_demo_petsc.py:207-230: Linear(2,50)-Tanh-Linear(50,2) on y**3, weights N(0, 0.1^2), biases 0."""

    def __init__(self, dtype=torch.float64, seed=0, hidden=50, cube=True, bias_std=0.0, dim=2):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.Tanh(), nn.Linear(hidden, dim)).to(dtype)
        g = torch.Generator().manual_seed(seed)
        for m in self.net.modules():
            if isinstance(m, nn.Linear):
                with torch.no_grad():
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g, dtype=torch.float64) * 0.1)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g, dtype=torch.float64) * bias_std)
        self.cube = cube
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        return self.net(y ** 3 if self.cube else y)


def spiral_inputs(batch, T=10, dtype=torch.float64, seed=0, h=0.025, dim=2):
    g = torch.Generator().manual_seed(seed)
    u0 = (torch.rand(batch, 1, dim, generator=g, dtype=torch.float64) * 2 - 1) * 2
    t = torch.arange(T, dtype=torch.float64) * h
    gout = torch.randn(T, batch, 1, dim, generator=g, dtype=torch.float64)
    return u0.to(dtype), t, gout.to(dtype)


class TimeMLP(nn.Module):
    """A generic time-dependent RHS that no fused recogniser matches (exercises the generic path)."""

    def __init__(self, d=6, hidden=16, dtype=torch.float64, seed=1):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.l1 = nn.Linear(d + 1, hidden).to(dtype)
        self.l2 = nn.Linear(hidden, d).to(dtype)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float64) * 0.4)

    def forward(self, t, y):
        tt = torch.full(y.shape[:-1] + (1,), float(t), dtype=y.dtype, device=y.device)
        return self.l2(torch.nn.functional.softplus(self.l1(torch.cat((y, tt), -1))))


def clone_module(mod, device=None, dtype=None):
    import copy

    m = copy.deepcopy(mod)
    if device is not None or dtype is not None:
        m = m.to(device=device, dtype=dtype)
    return m


def rel_err(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)
