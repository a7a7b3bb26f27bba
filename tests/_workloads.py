"""The three non-spiral workloads of BASELINE.json, re-implemented from their definitions in the reference (same math,
same parameter shapes and initialisation) so that parity tests and bench_configs.py can build them without importing
/root/reference:
  * CNFFunc      -- FFJORD tabular CNF right-hand side: ODEnet of ConcatSquashLinear + softplus, Hutchinson trace with a
                    fixed noise sample, state flattened as cat(z.view(-1), logp.view(-1))
                    (ffjord-pnode/lib/layers/diffeq_layers/basic.py:76-86, odefunc.py:53-57,322-385, cnf.py:72-93,140-150)
  * OdeConvBlock -- the SqueezeNext ODE block of examples-pnode/models/sqnxt_PETSc.py:70-121 (five conv+BN+ReLU layers)
  * KSImplicit / KSExplicit -- the SINODE IMEX pair of examples-sinode/KS/models/imex.py:6-70 (fixed circular 5-tap
                    stencil, 5-layer ReLU MLP returning -F(y))
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class ConcatSquashLinear(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self._layer = nn.Linear(dim_in, dim_out)
        self._hyper_bias = nn.Linear(1, dim_out, bias=False)
        self._hyper_gate = nn.Linear(1, dim_out)

    def forward(self, t, x):
        return self._layer(x) * torch.sigmoid(self._hyper_gate(t.view(1, 1))) + self._hyper_bias(t.view(1, 1))


class CNFNet(nn.Module):
    def __init__(self, dim, hidden_dims):
        super().__init__()
        dims = [dim] + list(hidden_dims) + [dim]
        self.layers = nn.ModuleList([ConcatSquashLinear(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, t, y):
        dx = y
        for i, layer in enumerate(self.layers):
            dx = layer(t, dx)
            if i < len(self.layers) - 1:
                dx = F.softplus(dx)
        return dx


class CNFFunc(nn.Module):
    """FlattenFunc(ODEfunc(ODEnet)) for states (z [B,D], logp [B,1]); `e` is the Hutchinson probe fixed per solve."""

    def __init__(self, batch, dim=6, hidden_dims=(60,), dtype=torch.float32, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        self.net = CNFNet(dim, hidden_dims).to(dtype)
        self.batch, self.dim = batch, dim
        g = torch.Generator().manual_seed(seed + 1)
        self.register_buffer("e", torch.randn(batch, dim, generator=g, dtype=torch.float64).to(dtype))
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        B, D = self.batch, self.dim
        z = y[: B * D].view(B, D)
        tt = torch.tensor(t).type_as(z)
        with torch.set_grad_enabled(True):
            z.requires_grad_(True)
            dz = self.net(tt, z)
            e_dzdx = torch.autograd.grad(dz, z, self.e, create_graph=True)[0]
            div = (e_dzdx * self.e).view(B, -1).sum(dim=1)
        return torch.cat((dz.reshape(-1), -div.reshape(-1)))


class OdeConvBlock(nn.Module):
    def __init__(self, dim, dtype=torch.float32, seed=0):
        super().__init__()
        torch.manual_seed(seed)
        r = dim // 2
        self.conv1 = nn.Conv2d(dim, r, 1, 1, bias=True)
        self.bn1 = nn.BatchNorm2d(r)
        self.conv2 = nn.Conv2d(r, r // 2, 1, 1, bias=True)
        self.bn2 = nn.BatchNorm2d(r // 2)
        self.conv3 = nn.Conv2d(r // 2, r, (1, 3), 1, (0, 1), bias=True)
        self.bn3 = nn.BatchNorm2d(r)
        self.conv4 = nn.Conv2d(r, r, (3, 1), 1, (1, 0), bias=True)
        self.bn4 = nn.BatchNorm2d(r)
        self.conv5 = nn.Conv2d(r, dim, 1, 1, bias=True)
        self.bn5 = nn.BatchNorm2d(dim)
        self.to(dtype)
        self.nfe = 0

    def forward(self, t, x):
        self.nfe += 1
        out = F.relu(self.bn1(self.conv1(x)))
        out = F.relu(self.bn2(self.conv2(out)))
        out = F.relu(self.bn3(self.conv3(out)))
        out = F.relu(self.bn4(self.conv4(out)))
        out = F.relu(self.bn5(self.conv5(out)))
        return out


class KSImplicit(nn.Module):
    """u_t = -u_xx - u_xxxx discretised with the fixed 5-point stencil of imex.py:20-36 (requires_grad=False)."""

    def __init__(self, dx, dtype=torch.float64):
        super().__init__()
        self.A = nn.Conv1d(1, 1, kernel_size=5, padding="same", padding_mode="circular", bias=False)
        K = torch.tensor([[[-1.0 / dx ** 4, 4.0 / dx ** 4 - 1.0 / dx ** 2, -6.0 / dx ** 4 + 2.0 / dx ** 2,
                            4.0 / dx ** 4 - 1.0 / dx ** 2, -1.0 / dx ** 4]]], dtype=dtype)
        self.A.weight = nn.Parameter(K, requires_grad=False)
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        return torch.squeeze(self.A(torch.unsqueeze(y, 1)), 1)


class KSExplicit(nn.Module):
    def __init__(self, n, hidden=None, dtype=torch.float64, seed=0):
        super().__init__()
        hidden = hidden or n * 25 // 8
        torch.manual_seed(seed)
        self.F = nn.Sequential(nn.Linear(n, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                               nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                               nn.Linear(hidden, n))
        for m in self.F.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0, std=0.01)
        self.to(dtype)
        self.nfe = 0

    def forward(self, t, y):
        self.nfe += 1
        return -self.F(y)


def ks_dx(n):
    return 22.0 / n
