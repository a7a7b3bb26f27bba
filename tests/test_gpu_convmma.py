"""Tensor-core evaluator of the conv ODE block (csrc/conv_mma.cu: implicit GEMMs through tcgen05, 3xTF32) against the stock
torch module evaluated in FLOAT64 and its autograd backward, on the GEMM-sized CIFAR shapes (blocks 3 and 4 of
examples-pnode/models/sqnxt_PETSc.py:70-121) and on small / ragged shapes forced onto it.  Bar: 1e-4 (fp32)."""
import copy

import pytest
import torch

from _problems import rel_err
from _workloads import OdeConvBlock

pytestmark = pytest.mark.gpu


def _callbacks(func, shape):
    from pnode_b200.convblock import ConvBlockCallbacks
    from pnode_b200.options import Options

    Options.clear_all()
    Options.insert_args(["-pnode_convblock_native", "1", "-pnode_convblock_mma", "1"])
    cb = ConvBlockCallbacks(func, torch.Size(shape))
    assert cb.native and cb.mma, "the tensor-core evaluator must accept this shape"
    return cb


def _setup(shape, seed=0, kink_free=False):
    """kink_free: BatchNorm shifts of +8 standard deviations keep every ReLU input away from 0.  A unit whose input lies within
    fp32 rounding of the kink takes the other branch in fp32 than in fp64; ONE such unit moves J^T w at a few hundred entries
    by O(1) (3e-3 of its norm at [256,128,8,8]), and with ~6 M units per evaluation about one is expected -- whatever the fp32
    implementation (torch's own differs from its fp64 self in the same way).  Full-size comparisons are therefore made in the
    kink-free regime (all arithmetic exercised, no branch decisions), the ReLU logic on shapes small enough to have no such unit."""
    N, Cc, H, W = shape
    func = OdeConvBlock(Cc, dtype=torch.float32, seed=seed).cuda()
    with torch.no_grad():
        g = torch.Generator().manual_seed(7)
        for m in func.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(0.3 * torch.randn(m.num_features, generator=g))
                if kink_free:
                    m.bias.copy_(8.0 * m.weight)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(shape, generator=g).cuda()
    w = torch.randn(shape, generator=g).cuda()
    ref = copy.deepcopy(func).double()
    xr = x.double().clone().requires_grad_(True)
    out = ref(0.0, xr)
    out.backward(w.double())
    return func, x, w, out.detach(), xr.grad, [p.grad for p in ref.parameters()], ref


SHAPES = [(8, 128, 8, 8), (16, 256, 4, 4), (4, 64, 16, 16), (3, 32, 6, 8), (2, 32, 4, 4), (5, 64, 2, 16), (256, 128, 8, 8),
          (256, 256, 4, 4)]


@pytest.mark.parametrize("kink_free", [True, False])
@pytest.mark.parametrize("shape", SHAPES)
def test_rhs_vjp_and_parameter_gradients(shape, kink_free):
    func, x, w, out_r, vu_r, gp_r, ref = _setup(shape, seed=shape[0], kink_free=kink_free)
    mine = copy.deepcopy(func)
    cb = _callbacks(mine, shape)
    cb.begin(True)
    out = cb.f(0.0, x.reshape(-1)).view(shape)
    vu, gp = cb.vjp(0.0, x.reshape(-1), w.reshape(-1))
    assert cb.reused_activations == 1
    cb.begin(True)  # forget the activation set: the same VJP re-evaluates the forward inside the call, bit-identically
    vu2, gp2 = cb.vjp(0.0, x.reshape(-1), w.reshape(-1))
    assert torch.equal(vu, vu2) and all(torch.equal(a, b) for a, b in zip(gp, gp2))
    tol = 1e-4
    assert rel_err(out, out_r) < tol, ("f", rel_err(out, out_r))
    if not kink_free:
        # stock initialisation: with more than ~1e5 ReLU units per evaluation one of them is likely to sit within fp32 rounding
        # of the kink (see _setup): the derivative bars then allow for that unit's branch
        units = shape[0] * shape[2] * shape[3] * sum(c.out_channels for c in [mine.conv1, mine.conv2, mine.conv3, mine.conv4, mine.conv5])
        tol = 1e-4 if units < 1e5 else 2e-2
    assert rel_err(vu.view(shape), vu_r) < tol, ("J^T w", rel_err(vu.view(shape), vu_r))
    flat = lambda gs: torch.cat([q.detach().double().reshape(-1) for q in gs])
    assert rel_err(flat(gp), flat(gp_r)) < tol, ("Jp^T w", rel_err(flat(gp), flat(gp_r)))
    gscale = float(flat(gp_r).norm()) / len(gp_r) ** 0.5
    for (n, _), a, b in zip(mine.named_parameters(), gp, gp_r):
        if "conv" in n and n.endswith("bias"):
            assert float(a.abs().max()) == 0.0  # exactly zero: the bias cancels in the BatchNorm that follows
            continue
        # a gradient that vanishes analytically (BatchNorm shifts when every unit is active) is rounding noise on both sides:
        # errors are measured against the parameter's own gradient or 1e-3 of the typical gradient, whichever is larger
        err = float((a.view_as(b).double() - b).norm()) / max(float(b.norm()), 1e-3 * gscale)
        assert err < 3 * tol, (n, err)
    # BatchNorm side effects: f once + each vjp once (re-evaluated or replayed from the stored statistics)
    assert int(mine.bn3.num_batches_tracked) == 3
    again = copy.deepcopy(func)
    again(0.0, x), again(0.0, x), again(0.0, x)
    for k in range(1, 6):
        a, b = getattr(mine, "bn%d" % k), getattr(again, "bn%d" % k)
        # batch means are ~1e-2 of the activations' spread: their own relative error is that much larger than the data's
        assert rel_err(a.running_mean, b.running_mean) < 1e-3 and rel_err(a.running_var, b.running_var) < 1e-3


def test_stage_combination_and_mu_accumulation_are_fused():
    shape = (8, 128, 8, 8)
    func, x, w, out_r, vu_r, gp_r, _ = _setup(shape)
    cb = _callbacks(copy.deepcopy(func), shape)
    cb.begin(True)
    g = torch.Generator().manual_seed(5)
    base = torch.randn(shape, generator=g).cuda()
    y, k = torch.empty_like(x), torch.empty_like(x)
    cb._native_f(x.reshape(-1), out=y, base=base, base_coef=1.0, k_coef=0.5, k=k)
    assert rel_err(k, out_r) < 1e-4 and rel_err(y, base.double() + 0.5 * out_r) < 1e-4
    mu0 = torch.randn(cb.nparams, generator=g).cuda()
    mu = mu0.clone()
    vu, none = cb.vjp_accumulate(0.0, x.reshape(-1), w.reshape(-1), mu, 0.25)
    want = mu0.double() + 0.25 * torch.cat([q.reshape(-1) for q in gp_r])
    assert none is None and rel_err(vu.view(shape), vu_r) < 1e-4 and rel_err(mu, want) < 1e-4
    mu2 = mu0.clone()
    vu2, _ = cb.vjp_accumulate(0.0, x.reshape(-1), w.reshape(-1), mu2, 0.25)
    assert torch.equal(vu, vu2) and torch.equal(mu, mu2)  # fixed-order reductions: bit-reproducible


@pytest.mark.parametrize("kink_free", [True, False])
@pytest.mark.parametrize("shape", [(256, 128, 8, 8), (256, 256, 4, 4)])
def test_cifar_blocks_3_and_4_through_the_drop_in(shape, kink_free):
    """RK4, t=[1.0], one step through ODEPetsc: the tensor-core evaluator is chosen on its own, results against the oracle.
    Kink-free regime (see _setup): 1e-4 on trajectory, lambda and mu.  Stock initialisation: the trajectory holds 1e-4 (ReLU is
    continuous), lambda and mu carry the O(1e-3) effect of the expected ~1 unit per evaluation that sits on the kink."""
    from oracle import OracleODEPetsc
    from pnode import petsc_adjoint
    from pnode_b200.options import Options

    Options.clear_all()
    Options.insert_args(["-ts_adapt_type", "none"])
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(shape, generator=g)
    gout = torch.randn((1,) + shape, generator=g)
    t = torch.tensor([1.0], dtype=torch.float64)
    res = []
    for dev in ("cpu", "cuda"):
        func = OdeConvBlock(shape[1])
        if kink_free:
            with torch.no_grad():
                for m in func.modules():
                    if isinstance(m, torch.nn.BatchNorm2d):
                        m.bias.copy_(8.0 * m.weight)
        func = func.to(dev)
        ode = OracleODEPetsc(["-ts_adapt_type", "none"]) if dev == "cpu" else petsc_adjoint.ODEPetsc()
        ode.setupTS(u0.to(dev), func, step_size=1.0, method="rk4")
        y0 = u0.to(dev).clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t.to(dev))
        (out * gout.to(dev)).sum().backward()
        res.append((out.detach().cpu(), y0.grad.cpu(), torch.cat([p.grad.reshape(-1) for p in func.parameters()]).cpu(), ode))
    o, p = res
    assert p[3].path == "generic+convblock-rhs" and p[3]._cb_im.mma
    errs = [rel_err(a, b) for a, b in zip(p[:3], o[:3])]
    assert errs[0] < 1e-4 and max(errs) < (1e-4 if kink_free else 2e-2), errs
