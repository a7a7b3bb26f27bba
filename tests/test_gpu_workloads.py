"""BASELINE configs 3-5 through the drop-in on the GPU (generic path: torch evaluates the RHS / VJP, csrc/vecops.cu does
the TS arithmetic) against the oracle on the same seeded inputs, at sizes the CPU oracle finishes in seconds."""
import copy

import pytest
import torch

from oracle import OracleODEPetsc
from pnode_b200.options import Options
from _problems import rel_err
from _workloads import CNFFunc, KSExplicit, KSImplicit, OdeConvBlock, cnf_to, ks_dx

pytestmark = pytest.mark.gpu


def _pair(argv, funcs, kw, u0, t, gout, step):
    from pnode import petsc_adjoint

    res = []
    for dev, make in (("cpu", lambda: OracleODEPetsc(argv)), ("cuda", lambda: petsc_adjoint.ODEPetsc())):
        Options.clear_all()
        Options.insert_args(argv)
        fs = [cnf_to(copy.deepcopy(f), dev) if hasattr(f, "base_func") else copy.deepcopy(f).to(dev) for f in funcs]
        k = dict(kw)
        if len(fs) == 2:
            k["func2"] = fs[1]
        ode = make()
        ode.setupTS(u0.to(dev), fs[0], step_size=step, enable_adjoint=True, **k)
        y0 = u0.to(dev).clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t.to(dev))
        (out * gout.to(dev)).sum().backward()
        grads = [p.grad for f in fs for p in f.parameters() if p.requires_grad]
        res.append((out.detach(), y0.grad, grads, ode, fs))
    return res


def _compare(p, o, tol, per_param=True):
    assert rel_err(p[0], o[0]) < tol, ("trajectory", rel_err(p[0], o[0]))
    assert rel_err(p[1], o[1]) < tol, ("lambda", rel_err(p[1], o[1]))
    assert len(p[2]) == len(o[2]) and len(p[2]) > 0
    flat = lambda gs: torch.cat([g.detach().double().cpu().reshape(-1) for g in gs])
    assert rel_err(flat(p[2]), flat(o[2])) < tol, ("mu (whole vector)", rel_err(flat(p[2]), flat(o[2])))
    if per_param:
        for a, b in zip(p[2], o[2]):
            assert rel_err(a, b) < tol, ("mu", rel_err(a, b))


@pytest.mark.parametrize("dtype,tol,ts_tol", [(torch.float64, 1e-10, "1e-6"), (torch.float32, 1e-4, "1e-4")])
def test_config3_ffjord_cnf_dopri5_adaptive(dtype, tol, ts_tol):
    """POWER-shaped 6-D CNF, hidden 60, B=1000, t=[0,1], dopri5 adaptive from h=0.05, Hutchinson VJP (second order)."""
    B, D = 1000, 6
    func = CNFFunc(B, D, (60,), dtype=dtype)
    g = torch.Generator().manual_seed(2)
    z = torch.randn(B, D, generator=g, dtype=torch.float64)
    u0 = torch.cat((z.view(-1), torch.zeros(B, dtype=torch.float64))).to(dtype)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    gout = torch.randn(2, B * (D + 1), generator=g, dtype=torch.float64).to(dtype)
    o, p = _pair(["-ts_rtol", ts_tol, "-ts_atol", ts_tol], [func], dict(method="dopri5"), u0, t, gout, 0.05)
    lo, lp = o[3].ts.log, p[3]._loop.attempts
    assert [a[2] for a in lo] == [a[2] for a in lp] and len(lo) >= 3
    for a, b in zip(lo, lp):
        assert a[1] == pytest.approx(b[1], rel=1e-8 if dtype == torch.float64 else 2e-2)
    assert p[3].np == 984  # np = 984 for the FFJORD defaults
    _compare(p, o, tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_config4_cifar_ode_block_rk4(dtype, tol):
    """SqueezeNext ODE block (conv + BatchNorm in train mode), RK4, t=[1.0] single point, Nt=2 => h=0.5."""
    C, HW, B = 32, 8, 16
    func = OdeConvBlock(C, dtype=dtype)
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(B, C, HW, HW, generator=g, dtype=torch.float64).to(dtype)
    gout = torch.randn(1, B, C, HW, HW, generator=g, dtype=torch.float64).to(dtype)
    t = torch.tensor([1.0], dtype=torch.float64)
    # parity needs IEEE fp32 convolutions: cuDNN's default TF32 path alone moves lambda by 2e-2 (measured, tools/diag_cifar.py)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        o, p = _pair(["-ts_adapt_type", "none", "-pnode_convblock_native", "1"], [func], dict(method="rk4"), u0, t, gout, 0.5)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert p[0].shape == (1, B, C, HW, HW)
    assert p[3].path == "generic+convblock-rhs" and p[3]._cb_im.native  # whole RHS / VJP in csrc/conv_block.cu
    # conv biases feeding a BatchNorm have an exactly-zero gradient (pure rounding noise on both sides): compare mu as a
    # whole vector, not parameter by parameter
    _compare(p, o, tol, per_param=False)
    # BatchNorm running statistics are mutated once per RHS evaluation, adjoint re-evaluations included
    assert p[4][0].nfe == o[4][0].nfe == 16
    assert rel_err(p[4][0].bn1.running_mean, o[4][0].bn1.running_mean) < tol * 10


@pytest.mark.parametrize("name", ["3", "l2"])
def test_config5_sinode_ks_imex(name):
    """KS: fixed stiff linear operator (implicit) + MLP (explicit), ARKIMEX, linear_solver='torch', -snes_type ksponly,
    h = 0.2, one step per call (examples-sinode/KS/runs64_a100.sh:23), fp64."""
    N, B = 64, 16
    f_im = KSImplicit(ks_dx(N))
    f_ex = KSExplicit(N)
    g = torch.Generator().manual_seed(4)
    u0 = 0.5 * torch.randn(B, N, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 0.2], dtype=torch.float64)
    gout = torch.randn(2, B, N, generator=g, dtype=torch.float64)
    argv = ["-ts_adapt_type", "none", "-snes_type", "ksponly", "-ts_arkimex_type", name]
    o, p = _pair(argv, [f_im, f_ex], dict(method="imex", imex_form=True, batch_size=B, linear_solver="torch",
                                          fixed_jacobian_across_solves=True), u0, t, gout, 0.2)
    assert p[3].npIM == 0 and p[3].npEX == 146464  # npEX = 146,464 at N = 64
    assert p[3].path == "generic+dense-mlp+circulant-rhs"  # tensor-core MLP + circulant implicit operator
    _compare(p, o, 1e-10)


def test_config5_fixed_jacobian_is_kept_across_solves_and_stage_graphs_are_reused():
    """fixed_jacobian(_across_solves)=True: the (shift I - J)^-1 factorisation survives from one odeint to the next until a
    parameter / buffer of the implicit function changes; the adjoint differentiates the stage graphs the forward kept."""
    from pnode import petsc_adjoint

    N, B = 64, 8
    f_im, f_ex = KSImplicit(ks_dx(N)).cuda(), KSExplicit(N).cuda()
    g = torch.Generator().manual_seed(4)
    u0 = (0.5 * torch.randn(B, N, generator=g, dtype=torch.float64)).cuda()
    t = torch.tensor([0.0, 0.2], dtype=torch.float64).cuda()
    gout = torch.randn(2, B, N, generator=g, dtype=torch.float64).cuda()
    Options.clear_all()
    Options.insert_args(["-ts_adapt_type", "none", "-snes_type", "ksponly", "-pnode_fused", "0"])  # the generic engine

    def solve(ode):
        f_ex.zero_grad(set_to_none=True)
        y0 = u0.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t)
        (out * gout).sum().backward()
        return out.detach().clone(), y0.grad.clone(), [p.grad.clone() for p in f_ex.parameters()]

    ode = petsc_adjoint.ODEPetsc()
    ode.setupTS(u0, f_im, step_size=0.2, method="imex", imex_form=True, func2=f_ex, batch_size=B, linear_solver="torch",
                fixed_jacobian_across_solves=True)
    a = solve(ode)
    inv1 = dict(ode._imp._inv)
    assert len(inv1) >= 1 and ode._cb_ex.reused_graphs >= 4  # ARK3: four explicit stage evaluations reused by the adjoint
    b = solve(ode)
    assert all(ode._imp._inv[k] is v for k, v in inv1.items()), "the factorisation was rebuilt although nothing changed"
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and all(torch.equal(x, y) for x, y in zip(a[2], b[2]))
    with torch.no_grad():
        f_im.A.weight.mul_(1.0)  # in-place touch: version counter moves, the Jacobian must be rebuilt
    solve(ode)
    assert all(ode._imp._inv[k] is not v for k, v in inv1.items())
    # without the promise the factorisation is rebuilt at every solve, as in the reference
    ode2 = petsc_adjoint.ODEPetsc()
    ode2.setupTS(u0, f_im, step_size=0.2, method="imex", imex_form=True, func2=f_ex, batch_size=B, linear_solver="torch")
    c = solve(ode2)
    inv2 = dict(ode2._imp._inv)
    solve(ode2)
    assert all(ode2._imp._inv[k] is not v for k, v in inv2.items())
    assert rel_err(c[0], a[0]) < 1e-12 and rel_err(c[1], a[1]) < 1e-12


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", [(256, 16, 32, 32), (7, 32, 4, 4), (3, 5, 2, 2)])
def test_bn_relu_kernels_match_torch(dtype, shape):
    """csrc/bn_relu.cu against torch's train-mode batch_norm + relu and its autograd backward (same inputs, fp64 1e-11)."""
    import ctypes as C

    from pnode_b200 import _lib

    lib = _lib.load()
    N, Cc, H, W = shape
    g = torch.Generator().manual_seed(N)
    x = torch.randn(shape, generator=g, dtype=torch.float64).to(dtype).cuda()
    dy = torch.randn(shape, generator=g, dtype=torch.float64).to(dtype).cuda()
    gamma = (torch.rand(Cc, generator=g, dtype=torch.float64) + 0.5).to(dtype).cuda()
    beta = torch.randn(Cc, generator=g, dtype=torch.float64).to(dtype).cuda()
    rm0, rv0 = torch.zeros(Cc, dtype=dtype).cuda(), torch.ones(Cc, dtype=dtype).cuda()
    # torch reference
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    yr = torch.relu(torch.nn.functional.batch_norm(xr, rm, rv, gr, br, True, 0.1, 1e-5))
    yr.backward(dy)
    # kernels
    work = torch.empty(int(lib.pnode_bn_work_bytes(Cc)), dtype=torch.uint8, device="cuda")
    y, dx = torch.empty_like(x), torch.empty_like(x)
    mean, invstd = torch.empty(Cc, dtype=dtype).cuda(), torch.empty(Cc, dtype=dtype).cuda()
    dg, db = torch.empty(Cc, dtype=dtype).cuda(), torch.empty(Cc, dtype=dtype).cuda()
    rm2, rv2 = rm0.clone(), rv0.clone()
    code = 0 if dtype == torch.float32 else 1
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pnode_bn_relu_forward(x.data_ptr(), y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm2.data_ptr(),
                                         rv2.data_ptr(), mean.data_ptr(), invstd.data_ptr(), N, Cc, H * W, 1e-5, 0.1,
                                         work.data_ptr(), code, st))
    _lib.check(lib.pnode_bn_relu_backward(dy.data_ptr(), x.data_ptr(), y.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
                                          invstd.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), N, Cc, H * W,
                                          work.data_ptr(), code, st))
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    assert rel_err(y, yr) < tol and rel_err(dx, xr.grad) < tol * 10
    assert rel_err(dg, gr.grad) < tol * 10 and rel_err(db, br.grad) < tol * 10
    assert rel_err(rm2, rm) < tol * 10 and rel_err(rv2, rv) < tol * 10


def test_config4_evaluator_choice_and_library_path():
    """auto: the hand-written conv kernels when there are pixels enough to fill the GPU, else library GEMM convolutions + the
    BatchNorm+ReLU kernels of csrc/bn_relu.cu; both evaluators give the same trajectory and gradients."""
    from pnode import petsc_adjoint
    from pnode_b200.convblock import ConvBlockCallbacks

    Options.clear_all()
    big = ConvBlockCallbacks(OdeConvBlock(32).cuda(), torch.Size((16, 32, 32, 32)))
    wide = ConvBlockCallbacks(OdeConvBlock(256).cuda(), torch.Size((256, 256, 4, 4)))
    wide64 = ConvBlockCallbacks(OdeConvBlock(256, dtype=torch.float64).cuda(), torch.Size((256, 256, 4, 4)))
    # many pixels, few channels: the CUDA-core kernels; few pixels, many channels: fp32 on the tensor cores (conv_mma.cu),
    # fp64 (no tensor-core evaluator, too few pixels for the CUDA-core one) on library convolutions + bn_relu.cu
    assert big.native and not big.mma and wide.native and wide.mma and not wide64.native
    C, HW, B = 16, 8, 8
    func = OdeConvBlock(C, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(B, C, HW, HW, generator=g, dtype=torch.float64)
    gout = torch.randn(1, B, C, HW, HW, generator=g, dtype=torch.float64)
    res = []
    for mode in ("1", "0"):
        Options.clear_all()
        Options.insert_args(["-ts_adapt_type", "none", "-pnode_convblock_native", mode])
        f = copy.deepcopy(func).cuda()
        ode = petsc_adjoint.ODEPetsc()
        ode.setupTS(u0.cuda(), f, step_size=0.5, method="rk4", enable_adjoint=True)
        y0 = u0.cuda().clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, torch.tensor([1.0], dtype=torch.float64).cuda())
        (out * gout.cuda()).sum().backward()
        res.append((out.detach(), y0.grad, [p.grad for p in f.parameters()], ode, f))
    assert res[0][3]._cb_im.native and not res[1][3]._cb_im.native and res[1][3].path == "generic+convblock-rhs"
    assert res[0][3]._cb_im.reused_activations == 8  # 2 steps x 4 stages: no forward re-evaluation in the adjoint
    _compare(res[0], res[1], 1e-9, per_param=False)


def test_config4_adaptive_dopri5_with_rejections_matches_stock_module_path():
    """The conv-block evaluator under the adaptive controller (dopri5, tight tolerance, a too-large first step so that attempts
    are rejected): same accept/reject sequence, trajectory and gradients as the stock module through autograd; activation sets
    of rejected attempts are dropped."""
    from pnode import petsc_adjoint

    C, HW, B = 16, 8, 8
    func = OdeConvBlock(C, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(B, C, HW, HW, generator=g, dtype=torch.float64)
    gout = torch.randn(2, B, C, HW, HW, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    res = []
    for extra in (["-pnode_convblock_native", "1"], ["-pnode_fused", "0"]):
        Options.clear_all()
        Options.insert_args(["-ts_rtol", "1e-7", "-ts_atol", "1e-7"] + extra)
        f = copy.deepcopy(func).cuda()
        ode = petsc_adjoint.ODEPetsc()
        ode.setupTS(u0.cuda(), f, step_size=1.0, method="dopri5", enable_adjoint=True)
        y0 = u0.cuda().clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t.cuda())
        (out * gout.cuda()).sum().backward()
        res.append((out.detach(), y0.grad, [p.grad for p in f.parameters()], ode, f))
    a, b = res
    la, lb = a[3]._loop.attempts, b[3]._loop.attempts
    assert [x[2] for x in la] == [x[2] for x in lb] and any(not x[2] for x in la) and a[3]._loop.steps >= 2
    assert a[3]._cb_im.native and len(a[3]._cb_im._saved) <= 6 * a[3]._loop.steps + 1
    _compare(a, b, 1e-8, per_param=False)


def test_config4_fused_rhs_equals_stock_module_path():
    """The conv-block evaluator vs the same module driven through torch autograd (-pnode_fused 0), both on the GPU."""
    from pnode import petsc_adjoint

    C, HW, B = 32, 8, 16
    func = OdeConvBlock(C, dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    u0 = torch.randn(B, C, HW, HW, generator=g, dtype=torch.float64)
    gout = torch.randn(2, B, C, HW, HW, generator=g, dtype=torch.float64)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64)
    res = []
    for argv in (["-ts_adapt_type", "none", "-pnode_convblock_native", "1"], ["-ts_adapt_type", "none", "-pnode_fused", "0"]):
        Options.clear_all()
        Options.insert_args(argv)
        f = copy.deepcopy(func).cuda()
        ode = petsc_adjoint.ODEPetsc()
        ode.setupTS(u0.cuda(), f, step_size=0.25, method="rk4", enable_adjoint=True)
        y0 = u0.cuda().clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t.cuda())
        (out * gout.cuda()).sum().backward()
        res.append((out.detach(), y0.grad, [p.grad for p in f.parameters()], ode, f))
    a, b = res
    assert a[3].path == "generic+convblock-rhs" and a[3]._cb_im.native and b[3].path == "generic"
    _compare(a, b, 1e-9, per_param=False)
    assert a[4].nfe == b[4].nfe == 32 and int(a[4].bn3.num_batches_tracked) == int(b[4].bn3.num_batches_tracked) == 32
    assert rel_err(a[4].bn5.running_var, b[4].bn5.running_var) < 1e-9
