"""A whole forward + discrete-adjoint pass of a fixed-step solve captured in a CUDA graph (torch.cuda.graph around the user's
own training step) and replayed on new inputs: the fixed-step paths launch from the host without reading the device, so
the capture holds every kernel of csrc/*.cu that the eager pass launches, and the replay reproduces the eager results bit
for bit.  (An adaptive solve decides on the device instead: its time loop is one graph of its own, csrc/cnf_rk.cu.)"""
import copy

import pytest
import torch

from pnode_b200.options import Options
from _problems import SpiralFunc, spiral_inputs
from _workloads import OdeConvBlock

pytestmark = pytest.mark.gpu


def _capture_and_replay(make_inputs, func, method, step, argv, t, expect_path, tol=0.0):
    from pnode import petsc_adjoint

    Options.clear_all()
    Options.insert_args(argv)
    f = copy.deepcopy(func).cuda()
    ode = petsc_adjoint.ODEPetsc()
    u_static = make_inputs(0).cuda()
    g_static = torch.randn((len(t),) + tuple(u_static.shape), generator=torch.Generator().manual_seed(5),
                           dtype=torch.float64).to(u_static.dtype).cuda()
    ode.setupTS(u_static, f, step_size=step, method=method, enable_adjoint=True)
    tt = t.cuda()

    def one_pass():
        for p in f.parameters():
            p.grad = None
        y0 = u_static.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, tt)
        (out * g_static).sum().backward()
        return out.detach(), y0.grad, [p.grad for p in f.parameters()]

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            one_pass()
    torch.cuda.current_stream().wait_stream(s)
    assert ode.path == expect_path
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = one_pass()
    for seed in (1, 2):
        u_static.copy_(make_inputs(seed).cuda())  # new data in the captured input buffer
        graph.replay()
        torch.cuda.synchronize()
        got = [captured[0].clone(), captured[1].clone(), [g.clone() for g in captured[2]]]
        want = one_pass()
        torch.cuda.synchronize()
        again = one_pass()
        torch.cuda.synchronize()
        pairs = [(got[0], want[0]), (got[1], want[1])] + list(zip(got[2], want[2]))
        eager = [(again[0], want[0]), (again[1], want[1])] + list(zip(again[2], want[2]))
        # conv biases in front of a train-mode BatchNorm have an exactly-zero gradient: what the kernels return there is
        # rounding noise, so every parameter gradient is measured against the scale of the largest one
        gscale = max(float(b.abs().max()) for _, b in pairs[2:])
        report = []
        for i, ((a, b), (c, d)) in enumerate(zip(pairs, eager)):
            floor = 1e-30 if i < 2 else 0.03 * gscale
            err = float((a - b).abs().max()) / max(float(b.abs().max()), floor)
            err_eager = float((c - d).abs().max()) / max(float(d.abs().max()), floor)
            report.append((i, tuple(a.shape), "%.2e" % err, "eager vs eager %.2e" % err_eager))
        assert all(float(r[2]) <= tol for r in report), report


def test_fused_spiral_pass_replays_from_a_cuda_graph():
    func = SpiralFunc(dtype=torch.float64)
    t = spiral_inputs(64, dtype=torch.float64)[1]
    _capture_and_replay(lambda seed: spiral_inputs(64, dtype=torch.float64, seed=seed)[0], func, "rk4", 0.025,
                        ["-ts_adapt_type", "none"], t, "fused-mlp-rk")


def test_conv_block_pass_replays_from_a_cuda_graph():
    """Generic engine + the hand-written conv evaluator: ~90 launches of csrc/conv_block.cu / vecops.cu per pass."""
    func = OdeConvBlock(32, dtype=torch.float32, seed=1)
    func.train()

    def inputs(seed):
        return torch.randn(8, 32, 16, 16, generator=torch.Generator().manual_seed(seed), dtype=torch.float64).float()

    # the weight-gradient partials fold into mu in the order their kernels finish: equal to fp32 rounding, not bit for bit
    _capture_and_replay(inputs, func, "rk4", 0.5, ["-ts_adapt_type", "none"], torch.tensor([0.0, 1.0], dtype=torch.float64),
                        "generic+convblock-rhs", tol=1e-3)


def _cnf_case(B, seed):
    from _workloads import CNFFunc, cnf_to

    func = cnf_to(CNFFunc(B, 6, (60,), dtype=torch.float64, seed=3), "cuda")
    with torch.no_grad():
        for prm in func.parameters():
            prm.mul_(3.0)  # enough stiffness for rejected attempts
    g = torch.Generator().manual_seed(seed)
    u0 = torch.cat((torch.randn(B, 6, generator=g, dtype=torch.float64).view(-1), torch.zeros(B, dtype=torch.float64))).cuda()
    gout = torch.randn(3, B * 7, generator=g, dtype=torch.float64).cuda()
    return func, u0, gout


def test_adaptive_cnf_solve_replays_from_a_cuda_graph():
    """An ADAPTIVE solve (dopri5, rejections, three output times) recorded into the caller's CUDA graph: a fixed budget of
    attempts with the step controller in the kernel, output states gathered on the device, the adjoint sweep reading its
    schedule from the control block (csrc/cnf_rk.cu).  Replays on new data -- which take DIFFERENT step sequences -- equal the
    eager solves bit for bit."""
    from pnode import petsc_adjoint

    B = 200
    func, u_static, gout = _cnf_case(B, 1)
    t = torch.tensor([0.0, 0.3, 1.0], dtype=torch.float64).cuda()
    Options.clear_all()
    Options.insert_args(["-ts_rtol", "1e-7", "-ts_atol", "1e-7"])
    ode = petsc_adjoint.ODEPetsc()
    ode.setupTS(u_static, func, step_size=0.5, method="dopri5", enable_adjoint=True)

    def one_pass():
        for p in func.parameters():
            p.grad = None
        y0 = u_static.clone().requires_grad_(True)
        out = ode.odeint_adjoint(y0, t)
        (out * gout).sum().backward()
        return out.detach(), y0.grad, [p.grad for p in func.parameters()]

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            one_pass()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = one_pass()
    sequences = set()
    for seed, scale in ((2, 1.0), (3, 2.5), (4, 0.2)):
        u_static.copy_(_cnf_case(B, seed)[1] * scale)
        graph.replay()
        torch.cuda.synchronize()
        got = [captured[0].clone(), captured[1].clone(), [g.clone() for g in captured[2]]]
        want = one_pass()
        torch.cuda.synchronize()
        sequences.add(tuple(a[2] for a in ode._loop.attempts))
        assert torch.isfinite(got[0]).all()
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        for a, b in zip(got[2], want[2]):
            assert torch.equal(a, b)
    assert len(sequences) > 1, "the replays should have taken different accept/reject sequences"
    assert any(False in q for q in sequences), "at least one replay should contain a rejected attempt"


def test_recorded_adaptive_solve_that_runs_out_of_attempts_returns_nan():
    from pnode import petsc_adjoint

    B = 40
    func, u_static, gout = _cnf_case(B, 1)
    t = torch.tensor([0.0, 1.0], dtype=torch.float64).cuda()
    Options.clear_all()
    Options.insert_args(["-ts_rtol", "1e-8", "-ts_atol", "1e-8", "-pnode_capture_attempts", "2"])
    ode = petsc_adjoint.ODEPetsc()
    ode.setupTS(u_static, func, step_size=0.05, method="dopri5", enable_adjoint=False)
    with torch.no_grad():
        ode.odeint(u_static, t)  # eager warm-up (also allocates the library's scratch buffers outside the graph)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = ode.odeint(u_static, t)
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[0], u_static) and torch.isnan(out[1]).all()
