"""GPU parity of the generic-path vector kernels (through the C ABI) against the oracle arithmetic on the same inputs."""
import math

import pytest
import torch

from oracle import wrms_norm

pytestmark = pytest.mark.gpu


def _ops(dtype):
    from pnode_b200.device import DeviceOps

    return DeviceOps(torch.device("cuda:0"), dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n", [0, 1, 3, 40, 1023, 4096 + 5, 2 * 1024 * 1024 + 1])
def test_lincomb(dtype, n):
    g = torch.Generator().manual_seed(n + 1)
    ops = _ops(dtype)
    for nterms in (0, 1, 6, 16, 19):
        vecs = [torch.randn(n, generator=g, dtype=torch.float64) for _ in range(nterms)]
        coefs = [float(c) for c in torch.randn(nterms, generator=g, dtype=torch.float64)]
        base = torch.randn(n, generator=g, dtype=torch.float64)
        ref = 0.75 * base.to(dtype).double()
        for v, c in zip(vecs, coefs):
            ref = ref + float(torch.tensor(c, dtype=dtype)) * v.to(dtype).double()
        out = torch.empty(n, dtype=dtype, device="cuda")
        ops.lincomb(out, base.to(dtype).cuda(), 0.75, [v.to(dtype).cuda() for v in vecs], coefs)
        tol = 1e-14 if dtype == torch.float64 else 2e-6
        assert torch.allclose(out.cpu().double(), ref, rtol=tol, atol=tol * 10)
        out2 = torch.empty(n, dtype=dtype, device="cuda")
        ops.lincomb(out2, None, 0.0, [v.to(dtype).cuda() for v in vecs], coefs)
        assert torch.allclose(out2.cpu().double(), ref - 0.75 * base.to(dtype).double(), rtol=tol, atol=tol * 10)


def test_lincomb_unaligned_views_and_aliasing():
    ops = _ops(torch.float64)
    big = torch.randn(4099, dtype=torch.float64, device="cuda")
    a, b = big[1:2050], big[2050:4099]  # 8-byte but not 16-byte aligned
    ref = a.cpu() + 2.0 * b.cpu()
    ops.lincomb(a, a, 1.0, [b], [2.0])  # in place
    assert torch.allclose(a.cpu(), ref, rtol=1e-15, atol=0)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n", [7, 7000, 3 * 1024 * 1024 + 3])
def test_complete_with_weighted_norm(dtype, n):
    g = torch.Generator().manual_seed(n)
    ops = _ops(dtype)
    u = torch.randn(n, generator=g, dtype=torch.float64).to(dtype)
    ks = [torch.randn(n, generator=g, dtype=torch.float64).to(dtype) for _ in range(7)]
    bw = [0.05 * (j + 1) for j in range(7)]
    ew = [1e-4 * (-1) ** j * (j + 1) for j in range(7)]
    unew_ref = u.double()
    err = torch.zeros(n, dtype=torch.float64)
    for k, b, e in zip(ks, bw, ew):
        unew_ref = unew_ref + float(torch.tensor(b, dtype=dtype)) * k.double()
        err = err + float(torch.tensor(e, dtype=dtype)) * k.double()
    unew = torch.empty(n, dtype=dtype, device="cuda")
    for rep in range(3):  # the work buffer (ticket) must be restored by the kernel
        sumsq = ops.complete(unew, u.cuda(), [k.cuda() for k in ks], bw, ew, 1e-4, 1e-3)
        got = math.sqrt(float(sumsq.item()) / n)
    tol = 1e-13 if dtype == torch.float64 else 3e-6
    assert torch.allclose(unew.cpu().double(), unew_ref, rtol=tol, atol=tol)
    ref = wrms_norm(unew_ref, unew_ref + err, 1e-4, 1e-3)
    assert got == pytest.approx(ref, rel=1e-9 if dtype == torch.float64 else 2e-3)
    # plain completion
    ops.complete(unew, u.cuda(), [k.cuda() for k in ks], bw)
    assert torch.allclose(unew.cpu().double(), unew_ref, rtol=tol, atol=tol)
    # bit-reproducible reduction
    s1 = ops.complete(unew, u.cuda(), [k.cuda() for k in ks], bw, ew, 1e-4, 1e-3).clone()
    s2 = ops.complete(unew, u.cuda(), [k.cuda() for k in ks], bw, ew, 1e-4, 1e-3).clone()
    assert torch.equal(s1, s2)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_multi_axpy(dtype):
    g = torch.Generator().manual_seed(11)
    ops = _ops(dtype)
    sizes = [100, 50, 0, 1, 7, 3200 * 33] + [5] * 40
    grads = [None if i == 4 else torch.randn(s, generator=g, dtype=torch.float64).to(dtype) for i, s in enumerate(sizes)]
    mu0 = torch.randn(sum(sizes), generator=g, dtype=torch.float64).to(dtype)
    ref = mu0.double().clone()
    off = 0
    for gr, s in zip(grads, sizes):
        if gr is not None:
            ref[off:off + s] += float(torch.tensor(0.3, dtype=dtype)) * gr.double()
        off += s
    mu = mu0.cuda()
    ops.multi_axpy(mu, [None if x is None else x.cuda() for x in grads], sizes, 0.3)
    tol = 1e-14 if dtype == torch.float64 else 2e-6
    assert torch.allclose(mu.cpu().double(), ref, rtol=tol, atol=tol)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 8e-16), (torch.float32, 4e-7)])
def test_kernel_tanh_accuracy(dtype, tol):
    import ctypes as C

    from pnode_b200 import _lib

    lib = _lib.load()
    x = torch.cat((torch.linspace(-25, 25, 200001, dtype=torch.float64), torch.logspace(-300, 1, 4001, dtype=torch.float64),
                   -torch.logspace(-30, 1, 4001, dtype=torch.float64), torch.tensor([0.0, 1e3, -1e3, 19.9, 20.1, 23.9, 24.1, -24.1, 1e300, -1e300, float('inf'), -float('inf')])))
    xd = x.to(dtype).cuda()
    out = torch.empty_like(xd)
    _lib.check(lib.pnode_tanh_probe(xd.data_ptr(), out.data_ptr(), xd.numel(), 0 if dtype == torch.float32 else 1,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    ref = torch.tanh(xd.double().cpu())
    err = (out.cpu().double() - ref).abs()
    assert float(err.max()) < tol  # absolute error
    # the fp64 evaluation is 1 - 2 / (e^{2x} + 1): absolute accuracy (what the state and the gradients see through W2 a and
    # 1 - a^2), exact at 0, never outside [-1, 1]
    o = out.cpu().double()
    assert float(o.abs().max()) <= 1.0 and float(o[x == 0.0].abs().max()) == 0.0
    small = 1e-10 if dtype == torch.float64 else 1e-4
    assert bool((o[x > small] > 0).all()) and bool((o[x < -small] < 0).all())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("n", [5, 4097, 1 << 20])
def test_mdot_matches_fp64_dots(dtype, n):
    g = torch.Generator().manual_seed(n)
    ops = _ops(dtype)
    w = torch.randn(n, generator=g, dtype=torch.float64).to(dtype)
    for nv in (0, 1, 7, 16, 31):
        vs = [torch.randn(n, generator=g, dtype=torch.float64).to(dtype) for _ in range(nv)]
        for rep in range(2):  # ticket restored between calls
            vals, ww = ops.mdot([v.cuda() for v in vs], w.cuda())
        ref = [float((v.double() * w.double()).sum()) for v in vs]
        assert len(vals) == nv
        scale = float(w.double().norm()) * (float(vs[0].double().norm()) if nv else 1.0)
        for a, b in zip(vals, ref):
            assert abs(a - b) <= 1e-12 * scale
        assert ww == pytest.approx(float((w.double() ** 2).sum()), rel=1e-12)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("nseg,seglen,nvec", [(1, 3, 2), (5, 1, 0), (37, 1000, 17), (256, 1024, 30), (3, 70001, 4)])
def test_segmented_dots_and_updates(dtype, nseg, seglen, nvec):
    """pnode_mdot_seg / pnode_lincomb_seg (block Krylov solver): per-segment dots in double, coefficients read from the
    device in the three modes, more than 16 vectors per call, out aliasing base."""
    ops = _ops(dtype)
    g = torch.Generator().manual_seed(nseg + seglen)
    vecs = [torch.randn(nseg * seglen, generator=g, dtype=torch.float64).to(dtype).cuda() for _ in range(nvec)]
    w = torch.randn(nseg * seglen, generator=g, dtype=torch.float64).to(dtype).cuda()
    got = ops.mdot_seg(vecs, w, nseg)
    wd = w.double().view(nseg, seglen)
    want = torch.stack([(v.double().view(nseg, seglen) * wd).sum(1) for v in vecs] + [(wd * wd).sum(1)])
    assert got.shape == (nvec + 1, nseg)
    scale = wd.norm(dim=1) * torch.stack([v.double().view(nseg, seglen).norm(dim=1) for v in vecs] + [wd.norm(dim=1)])
    assert float(((got - want).abs() / scale).max()) < 1e-14
    eps = 1e-15 if dtype == torch.float64 else 1e-6
    # Gram-Schmidt update with the dots just computed (NEG), coefficients as is (COEF), normalisation (RSQRT)
    out = torch.empty_like(w)
    ops.lincomb_seg(out, w, 1.0, vecs, got, 1, nseg)
    ref = wd - sum(want[j][:, None] * vecs[j].double().view(nseg, seglen) for j in range(nvec))
    assert float((out.double().view(nseg, seglen) - ref).abs().max() / ref.abs().max()) < 50 * eps * max(nvec, 1)
    acc = w.clone()
    ops.lincomb_seg(acc, acc, 0.5, vecs, got, 0, nseg)
    ref = 0.5 * wd + sum(want[j][:, None] * vecs[j].double().view(nseg, seglen) for j in range(nvec))
    assert float((acc.double().view(nseg, seglen) - ref).abs().max() / ref.abs().max()) < 50 * eps * max(nvec, 1)
    nn = got[-1:].clone()
    nn[0, 0] = 0.0  # a segment that broke down: zero basis vector
    ops.lincomb_seg(out, None, 0.0, [w], nn, 2, nseg)
    ref = wd / wd.norm(dim=1, keepdim=True)
    ref[0] = 0.0
    assert float((out.double().view(nseg, seglen) - ref).abs().max()) < 10 * eps
