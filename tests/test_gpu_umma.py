"""Tensor-core sliced products (csrc/umma_gemm.cu) through the C ABI against plain fp64 matrix products.

`python tests/test_gpu_umma.py` runs every case and prints a table without stopping at the first failure (bring-up aid)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

SHAPES = [(128, 64, 64), (128, 64, 128), (128, 64, 32), (128, 128, 256), (256, 192, 640), (100, 70, 200), (37, 5, 1000),
          (256, 1024, 3200), (256, 3200, 1024), (384, 3200, 256)]


def _ops():
    from pnode_b200 import sliced
    return sliced


def _unpack_i8(s):
    """Host reconstruction of an int8-sliced operand: x = 2^e sum_i q_i 2^-(L + B i), signed bytes; kind 0: balanced base-256
    digits (L = 6, B = 8, 6 slices), kind 2: base 128, digits in [-64, 64] (L = 6, B = 7, 8 slices)."""
    S, B = (8, 7) if s.kind == 2 else (6, 8)
    pitch = s.buf.numel() // (S * s.rows)
    q = s.buf.view(torch.int8).view(S, s.rows, pitch)[:, :, :s.k].to(torch.float64).cpu()
    e = s.exp.cpu().to(torch.float64)
    x = torch.zeros(s.rows, s.k, dtype=torch.float64)
    for i in range(S):
        x += q[i] * 2.0 ** (-(6 + B * i))
    ok = bool(q.abs().max() <= 64) if s.kind == 2 else bool(q[0].abs().max() <= 65)
    return x * (2.0 ** e)[:, None], ok, e


def case_slices(dtype):
    sl = _ops()
    torch.manual_seed(1)
    x = (torch.randn(50, 333, dtype=torch.float64) * torch.logspace(-6, 3, 50, dtype=torch.float64)[:, None]).to(dtype).cuda()
    x[7] = 0.0
    if dtype == torch.float64:
        errs, ok = [], True
        for ext in (False, True):
            xr, okq, e = _unpack_i8(sl.slice_rows(x, extended=ext))
            amax = x.abs().amax(1).cpu()
            ok = ok and okq and bool((amax < 2.0 ** e).all())
            errs.append(((xr - x.cpu()).abs().amax(1) / (2.0 ** e)).max().item())
            xc, okq, e = _unpack_i8(sl.slice_cols(x, extended=ext))
            ok = ok and okq
            errs.append(((xc - x.cpu().T).abs().amax(1) / (2.0 ** e)).max().item())
        return max(errs), 2.0 ** -47, ok
    s = sl.slice_rows(x)
    pitch = s.buf.numel() // (2 * s.rows)
    parts = s.buf.view(2, s.rows, pitch)[:, :, :s.k * 4].contiguous().view(torch.float32).view(2, s.rows, s.k)
    err_r = (parts[0].double() + parts[1].double() - x.double()).abs().max().item()
    lowbits = int((parts[0].contiguous().view(torch.int32) & 0x1fff).abs().max().item())
    s = sl.slice_cols(x)
    pitch = s.buf.numel() // (2 * s.rows)
    parts = s.buf.view(2, s.rows, pitch)[:, :, :s.k * 4].contiguous().view(torch.float32).view(2, s.rows, s.k)
    err_c = (parts[0].double() + parts[1].double() - x.double().T).abs().max().item()
    return max(err_r, err_c), 0.0, lowbits == 0


def case_exact_int(M, N, K):
    """Integer-valued fp64 operands below 64: slice 0 holds them exactly, the product must be bit-exact."""
    sl = _ops()
    g = torch.Generator().manual_seed(M * 131 + N * 17 + K)
    a = torch.randint(-63, 64, (M, K), generator=g).double()
    b = torch.randint(-63, 64, (N, K), generator=g).double()
    a[0, 0] = 63.0
    b[0, 0] = -63.0
    c = sl.gemm(sl.slice_rows(a.cuda()), sl.slice_rows(b.cuda()))
    ref = a @ b.T
    return (c.cpu() - ref).abs().max().item(), 0.0


def case_gemm(dtype, M, N, K, **kw):
    sl = _ops()
    ext = bool(kw.pop("extended", False))
    g = torch.Generator().manual_seed(M + 7 * N + 13 * K)
    a = torch.randn(M, K, generator=g, dtype=torch.float64)
    b = torch.randn(N, K, generator=g, dtype=torch.float64) * 0.01
    a, b = a.to(dtype), b.to(dtype)
    bias = torch.randn(N, generator=g, dtype=torch.float64).to(dtype) if kw.get("bias") else None
    mask = torch.randn(M, N, generator=g, dtype=torch.float64).to(dtype) if kw.get("mask") else None
    c0 = torch.randn(M, N, generator=g, dtype=torch.float64).to(dtype) if kw.get("accumulate") else None
    alpha = kw.get("alpha", 1.0)
    ref = a.double() @ b.double().T
    if bias is not None:
        ref = ref + bias.double()
    ref = alpha * ref
    if kw.get("relu"):
        ref = ref.clamp_min(0.0)
    if mask is not None:
        ref = torch.where(mask.double() > 0, ref, torch.zeros_like(ref))
    if c0 is not None:
        ref = ref + c0.double()
    out = None if c0 is None else c0.clone().cuda()
    c = sl.gemm(sl.slice_rows(a.cuda(), extended=ext), sl.slice_rows(b.cuda(), extended=ext), out=out, alpha=alpha,
                bias=None if bias is None else bias.cuda(), relu=bool(kw.get("relu")),
                mask=None if mask is None else mask.cuda(), accumulate=c0 is not None)
    scale = (a.double().abs() @ b.double().abs().T).max().item() * abs(alpha)
    return (c.double().cpu() - ref).abs().max().item() / scale, ((1e-15 if ext else 3e-13) if dtype == torch.float64 else 5e-6)


def case_cols(dtype, M, N, K):
    """dW-shaped product: both operands are column-sliced (reduction over the rows of the sources)."""
    sl = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(K, M, generator=g, dtype=torch.float64).to(dtype)
    y = torch.randn(K, N, generator=g, dtype=torch.float64).to(dtype)
    colsum = torch.zeros(M, dtype=dtype).cuda()
    c = sl.gemm(sl.slice_cols(x.cuda(), colsum=colsum, coef=0.5), sl.slice_cols(y.cuda()))
    ref = x.double().T @ y.double()
    scale = (x.double().abs().T @ y.double().abs()).max().item()
    e1 = (c.double().cpu() - ref).abs().max().item() / scale
    e2 = (colsum.double().cpu() - 0.5 * x.double().sum(0)).abs().max().item() / x.double().abs().sum(0).max().item()
    return max(e1, e2), (3e-13 if dtype == torch.float64 else 5e-6)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_slices_reconstruct(dtype):
    err, tol, ok = case_slices(dtype)
    assert ok and err <= tol, (err, tol, ok)


@pytest.mark.parametrize("shape", SHAPES[:7])
def test_integer_products_are_exact(shape):
    err, _ = case_exact_int(*shape)
    assert err == 0.0, err


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_matches_fp64(dtype, shape):
    err, tol = case_gemm(dtype, *shape)
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("shape", [(128, 64, 64), (256, 1024, 1024), (100, 70, 200)])
def test_extended_slices_reach_fp64_rounding(shape):
    """Eight digits (55 bits relative to the row maximum): the product is as good as a correctly rounded fp64 GEMM."""
    err, tol = case_gemm(torch.float64, *shape, extended=True)
    assert err <= tol, (err, tol)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_gemm_epilogue_options(dtype):
    for kw in ({"bias": True, "relu": True}, {"mask": True, "alpha": -0.3}, {"accumulate": True, "alpha": 0.2},
               {"bias": True, "mask": True, "accumulate": True}):
        err, tol = case_gemm(dtype, 200, 130, 300, **kw)
        assert err <= tol, (kw, err, tol)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_column_sliced_products(dtype):
    err, tol = case_cols(dtype, 320, 200, 256)
    assert err <= tol, (err, tol)
    err, tol = case_cols(dtype, 70, 33, 100)
    assert err <= tol, (err, tol)


if __name__ == "__main__":
    def run(name, fn, *a, **kw):
        try:
            r = fn(*a, **kw)
            torch.cuda.synchronize()
            print("%-60s %s" % (name, " ".join("%.3e" % v if isinstance(v, float) else str(v) for v in r)), flush=True)
        except Exception as e:  # keep going: the table is the point
            print("%-60s EXC %s: %s" % (name, type(e).__name__, str(e)[:300]), flush=True)

    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    dts = [d for d, n in ((torch.float64, "f64"), (torch.float32, "f32")) if which in ("all", n)]
    for dt in dts:
        run("slices %s" % dt, case_slices, dt)
    for shp in SHAPES[:7] if torch.float64 in dts else []:
        run("exact-int %s" % (shp,), case_exact_int, *shp)
    for dt in dts:
        for shp in SHAPES:
            run("gemm %s %s" % (dt, shp), case_gemm, dt, *shp)
        run("cols %s" % dt, case_cols, dt, 320, 200, 256)
        run("epi %s" % dt, case_gemm, dt, 200, 130, 300, bias=True, relu=True, mask=True, accumulate=True, alpha=0.7)
    if torch.float64 in dts:
        for shp in [(128, 64, 64), (256, 1024, 1024), (100, 70, 200)]:
            run("gemm extended %s" % (shp,), case_gemm, torch.float64, *shp, extended=True)
    # timing of the SINODE layer shape
    sl = _ops()
    for dt in dts:
        a = torch.randn(256, 3200, dtype=dt, device="cuda")
        b = torch.randn(3200, 3200, dtype=dt, device="cuda") * 0.01
        sa, sb = sl.slice_rows(a), sl.slice_rows(b)
        out = torch.empty(256, 3200, dtype=dt, device="cuda")
        for _ in range(3):
            sl.gemm(sa, sb, out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            sl.gemm(sa, sb, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 20
        ev[0].record()
        for _ in range(20):
            torch.matmul(a, b.T, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms_ref = ev[0].elapsed_time(ev[1]) / 20
        ev[0].record()
        for _ in range(20):
            sl.slice_rows(a, into=sa)
        ev[1].record()
        torch.cuda.synchronize()
        ms_sl = ev[0].elapsed_time(ev[1]) / 20
        print("time %s 256x3200x3200: sliced gemm %.3f ms (%.1f TFLOP/s equivalent), torch.matmul %.3f ms, slice_rows(256x3200) %.3f ms"
              % (dt, ms, 2 * 256 * 3200 * 3200 / ms / 1e9, ms_ref, ms_sl), flush=True)
