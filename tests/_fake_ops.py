"""TEST DOUBLE (lives in tests/, never shipped): a torch-CPU stand-in for pnode_b200.device.DeviceOps so that the HOST
logic of the engine (stage loops, step controller, adjoint recurrences, option handling) can be exercised by the
`-m "not gpu"` suite.  The product only ever constructs the CUDA DeviceOps."""
import torch


class FakeOps:
    def __init__(self, device=None, dtype=None):
        self.launches = 0

    def lincomb(self, out, base, base_coef, vecs, coefs):
        acc = torch.zeros_like(out) if base is None else base_coef * base
        for v, c in zip(vecs, coefs):
            acc = acc + c * v.reshape(-1)
        out.copy_(acc)
        self.launches += 1
        return out

    def complete(self, unew, u, ks, bw, ew=None, atol=0.0, rtol=0.0):
        acc = u.clone()
        for k, b in zip(ks, bw):
            acc = acc + b * k
        unew.copy_(acc)
        self.launches += 1
        if ew is None:
            return None
        err = torch.zeros_like(u)
        for k, e in zip(ks, ew):
            err = err + e * k
        x = (acc + err).double()
        a = acc.double()
        tol = atol + rtol * torch.maximum(a.abs(), x.abs())
        return (((a - x) / tol) ** 2).sum().reshape(1)

    def multi_axpy(self, mu, grads, sizes, coef):
        off = 0
        for g, n in zip(grads, sizes):
            if g is not None:
                mu[off:off + n] += coef * g.reshape(-1)
            off += n
        self.launches += 1


def _mdot(self, vecs, w):
    self.launches += 1
    wd = w.reshape(-1).double()
    return [float((v.reshape(-1).double() * wd).sum()) for v in vecs], float((wd * wd).sum())


FakeOps.mdot = _mdot


def _mdot_seg(self, vecs, w, nseg):
    self.launches += 1
    wd = w.reshape(nseg, -1).double()
    rows = [(v.reshape(nseg, -1).double() * wd).sum(1) for v in vecs] + [(wd * wd).sum(1)]
    return torch.stack(rows)


def _lincomb_seg(self, out, base, base_coef, vecs, coef, mode, nseg):
    self.launches += 1
    acc = torch.zeros(nseg, out.numel() // nseg, dtype=torch.float64) if base is None else \
        base_coef * base.reshape(nseg, -1).double()
    for j, v in enumerate(vecs):
        c = coef[j].double()
        c = c if mode == 0 else -c if mode == 1 else torch.where(c > 0, 1.0 / c.clamp_min(1e-300).sqrt(), torch.zeros_like(c))
        acc = acc + c[:, None] * v.reshape(nseg, -1).double()
    out.copy_(acc.reshape(out.shape).to(out.dtype))
    return out


FakeOps.mdot_seg = _mdot_seg
FakeOps.lincomb_seg = _lincomb_seg


def patch_cpu(monkeypatch):
    """Route ODEPetsc onto FakeOps and lift the CUDA-only gate -- for host-logic tests only."""
    import pnode_b200.petsc_adjoint as pa

    monkeypatch.setattr(pa, "DeviceOps", FakeOps)
    monkeypatch.setattr(pa, "_check_device", lambda t, what: None)
    return pa
