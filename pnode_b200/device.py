"""Thin typed wrappers that launch the C-ABI kernels on torch-owned CUDA memory.

PyTorch is plumbing here: it owns the allocations and the stream; every arithmetic op below is a kernel of
csrc/*.cu reached through include/pnode_b200.h.  Tensors must be CUDA tensors -- there is no CPU path.
"""
import ctypes as C

import torch

from . import _lib
from .errors import Error


def dtype_code(dt):
    if dt == torch.float32:
        return _lib.F32
    if dt == torch.float64:
        return _lib.F64
    raise Error(-10, "unsupported state dtype %s (the engine computes in float32 or float64, like a PETSc build)" % dt)


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw accessor: torch.cuda.current_stream() builds a
    Stream object per call, ~25 us -- a tenth of a whole fwd+adjoint pass of the small-batch configurations)."""
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:  # private accessors moved: the documented route
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise Error(-11, "%s must live on a CUDA device (got %s); pnode_b200 has no CPU fallback" % (what, t.device))


class DeviceOps:
    """Vector kernels of the generic path (one launch per RK stage / completion / adjoint stage)."""

    def __init__(self, device, dtype):
        self.lib = _lib.load()
        self.device = device
        self.dtype = dtype
        self.code = dtype_code(dtype)
        self._wrms_work = None
        self._sumsq = None
        self._mdot_work = None
        self._mdot_out = None
        self.launches = 0

    # out = base_coef*base + sum_j coefs[j]*vecs[j]
    def lincomb(self, out, base, base_coef, vecs, coefs):
        _require_cuda(out, "state vector")
        n = out.numel()
        nt = len(vecs)
        if nt > _lib.MAX_TERMS:  # split: fold the first 16 terms, then keep adding
            self.lincomb(out, base, base_coef, vecs[:_lib.MAX_TERMS], coefs[:_lib.MAX_TERMS])
            return self.lincomb(out, out, 1.0, vecs[_lib.MAX_TERMS:], coefs[_lib.MAX_TERMS:])
        vp = (C.c_void_p * max(nt, 1))(*[v.data_ptr() for v in vecs])
        cf = (C.c_double * max(nt, 1))(*[float(c) for c in coefs])
        _lib.check(self.lib.pnode_lincomb(out.data_ptr(), None if base is None else base.data_ptr(), float(base_coef),
                                          vp, cf, nt, n, self.code, _stream()))
        self.launches += 1
        return out

    # u_new = u + sum bw_j k_j ; optional embedded error -> device scalar sum of squares of the weighted error
    def complete(self, unew, u, ks, bw, ew=None, atol=0.0, rtol=0.0):
        _require_cuda(unew, "state vector")
        nt = len(ks)
        vp = (C.c_void_p * max(nt, 1))(*[k.data_ptr() for k in ks])
        b = (C.c_double * max(nt, 1))(*[float(c) for c in bw])
        sumsq = None
        if ew is not None:
            if self._wrms_work is None:
                self._wrms_work = torch.zeros(int(self.lib.pnode_wrms_work_bytes()), dtype=torch.uint8,
                                              device=self.device)
                self._sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
            e = (C.c_double * max(nt, 1))(*[float(c) for c in ew])
            sumsq = self._sumsq
            _lib.check(self.lib.pnode_rk_complete_wrms(unew.data_ptr(), u.data_ptr(), vp, b, e, nt, unew.numel(),
                                                       float(atol), float(rtol), sumsq.data_ptr(),
                                                       self._wrms_work.data_ptr(), self.code, _stream()))
        else:
            _lib.check(self.lib.pnode_rk_complete_wrms(unew.data_ptr(), u.data_ptr(), vp, b, None, nt, unew.numel(),
                                                       0.0, 0.0, None, None, self.code, _stream()))
        self.launches += 1
        return sumsq

    # mu[off_k : off_k+size_k] += coef * grads[k]   (None => zeros)
    def multi_axpy(self, mu, grads, sizes, coef):
        ns = len(grads)
        if ns == 0:
            return
        ptrs, keep = [], []  # `keep` holds contiguous copies until the launch below has been issued
        for g in grads:
            if g is None:
                ptrs.append(None)
            else:
                if g.dtype != mu.dtype or g.device != mu.device:
                    raise Error(-18, "parameter gradient is %s on %s, mu is %s on %s" % (g.dtype, g.device, mu.dtype, mu.device))
                if not g.is_contiguous():
                    g = g.contiguous()
                keep.append(g)
                ptrs.append(g.data_ptr())
        vp = (C.c_void_p * ns)(*ptrs)
        sz = (C.c_int64 * ns)(*[int(s) for s in sizes])
        _lib.check(self.lib.pnode_multi_axpy(mu.data_ptr(), vp, sz, ns, float(coef), self.code, _stream()))
        del keep
        self.launches += 1

    # [<v_j, w> for j] + [<w, w>] as a HOST list of floats (one device->host read): GMRES Gram-Schmidt coefficients
    # column-wise variants (block Krylov solver): vectors are [nseg][n / nseg]; results / coefficients stay on the device
    def mdot_seg(self, vecs, w, nseg):
        """Returns a device tensor [len(vecs) + 1, nseg]: the per-segment dots with w, and ||w||^2 per segment last."""
        n = w.numel()
        out = torch.empty(len(vecs) + 1, nseg, dtype=torch.float64, device=w.device)
        k = 0
        while True:
            chunk = vecs[k:k + 16]
            vp = (C.c_void_p * max(len(chunk), 1))(*[v.data_ptr() for v in chunk])
            tmp = out[k:] if k + len(chunk) == len(vecs) else torch.empty(len(chunk) + 1, nseg, dtype=torch.float64,
                                                                           device=w.device)
            _lib.check(self.lib.pnode_mdot_seg(tmp.data_ptr(), vp, len(chunk), w.data_ptr(), nseg, n // nseg, self.code,
                                               _stream()))
            self.launches += 1
            if tmp.data_ptr() != out[k:].data_ptr():
                out[k:k + len(chunk)].copy_(tmp[:len(chunk)])
            k += 16
            if k >= len(vecs):
                return out

    def lincomb_seg(self, out, base, base_coef, vecs, coef, mode, nseg):
        """out[s,:] = base_coef base[s,:] + sum_j c_j[s] vecs[j][s,:], c from the device tensor coef [>= len(vecs), nseg]."""
        n = out.numel()
        k = 0
        while True:
            chunk = vecs[k:k + 16]
            vp = (C.c_void_p * max(len(chunk), 1))(*[v.data_ptr() for v in chunk])
            b, bc = (base, base_coef) if k == 0 else (out, 1.0)
            _lib.check(self.lib.pnode_lincomb_seg(out.data_ptr(), None if b is None else b.data_ptr(), float(bc), vp,
                                                  None if not chunk else coef[k:].data_ptr(), len(chunk), int(mode), nseg,
                                                  n // nseg, self.code, _stream()))
            self.launches += 1
            k += 16
            if k >= len(vecs):
                return out

    def mdot(self, vecs, w):
        if self._mdot_work is None:
            self._mdot_work = torch.zeros(int(self.lib.pnode_mdot_work_bytes()), dtype=torch.uint8, device=self.device)
            self._mdot_out = torch.zeros(64, dtype=torch.float64, device=self.device)
        out, k = self._mdot_out, 0
        n = w.numel()
        if not vecs:
            _lib.check(self.lib.pnode_mdot(out.data_ptr(), (C.c_void_p * 1)(), 0, w.data_ptr(), n,
                                           self._mdot_work.data_ptr(), self.code, _stream()))
            self.launches += 1
            return [], float(out[0].item())
        vals = []
        while k < len(vecs):
            chunk = vecs[k:k + 16]
            vp = (C.c_void_p * len(chunk))(*[v.data_ptr() for v in chunk])
            _lib.check(self.lib.pnode_mdot(out.data_ptr(), vp, len(chunk), w.data_ptr(), n, self._mdot_work.data_ptr(),
                                           self.code, _stream()))
            self.launches += 1
            host = out[:len(chunk) + 1].tolist()
            vals += host[:-1]
            ww = host[-1]
            k += 16
        return vals, ww
