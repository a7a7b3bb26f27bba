"""SINODE right-hand sides on the tensor cores: recognisers + evaluators plugged into the generic stage loop.

  * DenseMlpCallbacks   -- f(t, u) = +-Sequential(Linear, ReLU, ..., Linear)(u)  (ODEFuncEX of examples-sinode/KS/models/
                           imex.py:46-70 and examples-sinode/Burgers/Burgers.py:134-160): one C call per evaluation and
                           one per adjoint stage into csrc/dense_mlp.cu (tcgen05 sliced products; weight gradients are
                           accumulated straight into mu).  The forward keeps an activation set per stage evaluation so the
                           adjoint does not re-evaluate (the reference re-evaluates: petsc_adjoint.py:64-70).
  * CirculantCallbacks  -- sample-independent linear f_I(u) = u C^T with C circulant (ODEFuncIM, imex.py:6-44: Conv1d with
                           circular padding): stencil kernel for f_I and J^T x.
  * CirculantSolver     -- torch_linearsolve.PCShell (pnode/torch_linearsolve.py:7-35) for that operator: (shift I - J)^-1
                           from the spectrum, applied to all samples as one sliced product.
Both recognisers are structural AND numerical (the module is probed and must reproduce the candidate closed form)."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib, sliced
from .device import _stream, dtype_code
from .engine import Callbacks, ImplicitSolver
from .errors import Error


# ---- ReLU MLP ---------------------------------------------------------------------------------------------------------
def _find_relu_chain(func):
    """The nn.Sequential(Linear, ReLU, ..., Linear) that owns ALL parameters of func, or None."""
    if not isinstance(func, nn.Module):
        return None
    fparams = {id(p) for p in func.parameters()}
    for m in func.modules():
        if not isinstance(m, nn.Sequential) or len(m) == 0 or len(m) % 2 == 0:
            continue
        mods = list(m)
        if not all(isinstance(x, nn.Linear) for x in mods[0::2]) or not all(type(x) is nn.ReLU for x in mods[1::2]):
            continue
        if {id(p) for p in m.parameters()} != fparams:
            continue
        lins = mods[0::2]
        if len(lins) > _lib.DMLP_MAX_LAYERS or any(a.out_features != b.in_features for a, b in zip(lins[:-1], lins[1:])):
            continue
        return lins
    return None


def recognise_relu_mlp(func, u_meta):
    """(linears, out_scale) if func(t, u) == out_scale * chain(u) for a [batch, n] state, else None."""
    lins = _find_relu_chain(func)
    if lins is None or u_meta.dim() != 2 or not u_meta.is_cuda:
        return None
    w0 = lins[0].weight
    if u_meta.shape[1] != lins[0].in_features or lins[-1].out_features != u_meta.shape[1]:
        return None
    if w0.dtype != u_meta.dtype or w0.device != u_meta.device or u_meta.dtype not in (torch.float32, torch.float64):
        return None
    nfe = getattr(func, "nfe", None)
    try:
        with torch.no_grad():
            g = torch.Generator(device="cpu").manual_seed(1234)
            y = torch.randn(3, u_meta.shape[1], generator=g, dtype=torch.float64).to(u_meta)
            ref = y
            for i, lin in enumerate(lins):
                ref = torch.nn.functional.linear(ref, lin.weight, lin.bias)
                if i < len(lins) - 1:
                    ref = torch.relu(ref)
            o1, o2 = func(0.3, y), func(1.7, y)
        if o1.shape != ref.shape or not torch.equal(o1, o2):
            return None
        tol = 1e-12 if u_meta.dtype == torch.float64 else 1e-5
        scale = ref.abs().max().item() + 1e-300
        for s in (1.0, -1.0):
            if (o1 - s * ref).abs().max().item() <= tol * scale:
                return lins, s
        return None
    except Exception:
        return None
    finally:
        if isinstance(nfe, int):
            func.nfe = nfe


class DenseMlpCallbacks(Callbacks):
    def __init__(self, func, tensor_size, lins, out_scale):
        super().__init__(func, tensor_size)
        self.lib = _lib.load()
        self.lins = lins
        self.out_scale = float(out_scale)
        p0 = lins[0].weight
        self.dtype, self.device = p0.dtype, p0.device
        self.code = dtype_code(self.dtype)
        self.batch = int(tensor_size[0])
        offs, off = {}, 0
        for p in self.params:  # mu layout = func.parameters() filtered by requires_grad (petsc_adjoint.py:603-614)
            offs[id(p)] = off
            off += p.numel()
        d = _lib.DmlpDesc()
        d.nlayers, d.dtype, d.batch = len(lins), self.code, self.batch
        d.dims[0] = lins[0].in_features
        for l, lin in enumerate(lins):
            d.dims[l + 1] = lin.out_features
            d.mu_w_off[l] = offs.get(id(lin.weight), -1)
            d.mu_b_off[l] = offs.get(id(lin.bias), -1) if lin.bias is not None else -1
        d.out_scale = self.out_scale
        self._desc = d
        self._refresh_pointers()
        nb = [int(self.lib.pnode_dmlp_weight_bytes(C.byref(d))), int(self.lib.pnode_dmlp_act_bytes(C.byref(d))),
              int(self.lib.pnode_dmlp_work_bytes(C.byref(d)))]
        if min(nb) < 0:
            raise Error(-50, self.lib.pnode_last_error().decode())
        self._wbuf = torch.empty(nb[0], dtype=torch.uint8, device=self.device)
        self._act_bytes = nb[1]
        self._act0 = torch.empty(nb[1], dtype=torch.uint8, device=self.device)  # evaluations that are not kept
        self._work = torch.empty(nb[2], dtype=torch.uint8, device=self.device)
        self._prepared_for = None
        self._saved, self._saved_bytes, self._saved_mark = {}, 0, set()
        from .options import Options

        self._save_budget = int(float(Options().getString("pnode_dmlp_save_mb", "16384")) * (1 << 20))
        self._keep = True
        self.reused_activations = 0
        self.launches = 0

    def _refresh_pointers(self):
        """Parameters are borrowed: addresses re-read at the start of every solve."""
        for l, lin in enumerate(self.lins):
            self._desc.d_weight[l] = lin.weight.data_ptr()
            self._desc.d_bias[l] = None if lin.bias is None else lin.bias.data_ptr()

    def _prepare(self):
        ver = tuple((p.data_ptr(), p._version) for lin in self.lins for p in lin.parameters())
        if ver != self._prepared_for:
            self._refresh_pointers()
            _lib.check(self.lib.pnode_dmlp_prepare(C.byref(self._desc), self._wbuf.data_ptr(), _stream()))
            self._prepared_for = ver
            self.launches += 2 * len(self.lins)

    def begin(self, forward, keep=True, comm=None):
        super().begin(forward, keep=keep, comm=comm)
        if forward:
            self._prepared_for = None  # weights are re-sliced once per solve (an optimiser step may have changed them)
            self._saved.clear()
            self._saved_bytes = 0
            self._keep = bool(keep)
        self._prepare()

    def release(self):
        super().release()
        self._saved.clear()
        self._saved_bytes = 0

    def mark(self):
        super().mark()
        self._saved_mark = set(self._saved.keys())

    def rollback(self):
        super().rollback()
        for k in [k for k in self._saved if k not in self._saved_mark]:
            del self._saved[k]
            self._saved_bytes -= self._act_bytes

    def _count(self):
        if isinstance(getattr(self.func, "nfe", None), int):
            self.func.nfe += 1

    def _forward(self, u, act):
        out = torch.empty_like(u)
        _lib.check(self.lib.pnode_dmlp_forward(C.byref(self._desc), self._wbuf.data_ptr(), u.data_ptr(), out.data_ptr(),
                                               None if act is None else act.data_ptr(), self._work.data_ptr(), _stream()))
        self.launches += len(self.lins) * (3 if act is not None else 2)
        return out

    def f(self, t, u, keep=False):
        self.nfe += 1
        self._count()
        self._prepare()
        act = None
        if keep and self._keep and self._saved_bytes + self._act_bytes <= self._save_budget:
            act = torch.empty(self._act_bytes, dtype=torch.uint8, device=u.device)
            self._saved[u.data_ptr()] = (u, u._version, act)  # holding u keeps its address unique
            self._saved_bytes += self._act_bytes
        return self._forward(u, act)

    def _vjp(self, u, w, vu, mu, coef):
        self._prepare()
        ent = self._saved.get(u.data_ptr())
        if ent is not None and ent[1] == u._version and ent[0].numel() == u.numel():
            act = ent[2]
            self.reused_activations += 1
        else:  # no kept activation set for this point: re-evaluate like the reference does
            act = self._act0
            self._forward(u, act)
        if not w.is_contiguous():
            w = w.contiguous()
        ev = None
        if mu is not None and self.record_layer_events:
            # batch-sharded runs: mark the moment each layer's slice of mu is complete for this call, so that the caller
            # can all-reduce the slices of the LAST call layer by layer behind the rest of the sweep (layer_slices())
            if self._layer_events is None:
                self._layer_events = [torch.cuda.Event() for _ in self.lins]
                self._ev_handles = (C.c_void_p * len(self.lins))()
            for l, e in enumerate(self._layer_events):
                if not e.cuda_event:
                    e.record()  # materialise the handle
                self._ev_handles[l] = e.cuda_event
            ev = self._ev_handles
            self._events_mu_ptr = mu.data_ptr()
        _lib.check(self.lib.pnode_dmlp_vjp_ev(C.byref(self._desc), self._wbuf.data_ptr(), act.data_ptr(), w.data_ptr(),
                                              None if vu is None else vu.data_ptr(), None if mu is None else mu.data_ptr(),
                                              float(coef), self._work.data_ptr(), ev, _stream()))
        self.launches += 4 * len(self.lins)

    record_layer_events = False
    _layer_events = None
    _events_mu_ptr = None

    def layer_slices(self, mu):
        """[(event, offset, length)] last layer first: after `event`, mu[offset : offset + length] (this function's slice of
        the parameter-gradient vector handed to the LAST vjp_accumulate call) receives no further contribution.  None when
        no call recorded events into this very mu."""
        if self._layer_events is None or self._events_mu_ptr != mu.data_ptr():
            return None
        out = []
        for l in range(len(self.lins) - 1, -1, -1):
            offs = [o for o in (self._desc.mu_w_off[l], self._desc.mu_b_off[l]) if o >= 0]
            if not offs:
                continue
            lin = self.lins[l]
            lo = min(offs)
            hi = max(self._desc.mu_w_off[l] + lin.weight.numel() if self._desc.mu_w_off[l] >= 0 else 0,
                     self._desc.mu_b_off[l] + lin.bias.numel() if self._desc.mu_b_off[l] >= 0 else 0)
            out.append((self._layer_events[l], lo, hi - lo))
        return out

    def vjp(self, t, u, w, want_u=True, want_params=True):
        self.nvjp += 1
        self._count()
        vu = torch.empty_like(u) if want_u else None
        mu = torch.zeros(self.nparams, dtype=u.dtype, device=u.device) if want_params and self.nparams else None
        self._vjp(u, w, vu, mu, 1.0)
        return vu, (list(torch.split(mu, self.sizes)) if mu is not None else [])

    def vjp_accumulate(self, t, u, w, mu, coef):
        """J^T w, and mu += coef * (df/dp)^T w inside the products' epilogues."""
        self.nvjp += 1
        self._count()
        vu = torch.empty_like(u)
        if mu is not None and mu.numel() != self.nparams:
            raise Error(-51, "dense MLP: mu slice has %d entries, the function has %d parameters" % (mu.numel(), self.nparams))
        self._vjp(u, w, vu, mu if self.nparams else None, coef)
        return vu, None


# ---- circulant linear operator ------------------------------------------------------------------------------------------
def recognise_circulant(func, u_meta):
    """First column c (fp64 CPU tensor) of the circulant matrix C with func(t, u) == u C^T for every sample, or None.
    Requires: no trainable parameter, time-independent, linear, sample-independent, circulant -- each verified on probes."""
    if not isinstance(func, nn.Module) or u_meta.dim() != 2 or not u_meta.is_cuda:
        return None
    if any(p.requires_grad for p in func.parameters()) or u_meta.dtype not in (torch.float32, torch.float64):
        return None
    n = int(u_meta.shape[1])
    nfe = getattr(func, "nfe", None)
    try:
        with torch.no_grad():
            e = torch.zeros(2, n, dtype=u_meta.dtype, device=u_meta.device)
            e[0, 0] = 1.0
            e[1, n // 3] = 1.0
            r = func(0.0, e)
            if r.shape != e.shape:
                return None
            c = r[0].double()
            if not torch.equal(torch.roll(r[0], n // 3), r[1]):
                return None
            g = torch.Generator(device="cpu").manual_seed(4321)
            y = torch.randn(3, n, generator=g, dtype=torch.float64).to(u_meta)
            o1, o2 = func(0.2, y), func(1.9, y)
            if not torch.equal(o1, o2):
                return None
            idx = (torch.arange(n, device=y.device)[:, None] - torch.arange(n, device=y.device)[None, :]) % n
            ref = y.double() @ c[idx].T
            tol = 1e-12 if u_meta.dtype == torch.float64 else 1e-5
            scale = (y.double().abs() @ c[idx].abs().T).max().item() + 1e-300
            if (o1.double() - ref).abs().max().item() > tol * scale:
                return None
            if (func(0.2, torch.zeros_like(y))).abs().max().item() != 0.0:
                return None
        return c.cpu()
    except Exception:
        return None
    finally:
        if isinstance(nfe, int):
            func.nfe = nfe


class CirculantCallbacks(Callbacks):
    def __init__(self, func, tensor_size, col, dtype, device):
        super().__init__(func, tensor_size)
        self.lib = _lib.load()
        self.dtype, self.device = dtype, device
        self.code = dtype_code(dtype)
        self.n = int(tensor_size[1])
        self.set_column(col)
        self.launches = 0

    def set_column(self, col):
        self.col = col.clone()
        nz = [i for i in range(self.n) if float(col[i]) != 0.0]
        self.stencil = len(nz) <= _lib.CIRC_MAX_TAPS
        # taps in the order a direct convolution sums them: most negative offset first
        nz.sort(key=lambda i: i if i <= self.n // 2 else i - self.n)
        self._offs = (C.c_int32 * max(len(nz), 1))(*nz)
        self._coefs = (C.c_double * max(len(nz), 1))(*[float(col[i]) for i in nz])
        self._ntaps = len(nz)
        self._dense = None

    def _dense_matrix(self):
        if self._dense is None:
            n = self.n
            idx = (torch.arange(n)[:, None] - torch.arange(n)[None, :]) % n
            self._dense = self.col[idx].to(device=self.device, dtype=self.dtype)
        return self._dense

    def _apply(self, x, transpose):
        if not self.stencil:  # many taps: a dense product (still sample-independent)
            M = self._dense_matrix()
            return (x.view(-1, self.n) @ (M if transpose else M.T)).reshape(-1)
        out = torch.empty_like(x)
        _lib.check(self.lib.pnode_circulant_apply(x.data_ptr(), out.data_ptr(), x.numel() // self.n, self.n, self._offs,
                                                  self._coefs, self._ntaps, int(transpose), self.code, _stream()))
        self.launches += 1
        return out

    def _count(self):
        if isinstance(getattr(self.func, "nfe", None), int):
            self.func.nfe += 1

    def f(self, t, u, keep=False):
        self.nfe += 1
        self._count()
        return self._apply(u if u.is_contiguous() else u.contiguous(), False)

    def jvp(self, t, u, v):
        self.nfe += 1
        return self._apply(v.contiguous(), False)

    def vjp(self, t, u, w, want_u=True, want_params=True):
        self.nvjp += 1
        self._count()
        return (self._apply(w if w.is_contiguous() else w.contiguous(), True) if want_u else None), []


class CirculantSolver(ImplicitSolver):
    """linear_solver='torch' for a recognised circulant operator: inverse of shift*I - J from the spectrum (per shift),
    applied to every sample as one tensor-core product.  f_I is linear, so the one Newton step of -snes_type ksponly (and the
    converged Newton iteration without it) is Y = shift (shift I - J)^-1 Z exactly: no residual evaluation, no cancellation."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.lib = _lib.load()
        self._cwork = None

    _probed_sig = None

    def reset(self):
        """Once per odeint (petsc_adjoint.py:792-799).  The operator is re-read (one evaluation on a unit impulse) only when
        a parameter or buffer of the function has been modified since the last look; an unchanged operator has the same
        inverses bit for bit, so they are kept (with or without fixed_jacobian)."""
        if not hasattr(self, "_circ"):
            self._circ = {}
            return  # constructor: nothing to probe yet
        cb = self.cb
        sig = self._signature()
        if sig is not None and sig == self._probed_sig:
            return
        with torch.no_grad():
            e = torch.zeros(1, cb.n, dtype=cb.dtype, device=cb.device)
            e[0, 0] = 1.0
            nfe = getattr(cb.func, "nfe", None)
            col = cb.func(0.0, e)[0].double().cpu()
            if isinstance(nfe, int):
                cb.func.nfe = nfe
        if not torch.equal(col, cb.col):
            cb.set_column(col)
            self._circ.clear()
        self._probed_sig = sig

    def _inverse(self, shift):
        key = float(shift)
        ent = self._circ.get(key)
        if ent is None:
            cb = self.cb
            n = cb.n
            if self._cwork is None:
                self._cwork = torch.empty(int(self.lib.pnode_circulant_work_bytes(n)), dtype=torch.uint8, device=cb.device)
            col = cb.col.to(cb.device)
            dense = torch.empty(n, n, dtype=cb.dtype, device=cb.device)
            _lib.check(self.lib.pnode_circulant_inverse(col.data_ptr(), n, key, dense.data_ptr(), cb.code,
                                                        self._cwork.data_ptr(), _stream()))
            # B operands of X = R A^-T and X = R A^-1.  The right-hand sides of the stiff stages are rough (K^I_0 = J u_n is
            # ~1e7 |u_n| for KS at N = 1024) and the inverse annihilates almost all of them: the product cancels by ~7 orders
            # of magnitude, so these operands carry the extra digit (55 bits relative to the row maximum)
            ent = (sliced.slice_rows(dense, extended=True), sliced.slice_cols(dense, extended=True))
            self._circ[key] = ent
        return ent

    def _apply(self, t, y, shift, rhs, transpose, alpha=1.0):
        n = self.cb.n
        R = rhs.view(-1, n)
        inv = self._inverse(shift)[1 if transpose else 0]
        out = torch.empty_like(R)
        sliced.gemm(sliced.slice_rows(R, extended=True), inv, out=out, alpha=alpha)
        return out.reshape(-1)

    def solve(self, t, Z, shift, guess, aff=None):
        if self.mass is not None or aff is not None:
            return super().solve(t, Z, shift, guess, aff=aff)
        return self._apply(t, None, shift, Z, False, alpha=float(shift))

    def solve_transpose(self, t, y, shift, rhs):
        if self.mass is not None:
            return super().solve_transpose(t, y, shift, rhs)
        return self._apply(t, y, shift, rhs, True)
