"""Fused-kernel paths: recognisers that map a user `nn.Module` onto a hand-written sweep kernel, and the host glue
that launches it.  A module that is not recognised (or `-pnode_fused 0`) runs on the generic path of engine.py -- both
execute on the GPU; neither is a fallback for a missing library.

Currently fused: tiny-state MLPs  f(t,y) = Linear(H,d)(tanh(Linear(d,H)(phi(y)))), phi = y**3 | y  -- the spiral model of
examples-pnode/ode_demo_petsc.py:207-230 -- under any fixed-step explicit RK scheme (csrc/mlp_rk.cu).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .controller import fixed_schedule
from .device import _stream, dtype_code

_STEP_DTYPE = np.dtype([("t", "<f8"), ("h", "<f8"), ("out_slot", "<i4"), ("in_slot", "<i4")])
assert _STEP_DTYPE.itemsize == C.sizeof(_lib.Step)


class MlpSpec:
    def __init__(self, lin1, lin2, phi, dim, hidden):
        self.lin1, self.lin2, self.phi, self.dim, self.hidden = lin1, lin2, phi, dim, hidden

    def params(self):
        return [self.lin1.weight, self.lin1.bias, self.lin2.weight, self.lin2.bias]


def recognise_mlp(func, u_meta):
    """Return an MlpSpec when `func(t, y)` is provably  lin2(tanh(lin1(phi(y))))  acting on the last axis of y:
    structural match of the module tree AND a numerical probe of the forward on random inputs at two times."""
    if not isinstance(func, nn.Module) or not u_meta.is_cuda or u_meta.dim() < 1:
        return None
    mods = [m for m in func.modules() if m is not func and not isinstance(m, (nn.Sequential, nn.ModuleList))]
    lins = [m for m in mods if isinstance(m, nn.Linear)]
    tanhs = [m for m in mods if isinstance(m, nn.Tanh)]
    if len(lins) != 2 or len(tanhs) != 1 or len(mods) != 3:
        return None
    lin1, lin2 = lins
    d = u_meta.shape[-1]
    if lin1.bias is None or lin2.bias is None:
        return None
    if lin1.in_features != d or lin2.out_features != d or lin1.out_features != lin2.in_features:
        return None
    plist = [p for p in func.parameters() if p.requires_grad]
    want = [lin1.weight, lin1.bias, lin2.weight, lin2.bias]
    if len(plist) != 4 or any(a is not b for a, b in zip(plist, want)):
        return None
    if any(p.dtype != u_meta.dtype or p.device != u_meta.device or not p.is_contiguous() for p in want):
        return None
    if len(list(func.buffers())) != 0:
        return None
    # numerical probe (a handful of tiny launches, once per func identity)
    with torch.no_grad():
        g = torch.Generator(device="cpu").manual_seed(1234)
        shape = (3,) + tuple(u_meta.shape[1:]) if u_meta.dim() > 1 else tuple(u_meta.shape)
        x = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).to(device=u_meta.device, dtype=u_meta.dtype)
        try:
            y0 = func(0.0, x)
            y1 = func(0.73, x)
        except Exception:
            return None
        if y0.shape != x.shape or not torch.equal(y0, y1):
            return None
        tol = 1e-5 if u_meta.dtype == torch.float32 else 1e-12
        for phi, fn in ((1, lambda v: v ** 3), (0, lambda v: v)):
            ref = lin2(torch.tanh(lin1(fn(x))))
            if torch.allclose(ref, y0, rtol=tol, atol=tol):
                return MlpSpec(lin1, lin2, phi, d, lin1.out_features)
    return None


def _tableau_struct(scheme):
    tab = _lib.RKTableau()
    tab.s = scheme.s
    tab.fsal = 1 if scheme.fsal else 0
    for i in range(scheme.s):
        tab.b[i] = scheme.b[i]
        tab.c[i] = scheme.c[i]
        for j in range(scheme.s):
            tab.a[i][j] = scheme.A[i][j]
    return tab


class FusedMlpRK:
    """Whole-sweep launches for a recognised MLP under a fixed-step explicit RK scheme."""

    def __init__(self, spec, scheme, dtype, device):
        self.lib = _lib.load()
        self.spec = spec
        self.scheme = scheme
        self.dtype = dtype
        self.device = device
        self.code = dtype_code(dtype)
        self.tab = _tableau_struct(scheme)
        self._sched_cache = {}
        self._work = None
        self.launches = 0

    @staticmethod
    def supported(spec, scheme, dtype):
        return bool(_lib.load().pnode_mlp_rk_supported(spec.dim, spec.hidden, spec.phi, dtype_code(dtype), scheme.s))

    def _desc(self):
        sp = self.spec
        d = _lib.MlpDesc()
        d.dim, d.hidden, d.phi, d.dtype = sp.dim, sp.hidden, sp.phi, self.code
        # live parameter storage: optimiser updates are seen by the next launch without a new setupTS
        d.d_w1, d.d_b1 = sp.lin1.weight.data_ptr(), sp.lin1.bias.data_ptr()
        d.d_w2, d.d_b2 = sp.lin2.weight.data_ptr(), sp.lin2.bias.data_ptr()
        return d

    def _schedule(self, times, step_size):
        key = (tuple(times), tuple(step_size) if isinstance(step_size, list) else float(step_size))
        hit = self._sched_cache.get(key)
        if hit is not None:
            return hit
        loop, steps = fixed_schedule(times, step_size, self.dtype == torch.float64)
        loop.check_complete()
        arr = np.zeros(len(steps), dtype=_STEP_DTYPE)
        single = len(times) == 1
        for n, (t, h, slot) in enumerate(steps):
            arr[n]["t"], arr[n]["h"] = t, h
            if single:
                arr[n]["out_slot"] = 0 if n == len(steps) - 1 else -1
                arr[n]["in_slot"] = -1
            else:
                arr[n]["out_slot"] = slot
                arr[n]["in_slot"] = 0 if n == 0 else steps[n - 1][2]
        dev = torch.from_numpy(arr.view(np.uint8)).to(self.device)
        entry = (dev, len(steps), loop)
        if len(self._sched_cache) > 64:
            self._sched_cache.clear()
        self._sched_cache[key] = entry
        return entry

    def forward(self, u0, times, step_size, save):
        """u0: flat [ntraj*dim].  Returns (sol [T, ntraj*dim], ckpt or None, schedule entry)."""
        sp = self.spec
        ntraj = u0.numel() // sp.dim
        sched, nsteps, loop = self._schedule(times, step_size)
        T = len(times)
        sol = torch.empty((T, u0.numel()), dtype=self.dtype, device=self.device)
        if T > 1:
            sol[0].copy_(u0)
        ckpt = None
        if save:
            ckpt = torch.empty((max(nsteps, 1), self.scheme.s, sp.dim, ntraj), dtype=self.dtype, device=self.device)
        if nsteps == 0:
            sol[-1].copy_(u0)
        desc = self._desc()
        _lib.check(self.lib.pnode_mlp_rk_forward(C.byref(desc), C.byref(self.tab), u0.data_ptr(), ntraj,
                                                 sched.data_ptr(), nsteps, sol.data_ptr(),
                                                 None if ckpt is None else ckpt.data_ptr(), _stream()))
        self.launches += 1
        return sol, ckpt, (sched, nsteps, loop)

    def adjoint(self, gout, ckpt, sched_entry, ntraj):
        """gout: contiguous [T, ntraj*dim].  Returns (lambda [ntraj*dim], mu [np])."""
        sp = self.spec
        sched, nsteps, _ = sched_entry
        T = gout.shape[0]
        lam = torch.empty(ntraj * sp.dim, dtype=self.dtype, device=self.device)
        npar = 2 * sp.hidden * sp.dim + sp.hidden + sp.dim
        mu = torch.empty(npar, dtype=self.dtype, device=self.device)
        desc = self._desc()
        if self._work is None:
            nbytes = int(self.lib.pnode_mlp_rk_adjoint_work_bytes(C.byref(desc)))
            self._work = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        if nsteps == 0:
            lam.copy_(gout[-1])
            mu.zero_()
            return lam, mu
        _lib.check(self.lib.pnode_mlp_rk_adjoint(C.byref(desc), C.byref(self.tab), ntraj, sched.data_ptr(), nsteps,
                                                 T - 1, gout.data_ptr(), ckpt.data_ptr(), lam.data_ptr(),
                                                 mu.data_ptr(), self._work.data_ptr(), _stream()))
        self.launches += 1
        return lam, mu
