"""The C-ABI library loads without a GPU and exports every function include/pnode_b200.h declares."""
import ctypes
import os
import re

from pnode_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pnode_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnode_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    decl = _declared()
    assert decl, "no declarations found"
    assert decl == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name


def test_loads_and_reports_version():
    lib = _lib.load()
    assert lib.pnode_abi_version() == 1
    assert lib.pnode_wrms_work_bytes() > 0
    assert lib.pnode_mlp_rk_supported(2, 50, 1, _lib.F64, 4) == 1
    assert lib.pnode_mlp_rk_supported(2, 50, 1, _lib.F32, 7) == 1
    assert lib.pnode_mlp_rk_supported(3, 50, 1, _lib.F64, 4) == 1 and lib.pnode_mlp_rk_supported(2, 100, 0, _lib.F32, 7) == 1
    assert lib.pnode_mlp_rk_supported(5, 50, 1, _lib.F64, 4) == 0 and lib.pnode_mlp_rk_supported(2, 64, 1, _lib.F64, 4) == 0


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Step) == 24
    assert ctypes.sizeof(_lib.RKTableau) == 8 + 8 * (49 + 7 + 7 + 7) + 8
    assert ctypes.sizeof(_lib.CnfDesc) == 16 + 11 * 8
    assert ctypes.sizeof(_lib.MlpDesc) == 16 + 4 * 8


def test_product_refuses_cpu_tensors():
    import pytest
    import torch
    from pnode_b200 import Error, petsc_adjoint

    ode = petsc_adjoint.ODEPetsc()
    with pytest.raises(Error):
        ode.setupTS(torch.zeros(4, 2), torch.nn.Linear(2, 2), method="rk4")


def test_product_does_not_import_oracle():
    import subprocess
    import sys

    code = "import sys; sys.path.insert(0, %r); import pnode, petsc4py, pnode_b200; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pnode_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_convblock_plan_runs_without_a_gpu():
    """Host half of csrc/conv_block.cu (shape validation, buffer sizing, parameter layout) -- no kernel is launched."""
    import ctypes as C

    import torch

    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _workloads import OdeConvBlock

    lib = _lib.load()
    for (N, Cc, H, W), dtype in (((256, 32, 32, 32), torch.float32), ((4, 16, 8, 8), torch.float64), ((256, 256, 4, 4), torch.float32)):
        func = OdeConvBlock(Cc, dtype=dtype)
        d = _lib.ConvBlockDesc()
        d.nlayers, d.dtype, d.N, d.H, d.W = 5, (0 if dtype == torch.float32 else 1), N, H, W
        for k in range(5):
            conv, bn = getattr(func, "conv%d" % (k + 1)), getattr(func, "bn%d" % (k + 1))
            l = d.layer[k]
            l.cin, l.cout = conv.in_channels, conv.out_channels
            l.kh, l.kw = conv.kernel_size
            l.ph, l.pw = conv.padding
            l.d_weight = l.d_bias = l.d_gamma = l.d_beta = 16  # non-null placeholders: the plan never dereferences them
            l.eps, l.momentum = bn.eps, bn.momentum
        assert lib.pnode_convblock_param_count(C.byref(d)) == sum(p.numel() for p in func.parameters())
        esz = 4 if dtype == torch.float32 else 8
        zbytes = sum(getattr(func, "conv%d" % (k + 1)).out_channels for k in range(5)) * N * H * W * esz
        act, work = lib.pnode_convblock_act_bytes(C.byref(d)), lib.pnode_convblock_work_bytes(C.byref(d))
        assert zbytes <= act <= zbytes + (1 << 20) and work >= zbytes  # z_1..z_L + statistics; g_k of every layer + partials
    d.W = 12
    assert lib.pnode_convblock_act_bytes(C.byref(d)) == -1 and b"power of two" in lib.pnode_last_error()
    d.W = 4
    d.layer[2].cin = 6
    assert lib.pnode_convblock_act_bytes(C.byref(d)) == -1 and b"multiples of 4" in lib.pnode_last_error()
    assert C.sizeof(_lib.ConvLayer) == 24 + 7 * 8 + 16 and C.sizeof(_lib.ConvBlockDesc) == 24 + 8 * C.sizeof(_lib.ConvLayer) + 32


def test_control_block_layout_matches_a_c_compiler(tmp_path):
    """pnode_cnf_ctl is written by the host and read / updated by the attempt kernel's step controller: the ctypes mirror
    must agree with what a C compiler makes of include/pnode_b200.h (size and the offsets the host relies on)."""
    import subprocess

    src = tmp_path / "ctl.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pnode_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pnode_cnf_ctl), offsetof(pnode_cnf_ctl, span), '
                   'offsetof(pnode_cnf_ctl, nspan), offsetof(pnode_cnf_ctl, max_steps), offsetof(pnode_cnf_ctl, sumsq), '
                   'offsetof(pnode_cnf_ctl, log_t), offsetof(pnode_cnf_ctl, log_accepted)); return 0; }\n')
    exe = tmp_path / "ctl"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    c = _lib.CnfCtl
    assert got == [ctypes.sizeof(c), c.span.offset, c.nspan.offset, c.max_steps.offset, c.sumsq.offset, c.log_t.offset,
                   c.log_accepted.offset]
