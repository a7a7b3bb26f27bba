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
    assert lib.pnode_mlp_rk_supported(3, 50, 1, _lib.F64, 4) == 0


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
