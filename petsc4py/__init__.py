"""Minimal `petsc4py` stand-in for scripts written against the reference.

Those scripts only do `petsc4py.init(sys.argv)` and `from petsc4py import PETSc` (examples-pnode/ode_demo_petsc.py:63-67,
tests/test_pnode.py:11,32,128-130) -- the PETSc solver objects themselves were created inside pnode, which is now
pnode_b200.  `init` feeds the `-ts_*` style options into the engine's options database.
"""
import sys as _sys

from pnode_b200.options import Options as _Options

__version__ = "0.0.pnode_b200"


def init(args=None, arch=None, comm=None):
    if args is None:
        args = _sys.argv
    _Options.insert_args(list(args)[1:] if args and not str(args[0]).startswith("-") else list(args or []))


def get_config():
    return {"PETSC_ARCH": "pnode_b200", "PETSC_DIR": ""}


from . import PETSc  # noqa: E402,F401
