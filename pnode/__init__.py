"""Import-path shim: reference scripts do `from pnode import petsc_adjoint` (pnode/__init__.py:3 of the reference);
the implementation lives in pnode_b200."""
from pnode_b200 import petsc_adjoint  # noqa: F401
