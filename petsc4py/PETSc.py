"""`petsc4py.PETSc` names that reference scripts touch: ScalarType, Options, Error, COMM_SELF / COMM_WORLD."""
import numpy as _np

from pnode_b200.errors import Error  # noqa: F401
from pnode_b200.options import Options  # noqa: F401

# The engine integrates in the dtype of the tensor handed to setupTS (float32 or float64), so it plays the role of both
# a double- and a single-precision PETSc build; report the double build the reference's test asserts on
# (tests/test_pnode.py:127-130).
ScalarType = _np.float64
RealType = _np.float64
IntType = _np.int32
COMM_SELF = "COMM_SELF"
COMM_WORLD = "COMM_WORLD"
DECIDE = -1
