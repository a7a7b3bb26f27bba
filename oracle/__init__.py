"""ORACLE -- test infrastructure only.

CPU restatement of the reference's algorithm for the hot path (pnode's Python callbacks + the PETSc TS/TSAdjoint
arithmetic they sit on).  Nothing under pnode_b200/, pnode/ or petsc4py/ may import this package: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the checker / the timed
CPU baseline.

Parity status: forward pinned on the reference's three ROBER known-answer values (tests/test_pnode.py:151-152,179-180,
200-201); gradients, adaptive step sequences and fp32 are "parity unpinned" by the reference (it asserts nothing about
them, and PETSc cannot be built or imported here) and are instead cross-checked against torch.autograd through the
unrolled scheme.
"""
from .odepetsc import OracleODEPetsc  # noqa: F401
from .petsc_ts import OracleTS, parse_petsc_options, wrms_norm  # noqa: F401
