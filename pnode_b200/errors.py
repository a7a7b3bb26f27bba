"""Error type raised for any failure inside the native engine.

The reference surfaces PETSc failures as `petsc4py.PETSc.Error`; the petsc4py shim re-exports this
class under that name so `except PETSc.Error` in user scripts keeps working.
"""


class Error(RuntimeError):
    def __init__(self, ierr=0, msg=""):
        self.ierr = ierr
        super().__init__("pnode_b200 error %d: %s" % (ierr, msg) if msg else "pnode_b200 error %d" % ierr)
