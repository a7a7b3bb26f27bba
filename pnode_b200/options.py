"""PETSc-style options database.

The reference's scripts leave every `-key [value]` they do not understand in `sys.argv`, hand it to
`petsc4py.init(sys.argv)` (examples-pnode/ode_demo_petsc.py:63-67, tests/test_pnode.py:25-32) and PETSc consumes it at
`ts.setFromOptions()` (pnode/petsc_adjoint.py:775).  The engine honours exactly the options those scripts use
(SURVEY.md section 5 "Config / flags"); the database is process-global like PETSc's.
"""

KNOWN = (
    "ts_type", "ts_rk_type", "ts_arkimex_type", "ts_adapt_type", "ts_atol", "ts_rtol", "ts_max_reject", "ts_dt",
    "ts_trajectory_type", "ts_trajectory_solution_only", "ts_trajectory_max_cps_ram", "ts_monitor", "snes_type",
    "snes_rtol", "snes_max_it", "ksp_rtol", "ksp_max_it", "pnode_inner_ksp_hpddm_type", "pnode_fused",
)


def _is_number(s):
    try:
        float(s)
        return True
    except (TypeError, ValueError):
        return False


class Options:
    """Mirror of `petsc4py.PETSc.Options` for the calls user scripts make (getAll/hasName/getString/setValue ...)."""

    _db = {}

    def __init__(self, prefix=None):
        self.prefix = prefix or ""

    # -- population --------------------------------------------------------------------------------------------------
    @classmethod
    def insert_args(cls, argv):
        argv = [str(a) for a in (argv or [])]
        i = 0
        while i < len(argv):
            a = argv[i]
            if a.startswith("-") and not a.startswith("--") and len(a) > 1 and not _is_number(a):
                key = a[1:]
                nxt = argv[i + 1] if i + 1 < len(argv) else None
                if nxt is not None and (not nxt.startswith("-") or _is_number(nxt)):
                    cls._db[key] = nxt
                    i += 2
                    continue
                cls._db[key] = None
            i += 1

    @classmethod
    def clear_all(cls):
        cls._db.clear()

    # -- queries ----------------------------------------------------------------------------------------------------
    def getAll(self):
        return dict(self._db)

    def hasName(self, name):
        return (self.prefix + name.lstrip("-")) in self._db

    def getString(self, name, default=None):
        return self._db.get(self.prefix + name.lstrip("-"), default)

    def getReal(self, name, default=None):
        v = self.getString(name, None)
        return default if v is None else float(v)

    def getInt(self, name, default=None):
        v = self.getString(name, None)
        return default if v is None else int(v)

    def getBool(self, name, default=False):
        key = self.prefix + name.lstrip("-")
        if key not in self._db:
            return default
        v = self._db[key]
        return True if v is None else str(v).lower() in ("1", "true", "yes", "on")

    def setValue(self, name, value):
        self._db[self.prefix + name.lstrip("-")] = None if value is None else str(value)

    def delValue(self, name):
        self._db.pop(self.prefix + name.lstrip("-"), None)

    def __contains__(self, name):
        return self.hasName(name)

    def __getitem__(self, name):
        return self._db[self.prefix + name.lstrip("-")]

    def __setitem__(self, name, value):
        self.setValue(name, value)
