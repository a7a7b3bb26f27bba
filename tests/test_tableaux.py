"""Tableau self-tests (the only guard without PETSc) + product tables == oracle tables."""
from fractions import Fraction

import pytest

from oracle import tableaux as otab
from pnode_b200 import tableaux as ptab


@pytest.mark.parametrize("name", sorted(ptab.RK))
def test_rk_tables(name):
    sc, oc = ptab.RK[name], otab.RK[name]
    assert sum(sc.b_exact) == 1
    if sc.bembed_exact is not None:
        assert sum(sc.bembed_exact) == 1
    A, b, be, c = oc.floats()
    assert sc.A == A and sc.b == b and sc.c == c and sc.bembed == be and sc.fsal == oc.fsal and sc.order == oc.order
    # order conditions up to 3 in exact arithmetic
    cs = [sum(r, Fraction(0)) for r in sc.A_exact]
    if sc.order >= 2:
        assert sum(bi * ci for bi, ci in zip(sc.b_exact, cs)) == Fraction(1, 2)
    if sc.order >= 3:
        assert sum(bi * ci * ci for bi, ci in zip(sc.b_exact, cs)) == Fraction(1, 3)
        assert sum(sc.b_exact[i] * sc.A_exact[i][j] * cs[j] for i in range(sc.s) for j in range(sc.s)) == Fraction(1, 6)
    if sc.fsal:
        assert sc.A_exact[-1] == sc.b_exact and sc.b_exact[-1] == 0


@pytest.mark.parametrize("name", sorted(ptab.ARK))
def test_ark_tables(name):
    sc, oc = ptab.ARK[name], otab.ARK[name]
    At, A, bt, b, be, ct, c = oc.floats()
    assert sc.At == At and sc.A == A and sc.b == b and sc.bt == bt and sc.bembed == be
    assert sc.c == pytest.approx(c, abs=0) and sc.ct == pytest.approx(ct, abs=0)
    assert abs(sum(sc.b) - 1) < 1e-15 and abs(sum(sc.bembed) - 1) < 1e-15
    if name != "l2":  # l2's implicit/explicit abscissae differ by construction (row sums of its two tables)
        assert max(abs(x - y) for x, y in zip(sc.c, sc.ct)) < 2e-15  # c == c~ (stage times agree)
    if name in ("ars122", "a2"):
        assert sum(sc.b_exact) == 1 and sum(sc.bembed_exact) == 1
    elif name != "l2":  # Kennedy-Carpenter publish 25-digit rational approximations: sum(b) = 1 - O(1e-25)
        assert abs(sum(sc.b_exact) - 1) < Fraction(1, 10 ** 20) and abs(sum(sc.bembed_exact) - 1) < Fraction(1, 10 ** 20)
    # second-order coupling conditions
    if sc.order >= 2:
        assert abs(sum(bi * ci for bi, ci in zip(sc.b, sc.c)) - 0.5) < 1e-14


def test_method_table_matches_reference_mapping():
    # pnode/petsc_adjoint.py:641-656
    assert ptab.METHODS == otab.METHOD_TO_SCHEME
    assert ptab.METHODS["rk2"] == ("rk", "2b") and ptab.METHODS["dopri5"] == ("rk", "5dp")
    assert "midpoint" not in ptab.METHODS and "rk3" not in ptab.METHODS  # fall through to the 3bs default
    assert ptab.TS_DEFAULT == ("rk", "3bs")
