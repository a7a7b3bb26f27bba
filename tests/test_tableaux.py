"""Tableau self-tests (the only guard without PETSc) + product tables == oracle tables."""
from fractions import Fraction

import pytest

from oracle import tableaux as otab
from pnode_b200 import tableaux as ptab


def _trees(n):
    """Rooted trees of order n as nested sorted tuples."""
    import itertools
    if n == 1:
        return [()]

    def parts(m, maxp):
        if m == 0:
            yield []
            return
        for q in range(min(m, maxp), 0, -1):
            for rest in parts(m - q, q):
                yield [q] + rest
    out = set()
    for part in parts(n - 1, n - 1):
        for combo in itertools.product(*[_trees(q) for q in part]):
            out.add(tuple(sorted(combo)))
    return sorted(out)


def _order(t):
    return 1 + sum(_order(c) for c in t)


def _gamma(t):
    g = _order(t)
    for c in t:
        g *= _gamma(c)
    return g


def _colourings(t):
    import itertools
    if not t:
        return [[]]
    per = [[(cc, sub) for cc in (0, 1) for sub in _colourings(c)] for c in t]
    return [list(x) for x in itertools.product(*per)]


def _phi(t, col, mats, s, one):
    """Stage weights of tree t; col gives, per child, which matrix (0 explicit / 1 implicit) carries the edge."""
    v = [one] * s
    for c, (cc, sub) in zip(t, col):
        pc = _phi(c, sub, mats, s, one)
        M = mats[cc]
        v = [v[i] * sum(M[i][j] * pc[j] for j in range(s)) for i in range(s)]
    return v


def _order_defect(mats, b, p, one):
    """max over all (coloured) rooted trees of order <= p of |b . phi(t) - 1/gamma(t)|."""
    worst = 0
    for n in range(1, p + 1):
        for t in _trees(n):
            if len(mats) == 2:
                cols = _colourings(t)
            else:  # single tableau: every edge uses mats[0]
                def mono(tt):
                    return [(0, mono(c)) for c in tt]
                cols = [mono(t)]
            for col in cols:
                val = sum(bi * vi for bi, vi in zip(b, _phi(t, col, mats, len(b), one)))
                worst = max(worst, abs(val - one / _gamma(t)))
    return worst


@pytest.mark.parametrize("name", sorted(ptab.RK))
def test_rk_order_conditions_exact(name):
    """ALL rooted-tree order conditions up to the stated order, in rational arithmetic (and the embedded weights one order
    lower): the guard for tables restated without PETSc at hand, e.g. the 8-stage Bogacki-Shampine 5(4)."""
    sc = ptab.RK[name]
    assert _order_defect((sc.A_exact,), sc.b_exact, sc.order, Fraction(1)) == 0
    assert _order_defect((sc.A_exact,), sc.b_exact, sc.order + 1, Fraction(1)) != 0
    if sc.bembed_exact is not None:
        assert _order_defect((sc.A_exact,), sc.bembed_exact, sc.order - 1, Fraction(1)) == 0


@pytest.mark.parametrize("name", sorted(ptab.ARK))
def test_ark_additive_order_conditions(name):
    """Every colouring of every rooted tree (explicit / implicit matrix per edge): the coupling conditions of the additive
    scheme up to its order.  1bee is registered by PETSc with order 2 for the step controller but is a first-order scheme."""
    sc = ptab.ARK[name]
    p = 1 if name == "1bee" else sc.order
    tol = 1e-14
    assert _order_defect((sc.A, sc.At), sc.b, p, 1.0) < tol
    assert _order_defect((sc.A, sc.At), sc.b, p + 1, 1.0) > 1e-6
    if sc.bembed is not None:
        assert _order_defect((sc.A, sc.At), sc.bembed, max(p - 1, 1), 1.0) < tol


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
    assert abs(sum(sc.b) - 1) < 1e-15 and (sc.bembed is None or abs(sum(sc.bembed) - 1) < 1e-15)
    if name not in ("l2", "1bee", "prssp2"):  # these have different implicit / explicit abscissae by construction
        assert max(abs(x - y) for x, y in zip(sc.c, sc.ct)) < 2e-15  # c == c~ (stage times agree)
    if name in ("ars122", "a2", "1bee", "ars443"):
        assert sum(sc.b_exact) == 1 and sum(sc.bembed_exact) == 1
    elif name in ("prssp2", "bpr3"):
        assert sum(sc.b_exact) == 1
    elif name not in ("l2", "2c", "2d", "2e"):  # Kennedy-Carpenter publish 25-digit rational approximations: sum(b) = 1 - O(1e-25)
        assert abs(sum(sc.b_exact) - 1) < Fraction(1, 10 ** 20) and abs(sum(sc.bembed_exact) - 1) < Fraction(1, 10 ** 20)
    # second-order coupling conditions
    if sc.order >= 2 and name != "1bee":
        assert abs(sum(bi * ci for bi, ci in zip(sc.b, sc.c)) - 0.5) < 1e-14


def test_method_table_matches_reference_mapping():
    # pnode/petsc_adjoint.py:641-656
    assert ptab.METHODS == otab.METHOD_TO_SCHEME
    assert ptab.METHODS["rk2"] == ("rk", "2b") and ptab.METHODS["dopri5"] == ("rk", "5dp")
    assert "midpoint" not in ptab.METHODS and "rk3" not in ptab.METHODS  # fall through to the 3bs default
    assert ptab.TS_DEFAULT == ("rk", "3bs")
