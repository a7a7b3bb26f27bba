"""ORACLE (test infrastructure, not product): Butcher tableaux of the PETSc TS schemes pnode selects.

pnode never stores a tableau itself; it names them (`setRKType("4")`, `TS.Type.ARKIMEX`, ... --
/root/reference/pnode/petsc_adjoint.py:641-656) and PETSc (absent from /root/reference, version unpinned:
/root/reference/setup.py:9, /root/reference/.github/workflows/build.sh:4) supplies the numbers.  They are restated
here from the published schemes (Bogacki-Shampine 1989, Dormand-Prince 1980, Kennedy-Carpenter 2003 ARK3(2)4L[2]SA /
ARK4(3)6L[2]SA, Ascher-Ruuth-Spiteri 1997, Giraldo et al. 2013 for `l2`) as exact rationals, SURVEY.md appendix B.

Everything is kept as `fractions.Fraction` (or python floats for the irrational `l2` gamma) so that the tests can check
sum(b) == 1 exactly and the order conditions in rational arithmetic.
"""
from fractions import Fraction as Fr
from math import sqrt


def _f(x):
    return x if isinstance(x, float) else Fr(x)


class RKTableau:
    """Explicit RK tableau.  A is strictly lower triangular (list of rows), b the completion weights,
    bembed the embedded (order-1 lower) weights or None, FSAL = last row of A equals b."""

    def __init__(self, name, order, A, b, bembed=None, fsal=False):
        self.name = name
        self.order = order
        self.s = len(b)
        self.A = [[_f(A[i][j]) if j < len(A[i]) else Fr(0) for j in range(self.s)] for i in range(self.s)]
        self.b = [_f(x) for x in b]
        self.bembed = None if bembed is None else [_f(x) for x in bembed]
        self.c = [sum(row, Fr(0)) for row in self.A]
        self.fsal = fsal

    def floats(self):
        A = [[float(x) for x in row] for row in self.A]
        b = [float(x) for x in self.b]
        be = None if self.bembed is None else [float(x) for x in self.bembed]
        c = [float(x) for x in self.c]
        return A, b, be, c


RK = {}
RK["1fe"] = RKTableau("1fe", 1, [[]], [1])
RK["2a"] = RKTableau("2a", 2, [[], [1]], [Fr(1, 2), Fr(1, 2)], [1, 0])
RK["2b"] = RKTableau("2b", 2, [[], [Fr(2, 3)]], [Fr(1, 4), Fr(3, 4)], [1, 0])
RK["3"] = RKTableau("3", 3, [[], [Fr(2, 3)], [Fr(-1, 3), 1]], [Fr(1, 4), Fr(1, 2), Fr(1, 4)])
RK["3bs"] = RKTableau(
    "3bs", 3,
    [[], [Fr(1, 2)], [0, Fr(3, 4)], [Fr(2, 9), Fr(1, 3), Fr(4, 9)]],
    [Fr(2, 9), Fr(1, 3), Fr(4, 9), 0],
    [Fr(7, 24), Fr(1, 4), Fr(1, 3), Fr(1, 8)],
    fsal=True,
)
RK["4"] = RKTableau("4", 4, [[], [Fr(1, 2)], [0, Fr(1, 2)], [0, 0, 1]], [Fr(1, 6), Fr(1, 3), Fr(1, 3), Fr(1, 6)])
RK["5dp"] = RKTableau(
    "5dp", 5,
    [
        [],
        [Fr(1, 5)],
        [Fr(3, 40), Fr(9, 40)],
        [Fr(44, 45), Fr(-56, 15), Fr(32, 9)],
        [Fr(19372, 6561), Fr(-25360, 2187), Fr(64448, 6561), Fr(-212, 729)],
        [Fr(9017, 3168), Fr(-355, 33), Fr(46732, 5247), Fr(49, 176), Fr(-5103, 18656)],
        [Fr(35, 384), 0, Fr(500, 1113), Fr(125, 192), Fr(-2187, 6784), Fr(11, 84)],
    ],
    [Fr(35, 384), 0, Fr(500, 1113), Fr(125, 192), Fr(-2187, 6784), Fr(11, 84), 0],
    [Fr(5179, 57600), 0, Fr(7571, 16695), Fr(393, 640), Fr(-92097, 339200), Fr(187, 2100), Fr(1, 40)],
    fsal=True,
)
RK["5f"] = RKTableau(
    "5f", 5,
    [
        [],
        [Fr(1, 4)],
        [Fr(3, 32), Fr(9, 32)],
        [Fr(1932, 2197), Fr(-7200, 2197), Fr(7296, 2197)],
        [Fr(439, 216), -8, Fr(3680, 513), Fr(-845, 4104)],
        [Fr(-8, 27), 2, Fr(-3544, 2565), Fr(1859, 4104), Fr(-11, 40)],
    ],
    [Fr(16, 135), 0, Fr(6656, 12825), Fr(28561, 56430), Fr(-9, 50), Fr(2, 55)],
    [Fr(25, 216), 0, Fr(1408, 2565), Fr(2197, 4104), Fr(-1, 5), 0],
)


# Bogacki-Shampine 5(4), 8 stages, FSAL (Bogacki & Shampine 1996; [PETSc] TSRK5BS)
RK["5bs"] = RKTableau(
    "5bs", 5,
    [[], [Fr(1, 6)], [Fr(2, 27), Fr(4, 27)], [Fr(183, 1372), Fr(-162, 343), Fr(1053, 1372)],
     [Fr(68, 297), Fr(-4, 11), Fr(42, 143), Fr(1960, 3861)],
     [Fr(597, 22528), Fr(81, 352), Fr(63099, 585728), Fr(58653, 366080), Fr(4617, 20480)],
     [Fr(174197, 959244), Fr(-30942, 79937), Fr(8152137, 19744439), Fr(666106, 1039181), Fr(-29421, 29068),
      Fr(482048, 414219)],
     [Fr(587, 8064), 0, Fr(4440339, 15491840), Fr(24353, 124800), Fr(387, 44800), Fr(2152, 5985), Fr(7267, 94080)]],
    [Fr(587, 8064), 0, Fr(4440339, 15491840), Fr(24353, 124800), Fr(387, 44800), Fr(2152, 5985), Fr(7267, 94080), 0],
    [Fr(2479, 34992), 0, Fr(123, 416), Fr(612941, 3411720), Fr(43, 1440), Fr(2272, 6561), Fr(79937, 1113912),
     Fr(3293, 556956)], fsal=True)


class ARKTableau:
    """Additive (implicit At / explicit A) tableau.  PETSc's TSARKIMEX keeps bt == b and ct == c for all schemes below."""

    def __init__(self, name, order, At, A, b, bembed=None):
        self.name = name
        self.order = order
        self.s = len(b)
        s = self.s
        self.At = [[_f(At[i][j]) if j < len(At[i]) else Fr(0) for j in range(s)] for i in range(s)]
        self.A = [[_f(A[i][j]) if j < len(A[i]) else Fr(0) for j in range(s)] for i in range(s)]
        self.b = [_f(x) for x in b]
        self.bt = list(self.b)
        self.bembed = None if bembed is None else [_f(x) for x in bembed]
        self.c = [sum(row[1:], row[0]) for row in self.A]
        self.ct = [sum(row[1:], row[0]) for row in self.At]

    def floats(self):
        fl = lambda M: [[float(x) for x in r] for r in M]
        be = None if self.bembed is None else [float(x) for x in self.bembed]
        return fl(self.At), fl(self.A), [float(x) for x in self.bt], [float(x) for x in self.b], be, \
            [float(x) for x in self.ct], [float(x) for x in self.c]


ARK = {}
# [PETSc] TSARKIMEX1BEE (backward Euler as two half steps, one full step embedded; registered order 2), 2C / 2D / 2E (Giraldo et
# al. 2013 family: gamma = 1 - 1/sqrt 2), PRSSP2 (Pareschi-Russo SSP2(3,3,2)), BPR3 (Boscarino-Pareschi-Russo BPR(3,5,3)),
# ARS443 (Ascher-Ruuth-Spiteri 1997).  Restated from the published schemes; additive order conditions incl. coupling are
# checked in tests/test_tableaux.py.
ARK["1bee"] = ARKTableau("1bee", 2, [[1, 0, 0], [0, Fr(1, 2), 0], [0, Fr(1, 2), Fr(1, 2)]], [[0, 0, 0], [0, 0, 0], [0, Fr(1, 2), 0]],
                         [0, Fr(1, 2), Fr(1, 2)], [1, 0, 0])
_r2 = sqrt(2.0)
_u2, _h2 = 1.0 - 1.0 / _r2, 1.0 / (2.0 * _r2)
_at2 = [[0.0, 0.0, 0.0], [_u2, _u2, 0.0], [_h2, _h2, _u2]]
_be2 = [(4.0 - _r2) / 8.0, (4.0 - _r2) / 8.0, _h2]
ARK["2c"] = ARKTableau("2c", 2, _at2, [[0.0, 0.0, 0.0], [2.0 - _r2, 0.0, 0.0], [0.5, 0.5, 0.0]], [_h2, _h2, _u2], _be2)
ARK["2d"] = ARKTableau("2d", 2, _at2, [[0.0, 0.0, 0.0], [2.0 - _r2, 0.0, 0.0], [0.75, 0.25, 0.0]], [_h2, _h2, _u2], _be2)
ARK["2e"] = ARKTableau("2e", 2, _at2, [[0.0, 0.0, 0.0], [2.0 - _r2, 0.0, 0.0],
                                       [(3.0 - 2.0 * _r2) / 6.0, (3.0 + 2.0 * _r2) / 6.0, 0.0]], [_h2, _h2, _u2], _be2)
ARK["prssp2"] = ARKTableau("prssp2", 2, [[Fr(1, 4), 0, 0], [0, Fr(1, 4), 0], [Fr(1, 3), Fr(1, 3), Fr(1, 3)]],
                           [[0, 0, 0], [Fr(1, 2), 0, 0], [Fr(1, 2), Fr(1, 2), 0]], [Fr(1, 3), Fr(1, 3), Fr(1, 3)])
ARK["bpr3"] = ARKTableau(
    "bpr3", 3,
    [[0] * 5, [Fr(1, 2), Fr(1, 2), 0, 0, 0], [Fr(5, 18), Fr(-1, 9), Fr(1, 2), 0, 0], [Fr(1, 2), 0, 0, Fr(1, 2), 0],
     [Fr(1, 4), 0, Fr(3, 4), Fr(-1, 2), Fr(1, 2)]],
    [[0] * 5, [1, 0, 0, 0, 0], [Fr(4, 9), Fr(2, 9), 0, 0, 0], [Fr(1, 4), 0, Fr(3, 4), 0, 0], [Fr(1, 4), 0, Fr(3, 4), 0, 0]],
    [Fr(1, 4), 0, Fr(3, 4), Fr(-1, 2), Fr(1, 2)])
ARK["ars443"] = ARKTableau(
    "ars443", 3,
    [[0] * 5, [0, Fr(1, 2), 0, 0, 0], [0, Fr(1, 6), Fr(1, 2), 0, 0], [0, Fr(-1, 2), Fr(1, 2), Fr(1, 2), 0],
     [0, Fr(3, 2), Fr(-3, 2), Fr(1, 2), Fr(1, 2)]],
    [[0] * 5, [Fr(1, 2), 0, 0, 0, 0], [Fr(11, 18), Fr(1, 18), 0, 0, 0], [Fr(5, 6), Fr(-5, 6), Fr(1, 2), 0, 0],
     [Fr(1, 4), Fr(7, 4), Fr(3, 4), Fr(-7, 4), 0]],
    [0, Fr(3, 2), Fr(-3, 2), Fr(1, 2), Fr(1, 2)], [Fr(1, 4), Fr(7, 4), Fr(3, 4), Fr(-7, 4), 0])
ARK["ars122"] = ARKTableau("ars122", 2, [[0, 0], [0, Fr(1, 2)]], [[0, 0], [Fr(1, 2), 0]], [0, 1], [Fr(1, 2), Fr(1, 2)])
ARK["a2"] = ARKTableau("a2", 2, [[0, 0], [Fr(1, 2), Fr(1, 2)]], [[0, 0], [1, 0]], [Fr(1, 2), Fr(1, 2)], [0, 1])
_g = 1.0 - 1.0 / sqrt(2.0)
ARK["l2"] = ARKTableau("l2", 2, [[_g, 0.0], [1.0 - 2.0 * _g, _g]], [[0.0, 0.0], [1.0, 0.0]], [0.5, 0.5], [0.0, 1.0])

_g3 = Fr(1767732205903, 4055673282236)
_ark3_last = [Fr(1471266399579, 7840856788654), Fr(-4482444167858, 7529755066697), Fr(11266239266428, 11593286722821), _g3]
ARK["3"] = ARKTableau(
    "3", 3,
    [
        [0],
        [_g3, _g3],
        [Fr(2746238789719, 10658868560708), Fr(-640167445237, 6845629431997), _g3],
        _ark3_last,
    ],
    [
        [0],
        [Fr(1767732205903, 2027836641118)],
        [Fr(5535828885825, 10492691773637), Fr(788022342437, 10882634858940)],
        [Fr(6485989280629, 16251701735622), Fr(-4246266847089, 9704473918619), Fr(10755448449292, 10357097424841)],
    ],
    _ark3_last,
    [Fr(2756255671327, 12835298489170), Fr(-10771552573575, 22201958757719), Fr(9247589265047, 10645013368117),
     Fr(2193209047091, 5459859503100)],
)

_q = Fr(1, 4)
_ark4_last = [Fr(82889, 524892), 0, Fr(15625, 83664), Fr(69875, 102672), Fr(-2260, 8211), _q]
ARK["4"] = ARKTableau(
    "4", 4,
    [
        [0],
        [_q, _q],
        [Fr(8611, 62500), Fr(-1743, 31250), _q],
        [Fr(5012029, 34652500), Fr(-654441, 2922500), Fr(174375, 388108), _q],
        [Fr(15267082809, 155376265600), Fr(-71443401, 120774400), Fr(730878875, 902184768), Fr(2285395, 8070912), _q],
        _ark4_last,
    ],
    [
        [0],
        [Fr(1, 2)],
        [Fr(13861, 62500), Fr(6889, 62500)],
        [Fr(-116923316275, 2393684061468), Fr(-2731218467317, 15368042101831), Fr(9408046702089, 11113171139209)],
        [Fr(-451086348788, 2902428689909), Fr(-2682348792572, 7519795681897), Fr(12662868775082, 11960479115383),
         Fr(3355817975965, 11060851509271)],
        [Fr(647845179188, 3216320057751), Fr(73281519250, 8382639484533), Fr(552539513391, 3454668386233),
         Fr(3354512671639, 8306763924573), Fr(4040, 17871)],
    ],
    _ark4_last,
    [Fr(4586570599, 29645900160), 0, Fr(178811875, 945068544), Fr(814220225, 1159782912), Fr(-3700637, 11593932),
     Fr(61727, 225920)],
)

_g5 = Fr(41, 200)
_ark5_last = [Fr(-872700587467, 9133579230613), 0, 0, Fr(22348218063261, 9555858737531), Fr(-1143369518992, 8141816002931),
              Fr(-39379526789629, 19018526304540), Fr(32727382324388, 42900044865799), _g5]
ARK["5"] = ARKTableau(
    "5", 5,
    [
        [0],
        [_g5, _g5],
        [Fr(41, 400), Fr(-567603406766, 11931857230679), _g5],
        [Fr(683785636431, 9252920307686), 0, Fr(-110385047103, 1367015193373), _g5],
        [Fr(3016520224154, 10081342136671), 0, Fr(30586259806659, 12414158314087), Fr(-22760509404356, 11113319521817), _g5],
        [Fr(218866479029, 1489978393911), 0, Fr(638256894668, 5436446318841), Fr(-1179710474555, 5321154724896),
         Fr(-60928119172, 8023461067671), _g5],
        [Fr(1020004230633, 5715676835656), 0, Fr(25762820946817, 25263940353407), Fr(-2161375909145, 9755907335909),
         Fr(-211217309593, 5846859502534), Fr(-4269925059573, 7827059040749), _g5],
        _ark5_last,
    ],
    [
        [0],
        [Fr(41, 100)],
        [Fr(367902744464, 2072280473677), Fr(677623207551, 8224143866563)],
        [Fr(1268023523408, 10340822734521), 0, Fr(1029933939417, 13636558850479)],
        [Fr(14463281900351, 6315353703477), 0, Fr(66114435211212, 5879490589093), Fr(-54053170152839, 4284798021562)],
        [Fr(14090043504691, 34967701212078), 0, Fr(15191511035443, 11219624916014), Fr(-18461159152457, 12425892160975),
         Fr(-281667163811, 9011619295870)],
        [Fr(19230459214898, 13134317526959), 0, Fr(21275331358303, 2942455364971), Fr(-38145345988419, 4862620318723),
         Fr(-1, 8), Fr(-1, 8)],
        [Fr(-19977161125411, 11928030595625), 0, Fr(-40795976796054, 6384907823539), Fr(177454434618887, 12078138498510),
         Fr(782672205425, 8267701900261), Fr(-69563011059811, 9646580694205), Fr(7356628210526, 4942186776405)],
    ],
    _ark5_last,
    [Fr(-975461918565, 9796059967033), 0, 0, Fr(78070527104295, 32432590147079), Fr(-548382580838, 3424219808633),
     Fr(-33438840321285, 15594753105479), Fr(3629800801594, 4656183773603), Fr(4035322873751, 18575991585200)],
)

# pnode's method= strings -> PETSc scheme (petsc_adjoint.py:641-656).  Anything not listed keeps the TS default
# set at petsc_adjoint.py:638 (TSRK, whose default type is 3bs) -- SURVEY.md appendix C.1.
METHOD_TO_SCHEME = {
    "euler": ("rk", "1fe"),
    "rk2": ("rk", "2b"),
    "bosh3": ("rk", "3bs"),
    "fixed_bosh3": ("rk", "3bs"),
    "rk4": ("rk", "4"),
    "dopri5": ("rk", "5dp"),
    "fixed_dopri5": ("rk", "5dp"),
    "beuler": ("beuler", None),
    "cn": ("cn", None),
    "imex": ("arkimex", "3"),
}
DEFAULT_SCHEME = ("rk", "3bs")
