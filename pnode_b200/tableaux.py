"""Butcher tableaux of the schemes reachable through pnode's `method=` strings and the `-ts_rk_type` /
`-ts_arkimex_type` options (pnode/petsc_adjoint.py:641-656, 775).  The numbers are the published ones of each scheme
(PETSc registers the same tables under these names); rows are written as rational strings and
converted once, so `sum(b) == 1` can be asserted exactly (tests/test_tableaux.py).
"""
from fractions import Fraction
from math import sqrt


def _row(text):
    out = []
    for tok in text.split():
        out.append(float(tok[1:]) if tok.startswith("~") else Fraction(tok))
    return out


def _square(rows, s):
    M = [[Fraction(0)] * s for _ in range(s)]
    for i, r in enumerate(rows):
        for j, v in enumerate(_row(r)):
            M[i][j] = v
    return M


class Scheme:
    """kind 'rk': A explicit.  kind 'arkimex': At implicit part, A explicit part (bt == b, ct == c in all of these)."""

    def __init__(self, name, kind, order, A, b, bembed=None, At=None, fsal=False):
        self.name, self.kind, self.order, self.fsal = name, kind, order, fsal
        self.s = len(_row(b))
        self.A_exact = _square(A, self.s)
        self.b_exact = _row(b)
        self.bembed_exact = None if bembed is None else _row(bembed)
        self.At_exact = None if At is None else _square(At, self.s)
        f = lambda M: [[float(x) for x in r] for r in M]
        self.A = f(self.A_exact)
        self.b = [float(x) for x in self.b_exact]
        self.bembed = None if bembed is None else [float(x) for x in self.bembed_exact]
        self.c = [float(sum(r[1:], r[0])) for r in self.A_exact]
        if At is not None:
            self.At = f(self.At_exact)
            self.bt = list(self.b)
            self.ct = [float(sum(r[1:], r[0])) for r in self.At_exact]


RK = {
    "1fe": Scheme("1fe", "rk", 1, [""], "1"),
    "2a": Scheme("2a", "rk", 2, ["", "1"], "1/2 1/2", "1 0"),
    "2b": Scheme("2b", "rk", 2, ["", "2/3"], "1/4 3/4", "1 0"),
    "3": Scheme("3", "rk", 3, ["", "2/3", "-1/3 1"], "1/4 1/2 1/4"),
    "3bs": Scheme("3bs", "rk", 3, ["", "1/2", "0 3/4", "2/9 1/3 4/9"], "2/9 1/3 4/9 0", "7/24 1/4 1/3 1/8", fsal=True),
    "4": Scheme("4", "rk", 4, ["", "1/2", "0 1/2", "0 0 1"], "1/6 1/3 1/3 1/6"),
    "5f": Scheme(
        "5f", "rk", 5,
        ["", "1/4", "3/32 9/32", "1932/2197 -7200/2197 7296/2197", "439/216 -8 3680/513 -845/4104",
         "-8/27 2 -3544/2565 1859/4104 -11/40"],
        "16/135 0 6656/12825 28561/56430 -9/50 2/55", "25/216 0 1408/2565 2197/4104 -1/5 0"),
    # Bogacki-Shampine 5(4), 8 stages, FSAL ([PETSc] TSRK5BS); order conditions checked exactly in tests/test_tableaux.py
    "5bs": Scheme(
        "5bs", "rk", 5,
        ["", "1/6", "2/27 4/27", "183/1372 -162/343 1053/1372", "68/297 -4/11 42/143 1960/3861",
         "597/22528 81/352 63099/585728 58653/366080 4617/20480",
         "174197/959244 -30942/79937 8152137/19744439 666106/1039181 -29421/29068 482048/414219",
         "587/8064 0 4440339/15491840 24353/124800 387/44800 2152/5985 7267/94080"],
        "587/8064 0 4440339/15491840 24353/124800 387/44800 2152/5985 7267/94080 0",
        "2479/34992 0 123/416 612941/3411720 43/1440 2272/6561 79937/1113912 3293/556956", fsal=True),
    "5dp": Scheme(
        "5dp", "rk", 5,
        ["", "1/5", "3/40 9/40", "44/45 -56/15 32/9", "19372/6561 -25360/2187 64448/6561 -212/729",
         "9017/3168 -355/33 46732/5247 49/176 -5103/18656", "35/384 0 500/1113 125/192 -2187/6784 11/84"],
        "35/384 0 500/1113 125/192 -2187/6784 11/84 0",
        "5179/57600 0 7571/16695 393/640 -92097/339200 187/2100 1/40", fsal=True),
}

_G2 = repr(1.0 - 1.0 / sqrt(2.0))
_G2b = repr(1.0 - 2.0 * (1.0 - 1.0 / sqrt(2.0)))
_G3 = "1767732205903/4055673282236"
_ARK3_B = "1471266399579/7840856788654 -4482444167858/7529755066697 11266239266428/11593286722821 " + _G3
_ARK4_B = "82889/524892 0 15625/83664 69875/102672 -2260/8211 1/4"
_ARK5_B = ("-872700587467/9133579230613 0 0 22348218063261/9555858737531 -1143369518992/8141816002931 "
           "-39379526789629/19018526304540 32727382324388/42900044865799 41/200")

_R2 = sqrt(2.0)
_U2 = "~" + repr(1.0 - 1.0 / _R2)          # 1 - 1/sqrt(2)
_H2 = "~" + repr(1.0 / (2.0 * _R2))        # 1/(2 sqrt(2))
_AT2 = ["0", _U2 + " " + _U2, _H2 + " " + _H2 + " " + _U2]   # implicit part shared by 2c / 2d / 2e
_B2 = _H2 + " " + _H2 + " " + _U2
_BE2 = "~%r ~%r %s" % ((4.0 - _R2) / 8.0, (4.0 - _R2) / 8.0, _H2)
_A21 = "~" + repr(2.0 - _R2)

ARK = {
    # [PETSc] TSARKIMEX1BEE: backward Euler (two half steps) with one full step as the embedded solution; registered order 2
    "1bee": Scheme("1bee", "arkimex", 2, A=["0", "0", "0 1/2"], At=["1", "0 1/2", "0 1/2 1/2"], b="0 1/2 1/2", bembed="1 0 0"),
    # [PETSc] TSARKIMEX2C / 2D / 2E: one explicit + two L-stable implicit stages (gamma = 1 - 1/sqrt 2), three explicit parts
    "2c": Scheme("2c", "arkimex", 2, A=["0", _A21, "1/2 1/2"], At=_AT2, b=_B2, bembed=_BE2),
    "2d": Scheme("2d", "arkimex", 2, A=["0", _A21, "3/4 1/4"], At=_AT2, b=_B2, bembed=_BE2),
    "2e": Scheme("2e", "arkimex", 2, A=["0", _A21, "~%r ~%r" % ((3.0 - 2.0 * _R2) / 6.0, (3.0 + 2.0 * _R2) / 6.0)], At=_AT2,
                 b=_B2, bembed=_BE2),
    # Pareschi-Russo SSP2(3,3,2)
    "prssp2": Scheme("prssp2", "arkimex", 2, A=["0", "1/2", "1/2 1/2"], At=["1/4", "0 1/4", "1/3 1/3 1/3"], b="1/3 1/3 1/3"),
    # Boscarino-Pareschi-Russo BPR(3,5,3)
    "bpr3": Scheme("bpr3", "arkimex", 3, A=["0", "1", "4/9 2/9", "1/4 0 3/4", "1/4 0 3/4"],
                   At=["0", "1/2 1/2", "5/18 -1/9 1/2", "1/2 0 0 1/2", "1/4 0 3/4 -1/2 1/2"], b="1/4 0 3/4 -1/2 1/2"),
    # Ascher-Ruuth-Spiteri (4,4,3); the explicit weights serve as the embedded solution
    "ars443": Scheme("ars443", "arkimex", 3, A=["0", "1/2", "11/18 1/18", "5/6 -5/6 1/2", "1/4 7/4 3/4 -7/4"],
                     At=["0", "0 1/2", "0 1/6 1/2", "0 -1/2 1/2 1/2", "0 3/2 -3/2 1/2 1/2"], b="0 3/2 -3/2 1/2 1/2",
                     bembed="1/4 7/4 3/4 -7/4 0"),
    "ars122": Scheme("ars122", "arkimex", 2, A=["0", "1/2"], At=["0", "0 1/2"], b="0 1", bembed="1/2 1/2"),
    "a2": Scheme("a2", "arkimex", 2, A=["0", "1"], At=["0", "1/2 1/2"], b="1/2 1/2", bembed="0 1"),
    "l2": Scheme("l2", "arkimex", 2, A=["0", "1"], At=["~" + _G2, "~" + _G2b + " ~" + _G2], b="1/2 1/2", bembed="0 1"),
    "3": Scheme(
        "3", "arkimex", 3,
        A=["0", "1767732205903/2027836641118", "5535828885825/10492691773637 788022342437/10882634858940",
           "6485989280629/16251701735622 -4246266847089/9704473918619 10755448449292/10357097424841"],
        At=["0", _G3 + " " + _G3, "2746238789719/10658868560708 -640167445237/6845629431997 " + _G3, _ARK3_B],
        b=_ARK3_B,
        bembed="2756255671327/12835298489170 -10771552573575/22201958757719 9247589265047/10645013368117 "
               "2193209047091/5459859503100"),
    "4": Scheme(
        "4", "arkimex", 4,
        A=["0", "1/2", "13861/62500 6889/62500",
           "-116923316275/2393684061468 -2731218467317/15368042101831 9408046702089/11113171139209",
           "-451086348788/2902428689909 -2682348792572/7519795681897 12662868775082/11960479115383 "
           "3355817975965/11060851509271",
           "647845179188/3216320057751 73281519250/8382639484533 552539513391/3454668386233 "
           "3354512671639/8306763924573 4040/17871"],
        At=["0", "1/4 1/4", "8611/62500 -1743/31250 1/4", "5012029/34652500 -654441/2922500 174375/388108 1/4",
            "15267082809/155376265600 -71443401/120774400 730878875/902184768 2285395/8070912 1/4", _ARK4_B],
        b=_ARK4_B,
        bembed="4586570599/29645900160 0 178811875/945068544 814220225/1159782912 -3700637/11593932 61727/225920"),
    "5": Scheme(
        "5", "arkimex", 5,
        A=["0", "41/100", "367902744464/2072280473677 677623207551/8224143866563",
           "1268023523408/10340822734521 0 1029933939417/13636558850479",
           "14463281900351/6315353703477 0 66114435211212/5879490589093 -54053170152839/4284798021562",
           "14090043504691/34967701212078 0 15191511035443/11219624916014 -18461159152457/12425892160975 "
           "-281667163811/9011619295870",
           "19230459214898/13134317526959 0 21275331358303/2942455364971 -38145345988419/4862620318723 -1/8 -1/8",
           "-19977161125411/11928030595625 0 -40795976796054/6384907823539 177454434618887/12078138498510 "
           "782672205425/8267701900261 -69563011059811/9646580694205 7356628210526/4942186776405"],
        At=["0", "41/200 41/200", "41/400 -567603406766/11931857230679 41/200",
            "683785636431/9252920307686 0 -110385047103/1367015193373 41/200",
            "3016520224154/10081342136671 0 30586259806659/12414158314087 -22760509404356/11113319521817 41/200",
            "218866479029/1489978393911 0 638256894668/5436446318841 -1179710474555/5321154724896 "
            "-60928119172/8023461067671 41/200",
            "1020004230633/5715676835656 0 25762820946817/25263940353407 -2161375909145/9755907335909 "
            "-211217309593/5846859502534 -4269925059573/7827059040749 41/200", _ARK5_B],
        b=_ARK5_B,
        bembed="-975461918565/9796059967033 0 0 78070527104295/32432590147079 -548382580838/3424219808633 "
               "-33438840321285/15594753105479 3629800801594/4656183773603 4035322873751/18575991585200"),
}

# method= -> (ts type, scheme name).  Strings outside this table do NOT raise in the reference: the TS keeps the type set
# at petsc_adjoint.py:638 (TSRK, default 3bs) -- this is what "midpoint", "rk3" and "dopri5_fixed" get.
METHODS = {
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
TS_DEFAULT = ("rk", "3bs")
ARKIMEX_DEFAULT = "3"
RK_DEFAULT = "3bs"


def lookup(kind, name):
    if kind == "rk":
        if name not in RK:
            raise KeyError("unknown -ts_rk_type %r (known: %s)" % (name, ", ".join(sorted(RK))))
        return RK[name]
    if kind == "arkimex":
        if name not in ARK:
            raise KeyError("unknown -ts_arkimex_type %r (known: %s)" % (name, ", ".join(sorted(ARK))))
        return ARK[name]
    return None
