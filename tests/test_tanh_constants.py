"""The fp64 tanh of the fused sweeps (csrc/mlp_rk.cu, PNODE_F64_TANH_V 1) restated on the CPU from the constants in the
source: the constants are what their comments say (correctly rounded), and the evaluation scheme -- one-step argument
reduction, 1024-entry table, degree-3 polynomial with the adjusted r^2 coefficient, 1 - 2 / (e^{2x} + 1) -- is accurate to
6e-16 absolute against a 50-digit reference.  (The kernel itself is checked on the GPU: test_gpu_vecops.py.)"""
import math
import os
import re
from decimal import Decimal, getcontext

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pnode_b200", "csrc", "mlp_rk.cu")


def _constants():
    txt = open(SRC).read()
    out = {}
    for name in ("TANH_C", "TANH_NEG_STEP", "TANH_C2"):
        m = re.search(r"constexpr double %s = (-?[0-9.eE+-]+);" % name, txt)
        assert m, name
        out[name] = float(m.group(1))
    m = re.search(r"#else\s+constexpr int EXP_TAB = (\d+);", txt)
    out["EXP_TAB"] = int(m.group(1))
    return out


def test_constants_are_correctly_rounded():
    getcontext().prec = 60
    c = _constants()
    n = c["EXP_TAB"]
    assert n == 1024
    ln2 = Decimal(2).ln()
    assert c["TANH_C"] == float(Decimal(2 * n) / ln2)
    assert c["TANH_NEG_STEP"] == -float(ln2 / Decimal(2 * n))
    h = math.log(2.0) / (2 * n)
    delta = (math.sqrt(2.0) - 1.0) * h * h / 12.0
    assert abs(c["TANH_C2"] - (2.0 + 4.0 * delta)) < 1e-15


def test_scheme_is_accurate_to_a_few_ulp_absolute():
    getcontext().prec = 50
    c = _constants()
    n = c["EXP_TAB"]
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-6, 6, 6000), rng.uniform(-1e-3, 1e-3, 1000), rng.uniform(-24, 24, 2000),
                         np.array([0.0, 1e-300, -1e-12, 23.99, -23.99, 24.0, 30.0, -1e9])])
    magic = 6755399441055744.0
    x = np.clip(xs, -24.0, 24.0)
    kf = (x * c["TANH_C"] + magic) - magic
    k = kf.astype(np.int64)
    # fma(kf, step, x): the product is exact inside the FMA
    rh = np.array([float(Decimal(float(a)) + Decimal(float(b)) * Decimal(c["TANH_NEG_STEP"])) for a, b in zip(x, kf)])
    p = rh * (4.0 / 3.0) + c["TANH_C2"]
    p = p * rh + 2.0
    q = p * rh + 1.0
    table = np.array([float(Decimal(2) ** (Decimal(j) / Decimal(n))) for j in range(n)])
    s = np.ldexp(table[k % n], (k >> int(math.log2(n))).astype(np.int64))
    d = s * q + 1.0
    got = 1.0 - 2.0 / d
    ref = np.array([float(((Decimal(float(v)) * 2).exp() - 1) / ((Decimal(float(v)) * 2).exp() + 1)) for v in x])
    assert float(np.abs(got - ref).max()) < 6e-16
    assert got[list(xs).index(0.0)] == 0.0 and np.all(np.abs(got) <= 1.0)
