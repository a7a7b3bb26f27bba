"""Pins the oracle on the ONLY known-answer values the reference holds for this path
(/root/reference/tests/test_pnode.py:133-201): ROBER, three steps [1e-5, 9e-5, 9e-4], fp64, -ts_adapt_type none."""
import pytest
import torch

from oracle import OracleODEPetsc
from _problems import PETSC_ARGS, ROBER_STEPS, ROBER_T, Rober, RoberEX, RoberIM, rober_truth


def _run(kw, funcs):
    true_y = rober_truth()
    ode = OracleODEPetsc(PETSC_ARGS)
    ode.setupTS(true_y[0], funcs[0], step_size=ROBER_STEPS, enable_adjoint=True, **kw)
    pred = ode.odeint_adjoint(true_y[0], ROBER_T)
    loss = torch.mean(torch.abs(pred - true_y))
    loss.backward()
    return loss.item(), torch.std(torch.abs(pred - true_y)).item()


def test_scalartype():
    import numpy as np
    from petsc4py import PETSc  # the shim

    assert PETSc.ScalarType == np.float64  # test_pnode.py:127-130


def test_golden_implicit_cn():
    loss, std = _run(dict(method="cn", implicit_form=True), (Rober(),))
    assert loss == pytest.approx(1.85e-6, abs=1e-6)  # test_pnode.py:151
    assert std == pytest.approx(3.36e-6, abs=1e-6)  # test_pnode.py:152
    assert loss == pytest.approx(1.8492e-6, rel=2e-4)  # SURVEY.md appendix D.1 (scratch NumPy probe)


def test_golden_imex_ark3():
    f_im, f_ex = RoberIM(), RoberEX()
    loss, std = _run(dict(method="imex", implicit_form=True, imex_form=True, func2=f_ex), (f_im,))
    assert loss == pytest.approx(3.11e-6, abs=3e-6)  # test_pnode.py:179
    assert std == pytest.approx(5.65e-6, abs=3e-6)  # test_pnode.py:180
    assert loss == pytest.approx(3.1138e-6, rel=2e-4)
    assert std == pytest.approx(5.6592e-6, rel=2e-4)


def test_golden_explicit_default_3bs():
    # method="rk3" is not in setupTS's table => PETSc's default TSRK type 3bs
    loss, std = _run(dict(method="rk3"), (Rober(),))
    assert loss == pytest.approx(1.85e-6, abs=1e-6)  # test_pnode.py:200
    assert std == pytest.approx(3.21e-6, abs=1e-6)  # test_pnode.py:201
    assert loss == pytest.approx(1.8495e-6, rel=2e-4)


def test_golden_discriminates_scheme():
    # rk4 gives 2.09e-6 / 4.08e-6: the golden really pins "3bs", not just "some RK"
    loss, std = _run(dict(method="rk4"), (Rober(),))
    assert loss == pytest.approx(2.0922e-6, rel=1e-3)
    assert std == pytest.approx(4.0756e-6, rel=1e-3)
