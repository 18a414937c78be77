"""`saveat` inside control pieces (problem.kwargs; lib/CorleoneOED/src/oed.jl:106-113): the checker's restatement of Tsit5's
dense output (OrdinaryDiffEq's interpolant for Tsit5, a third-party algorithm: Tsitouras 2011) pinned on its defining
properties and on closed forms.  CPU only."""
import ctypes

import numpy as np
import pytest

from oracle import pyoracle

# Tsit5 tableau
C_ = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
A = np.zeros((7, 7))
A[1, 0] = 0.161
A[2, :2] = [-0.008480655492356989, 0.335480655492357]
A[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
A[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
A[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]
A[6, :6] = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]


def weights(theta):
    L = pyoracle.lib()
    L.co_tsit5_interp_weights.argtypes = [ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    L.co_tsit5_interp_weights.restype = None
    b = np.zeros(7)
    L.co_tsit5_interp_weights(float(theta), b.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return b


def test_interpolant_ends_and_order_conditions():
    """b(0) = 0, b(1) = the propagation weights (row 7 of the tableau, b_7(1) = 0), and for every theta the order
    conditions of a 4th-order continuous extension: sum b = theta, b.c = theta^2/2, b.c^2 = theta^3/3, b.Ac = theta^3/6,
    b.c^3 = theta^4/4, (b o c).Ac = theta^4/8, b.Ac^2 = theta^4/12, b.AAc = theta^4/24"""
    assert np.array_equal(weights(0.0), np.zeros(7))
    assert np.max(np.abs(weights(1.0) - A[6])) <= 1e-14   # coefficients up to 88: rounding of the Horner sums
    Ac, Ac2 = A @ C_, A @ (C_ ** 2)
    for th in np.linspace(0.05, 1.0, 20):
        b = weights(th)
        cond = [(b.sum(), th), (b @ C_, th ** 2 / 2), (b @ C_ ** 2, th ** 3 / 3), (b @ Ac, th ** 3 / 6),
                (b @ C_ ** 3, th ** 4 / 4), ((b * C_) @ Ac, th ** 4 / 8), (b @ Ac2, th ** 4 / 12), (b @ (A @ Ac), th ** 4 / 24)]
        for got, want in cond:
            assert abs(got - want) <= 3e-14, (th, got, want)


def lin_spec(saveat, controls=()):
    return dict(model="lin1d", u0=[1.0], p0=[-2.0], tspan=(0.0, 1.0), controls=list(controls), saveat=saveat)


def test_samples_times_and_layout():
    """a piece saves its start (first piece only), the saveat times STRICTLY inside it, and its end (OrdinaryDiffEq keeps
    t0 < t < tf of `saveat`; the ends follow save_start / save_end, src/single_shooting.jl:443-453)"""
    O = pyoracle.OracleLayer(lin_spec([0.0, 0.1, 0.35, 0.5, 0.77, 1.0], [(1, np.array([0.0, 0.5]), np.array([-2.0, -1.0]))]))
    r = O.eval(O.default_ps(), K=8)
    assert np.array_equal(r["t"], [0.0, 0.1, 0.35, 0.5, 0.77, 1.0])
    assert O.nsamples == 6
    # the control row of a sample is the control of the piece the sample belongs to (single_shooting.jl:456)
    assert np.array_equal(r["u"][1], [-2.0, -2.0, -2.0, -2.0, -1.0, -1.0])
    # without saveat: the piece ends only
    O0 = pyoracle.OracleLayer(lin_spec(None, [(1, np.array([0.0, 0.5]), np.array([-2.0, -1.0]))]))
    r0 = O0.eval(O0.default_ps(), K=8)
    assert np.array_equal(r0["t"], [0.0, 0.5, 1.0]) and np.array_equal(r0["u"][:, [1, 2]], r["u"][:, [3, 5]])


@pytest.mark.parametrize("method", ["tsit5", "tsit5_adaptive"])
def test_dense_output_converges_to_the_closed_form(method):
    """x' = p x: the interior samples are x(t) = e^{pt} to the accuracy of the interpolant (4th order: the error falls
    by ~2^5 per halving of the fixed step); the adaptive mode reaches its tolerance"""
    sv = np.array([0.013, 0.21, 0.4999, 0.5, 0.83, 0.97])
    O = pyoracle.OracleLayer(lin_spec(sv))
    exact = np.exp(-2.0 * np.r_[0.0, sv, 1.0])
    if method == "tsit5":
        errs = []
        for K in (4, 8, 16, 32):
            r = O.eval(O.default_ps(), K=K)
            errs.append(np.max(np.abs(r["u"][0] - exact)))
        rate = np.log2(errs[0] / errs[-1]) / 3.0   # per halving, averaged (where the times fall inside a step changes with K)
        assert errs[-1] < 2e-10 and np.all(np.diff(errs) < 0) and rate > 4.5, (errs, rate)
    else:
        pyoracle.lib().co_set_tolerances(ctypes.c_double(1e-10), ctypes.c_double(1e-10))
        try:
            r = O.eval(O.default_ps(), method="tsit5_adaptive")
        finally:
            pyoracle.lib().co_set_tolerances(ctypes.c_double(1e-6), ctypes.c_double(1e-3))
        assert np.max(np.abs(r["u"][0] - exact)) < 5e-9


def test_tangents_of_the_samples_are_the_derivative_of_the_interpolated_values():
    """the dense output is linear in (y, k_1..k_7), so the tangents pass through it: d x(t_s) / d p = t_s e^{p t_s}"""
    sv = np.array([0.3, 0.61])
    O = pyoracle.OracleLayer(dict(model="lin1d", u0=[1.0], p0=[-2.0], tspan=(0.0, 1.0), controls=[], saveat=sv))
    u, J = O.traj_jacobian(O.default_ps(), K=64)
    t = np.r_[0.0, sv, 1.0]
    assert np.max(np.abs(J[0, :, 0] - t * np.exp(-2.0 * t))) < 1e-10
