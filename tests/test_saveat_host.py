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


def test_saveat_samples_vs_extended_precision_restatement():
    """An independent restatement of the fixed-step Tsit5 loop WITH its dense output in 40-digit arithmetic (mpmath): the
    checker's saveat samples agree to rounding, so the C interpolation arithmetic (weights, step bookkeeping, which step a
    time belongs to) is pinned by something that shares no code with it."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    Amp = [[mp.mpf(float(A[i, j])) for j in range(i)] for i in range(7)]
    R = [["1.0", "-2.763706197274826", "2.9132554618219126", "-1.0530884977290216"],
         ["0.0", "0.13169999999999998", "-0.2234", "0.1017"],
         ["0.0", "3.9302962368947516", "-5.941033872131505", "2.490627285651253"],
         ["0.0", "-12.411077166933676", "30.33818863028232", "-16.548102889244902"],
         ["0.0", "37.50931341651104", "-88.1789048947664", "47.37952196281928"],
         ["0.0", "-27.896526289197286", "65.09189467479366", "-34.87065786149661"],
         ["0.0", "1.5", "-4.0", "2.5"]]
    R = [[mp.mpf(float(x)) for x in row] for row in R]
    rng = np.random.default_rng(9)
    grid = np.array([0.0, 0.7, 1.3, 2.05])
    ctrl = rng.uniform(0, 1, len(grid))
    saveat = np.array([0.05, 0.31, 0.7, 0.99, 1.3000001, 2.0, 2.4, 2.9])     # inside pieces, on piece ends, just behind one
    spec = dict(model="lotka3", u0=[0.5, 0.7, 0.0], p0=[0.0, 1.0, 1.0], tspan=(0.0, 3.0), controls=[(1, grid, ctrl)],
                tunable_ic=[], quadrature_indices=[], shooting_points=None, saveat=list(saveat))
    O = pyoracle.OracleLayer(spec)
    K = 3
    r = O.eval(O.default_ps(), K=K)

    def ff(y, u):
        x12 = y[0] * y[1]
        return [y[0] - x12 - mp.mpf(0.4) * u * y[0], -y[1] + x12 - mp.mpf(0.2) * u * y[1], (y[0] - 1) ** 2 + (y[1] - 1) ** 2]

    x = [mp.mpf("0.5"), mp.mpf("0.7"), mp.mpf(0)]
    times, samples = [mp.mpf(0)], [list(x)]
    edges = [mp.mpf(float(t)) for t in np.r_[grid, 3.0]]
    for j in range(len(grid)):
        t0, t1, u = edges[j], edges[j + 1], mp.mpf(float(ctrl[j]))
        h = (t1 - t0) / K
        inside = [mp.mpf(float(s)) for s in saveat if float(t0) < s < float(t1)]
        for s in range(K):
            ks = [ff(x, u)]
            for st in range(1, 6):
                ks.append(ff([x[c] + h * sum(Amp[st][m] * ks[m][c] for m in range(st)) for c in range(3)], u))
            xn = [x[c] + h * sum(Amp[6][m] * ks[m][c] for m in range(6)) for c in range(3)]
            ks.append(ff(xn, u))
            tp, tn = t0 + s * h, (t1 if s == K - 1 else t0 + (s + 1) * h)
            for sv in inside:
                if tp < sv <= tn:
                    th = (sv - tp) / h
                    b = [th * (R[0][0] + th * (R[0][1] + th * (R[0][2] + th * R[0][3])))]
                    b += [th * th * (R[i][1] + th * (R[i][2] + th * R[i][3])) for i in range(1, 7)]
                    times.append(sv)
                    samples.append([x[c] + h * sum(b[m] * ks[m][c] for m in range(7)) for c in range(3)])
            x = xn
        times.append(t1)
        samples.append(list(x))
    ref_t = np.array([float(t) for t in times])
    ref_u = np.array([[float(v) for v in smp] for smp in samples]).T
    assert np.array_equal(r["t"], ref_t)
    assert np.max(np.abs(r["u"][:3] - ref_u) / np.maximum(1.0, np.abs(ref_u))) <= 1e-13
