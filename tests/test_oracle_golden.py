"""Pin the CPU oracle on every golden value the reference's own tests hold for the path.

All reference goldens come from ADAPTIVE Tsit5; a converged fixed-step run of the oracle must reproduce them
at the tolerance stated per test (the reference test's own atol where that is reachable).
"""
import math

import numpy as np
import pytest

import problems as P
from oracle import pyoracle as po


def test_G1_lotka_ms4_defects():
    """test/core/multiple_shooting.jl:54-72: 15 shooting defects, atol = 1e-10"""
    L = po.OracleLayer(P.lotka_spec())
    assert (L.I, L.nvars, L.ndef, L.nsamples) == (4, 137, 15, 121)
    assert list(L.block_structure()) == [0, 32, 67, 102, 137]
    G1 = np.array([1.375754960969482, 0.22357357513551113, 1.2417260108009396,
                   1.3757549609694817, 0.22357357513551046, 1.2417260108009387,
                   1.3757549609694821, 0.2235735751355138, 1.2417260108009425] + [0.0] * 6)
    r = L.eval(L.default_ps(), K=8)
    assert np.allclose(r["defk"], G1, atol=1e-10, rtol=0)
    # stage-major ordering: (u0, p) per stage
    exp_s = np.concatenate([np.r_[G1[3 * i:3 * i + 3], 0.0, 0.0] for i in range(3)])
    assert np.allclose(r["defs"], exp_s, atol=1e-10, rtol=0)
    t = r["t"]
    assert len(np.unique(t)) == len(t) == 121 and t[0] == 0.0 and t[-1] == 12.0


def test_G2_G3_lotka_ms_objective_and_cons():
    """test/examples/lotka_ms.jl:45-48,76-79"""
    L = po.OracleLayer(P.lotka_spec())
    r = L.eval(L.default_ps(), K=8)
    assert abs(r["last"][2] - 1.2417260108009376) < 1e-9       # reference atol 1e-4
    cons = np.array([1.3757549609694821, 0.2235735751355118, 1.24172601080094, 1.375754960969481,
                     0.22357357513551102, 1.2417260108009385, 1.3757549609694824, 0.2235735751355129,
                     1.2417260108009414])
    assert np.allclose(r["defk"][:9], cons, atol=1e-10, rtol=0)
    Q = po.OracleLayer(P.lotka_spec(quadrature=True))
    rq = Q.eval(Q.default_ps(), K=8)
    assert Q.ndef == 12
    assert abs(rq["last"][2] - 4 * 1.2417260108009376) < 1e-9
    consq = np.array([1.3757549609694821, 0.2235735751355118, 1.375754960969481, 0.22357357513551102,
                      1.3757549609694824, 0.2235735751355129] + [0.0] * 6)
    assert np.allclose(rq["defk"], consq, atol=1e-10, rtol=0)


def test_G3_lotka_ss_objective():
    """test/examples/lotka_oc.jl:80: 6.062277454291031 (reference atol 1e-4)"""
    L = po.OracleLayer(P.lotka_spec(shooting_points=None))
    assert L.nvars == 122 and L.I == 1
    r = L.eval(L.default_ps(), K=4)
    assert abs(r["last"][2] - 6.062277454291031) < 1e-9


def test_G4_G5_forward_and_hybrid_initialisation():
    """test/core/multiple_shooting.jl:131-152 (forward: zero defects) and :202-222 (hybrid golden, atol 1e-10)"""
    L = po.OracleLayer(P.lotka_spec())
    ps = L.forward_init(L.default_ps(), K=8)
    r = L.eval(ps, K=8)
    assert np.all(r["defk"] == 0.0)                                   # @test all(iszero, ...)
    # hybrid: interval 2 custom Dict(2 => 3.0, 3 => -5.0) on top of u0 -> [0.5, 3.0, -5.0], intervals 3,4 forward
    # with fixed_indices = [2, 4]; interval 4 keeps the custom value [1.0, 1.1, 1.34] -> but hybrid only takes
    # intervals [3, 4] from the forward result, where 4 is fixed to the default u0
    ps0 = L.default_ps()
    blocks = L.block_structure()
    ps_c = ps0.copy()
    ps_c[blocks[1]:blocks[1] + 3] = [0.5, 3.0, -5.0]
    ps_f = L.forward_init(ps_c, fixed_indices=[2, 4], K=8)
    node3 = ps_f[blocks[2]:blocks[2] + 3]
    assert np.allclose(node3, [0.2875960819428889, 0.27699894224421623, -1.088467272254529], atol=1e-10, rtol=0)
    assert np.array_equal(ps_f[blocks[3]:blocks[3] + 3], [0.5, 0.7, 0.0])


def test_G7_oed_layout_and_criteria():
    """lib/CorleoneOED/test/core/lotka_oed.jl:56-59,124"""
    L = po.OracleLayer(P.lotka_oed_spec())
    assert L.nvars == 2 + 48 + 24 + 79
    r = L.eval(L.default_ps(), K=8)
    assert np.array_equal(r["u"][:, 0], np.r_[[0.5, 0.7], np.zeros(4 + 3 + 1), np.ones(2)])
    F = r["last"][6:9]
    A = np.trace(np.linalg.inv(np.array([[F[0], F[1]], [F[1], F[2]]])))
    assert abs(A / 0.05352869250783344 - 1) < 1.5e-8            # reference: isapprox, rtol sqrt(eps)
    M = po.OracleLayer(P.lotka_oed_spec(shooting_points=[0.0, 3.0, 6.0, 9.0]))
    rm = M.eval(M.default_ps(), K=8)
    F = rm["last"][6:9]
    A = np.trace(np.linalg.inv(np.array([[F[0], F[1]], [F[1], F[2]]])))
    # this golden was produced at the solver's DEFAULT tolerances (reltol 1e-3); converged value differs by 7e-8
    assert abs(A / 0.6199466255548527 - 1) < 2e-7


def test_G8_control_symbol_permutation():
    """test/core/local_controls.jl:45-79: controls given in order [2,3,1]; sample rows stay in layer order"""
    s = P.egerstedt_spec()
    L = po.OracleLayer(s)
    r = L.eval(L.default_ps(), K=2)
    u = r["u"]
    assert np.all((0.3 <= u[3]) & (u[3] <= 0.5))     # row 4 = first control given = con2
    assert np.all((0.6 <= u[4]) & (u[4] <= 0.8))     # con3
    assert np.all((0.0 <= u[5]) & (u[5] <= 0.2))     # con1
    it = L.interval(0)
    assert np.array_equal(it["index_grid"][:, 0], [1, 21, 41]) and np.array_equal(it["index_grid"][:, -1], [20, 40, 60])


def test_G9_analytic_1d_oed():
    """lib/CorleoneOED/test/core/1d_oed.jl:11-31: x' = p x, x0 = 1, p = -2 => G = t e^{pt}, F = int t^2 e^{2pt}"""
    L = po.OracleLayer(P.lin1d_oed_spec())
    r = L.eval(L.default_ps(), K=2)
    assert np.allclose(r["t"], np.arange(101) / 100.0, atol=0, rtol=0)
    x, G, F = r["last"][:3]
    Fex = (1 - math.exp(-4) * (1 + 4 + 8)) / 32       # int_0^1 t^2 e^{-4t} dt
    assert abs(x - math.exp(-2)) < 1e-12 and abs(G - math.exp(-2)) < 1e-12 and abs(F - Fex) < 1e-12


def test_lqr_closed_form():
    """LQR: x(t+h) = e^{ah} x + b u (e^{ah} - 1)/a per piece; cost integral in closed form"""
    rng = np.random.default_rng(1)
    s = P.lqr_spec(10, rng)
    L = po.OracleLayer(s)
    ps = L.default_ps() + 0.1 * rng.standard_normal(L.nvars)
    r = L.eval(ps, K=8)
    blocks = L.block_structure()
    a, b = -1.0, 1.0
    for i in range(L.I):
        v = ps[blocks[i]:blocks[i + 1]]
        nu0 = 0 if i == 0 else 2
        x, c = (1.0, 0.0) if i == 0 else (v[0], v[1])
        a, b = v[nu0], v[nu0 + 1]
        for j in range(4):
            u, h = v[nu0 + 2 + j], 0.03
            # x(s) = e^{as}(x + bu/a) - bu/a; cost' = 10 (x - 3)^2 + 0.1 u^2
            A, Cc = x + b * u / a, -b * u / a - 3.0
            c += 10 * (A * A * (math.exp(2 * a * h) - 1) / (2 * a) + 2 * A * Cc * (math.exp(a * h) - 1) / a + Cc * Cc * h) \
                + 0.1 * u * u * h
            x = math.exp(a * h) * A - b * u / a
        assert abs(r["xend"][0, i] - x) < 1e-12 and abs(r["xend"][1, i] - c) < 1e-11


def test_sensitivities_match_finite_differences():
    rng = np.random.default_rng(5)
    s = P.lotka_spec(controls=rng.uniform(0, 1, 120))
    L = po.OracleLayer(s)
    ps = L.default_ps() + 0.05 * rng.standard_normal(L.nvars)
    Jx, Jd, Jl = L.jacobians(ps, K=2)
    for v in rng.choice(L.nvars, 12, replace=False):
        e = np.zeros(L.nvars)
        e[v] = 1e-6
        rp, rm = L.eval(ps + e, K=2), L.eval(ps - e, K=2)
        fd = (rp["defk"] - rm["defk"]) / 2e-6
        assert np.allclose(Jd[:, v], fd, atol=2e-8, rtol=1e-7)
        fdx = (rp["xend"] - rm["xend"]).T.ravel() / 2e-6
        assert np.allclose(Jx[:, v], fdx, atol=2e-8, rtol=1e-7)


def test_tsit5_order_conditions_and_convergence():
    """fixed-step Tsit5 converges with order 5, RK4 with order 4 (on the Lotka SS objective)"""
    L = po.OracleLayer(P.lotka_spec(shooting_points=None))
    ps = L.default_ps()
    ref = L.eval(ps, K=32)["last"]
    for method, order in (("tsit5", 5), ("rk4", 4)):
        e1 = np.abs(L.eval(ps, method=method, K=1)["last"][:3] - ref[:3]).max()
        e2 = np.abs(L.eval(ps, method=method, K=2)["last"][:3] - ref[:3]).max()
        assert order - 0.8 < math.log2(e1 / e2) < order + 1.5, (method, e1, e2)


def test_fp64_oracle_vs_extended_precision_restatement():
    """An independent restatement of the same fixed-step Tsit5 piece loop in 40-digit arithmetic (mpmath): the FP64
    oracle agrees to rounding (<= 1e-13 relative), so differences at the 1e-10 parity tolerance are never rounding."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    A = [[], ["0.161"], ["-0.008480655492356989", "0.335480655492357"],
         ["2.8971530571054935", "-6.359448489975075", "4.3622954328695815"],
         ["5.325864828439257", "-11.748883564062828", "7.4955393428898365", "-0.09249506636175525"],
         ["5.86145544294642", "-12.92096931784711", "8.159367898576159", "-0.071584973281401", "-0.028269050394068383"],
         ["0.09646076681806523", "0.01", "0.4798896504144996", "1.379008574103742", "-3.290069515436081", "2.324710524099774"]]
    A = [[mp.mpf(float(x)) for x in row] for row in A]   # the tableau as the FP64 constants both codes use

    rng = np.random.default_rng(5)
    ctrl = rng.uniform(0, 1, 120)
    spec = P.lotka_spec(controls=ctrl)
    L = po.OracleLayer(spec)
    K = 3
    r = L.eval(L.default_ps(), K=K)
    # interval 1: (0, 3), 30 pieces of 0.1 with controls ctrl[0:30], from u0 = [0.5, 0.7, 0]
    x = [mp.mpf("0.5"), mp.mpf("0.7"), mp.mpf(0)]
    grid = [mp.mpf(float(k / 10.0)) for k in range(31)]
    for j in range(30):
        h = (grid[j + 1] - grid[j]) / K
        u = mp.mpf(float(ctrl[j]))
        # 0.4 and 0.2 enter the FP64 code as FP64 constants
        def ff(y, u=u):
            x12 = y[0] * y[1]
            return [y[0] - x12 - mp.mpf(0.4) * u * y[0], -y[1] + x12 - mp.mpf(0.2) * u * y[1],
                    (y[0] - 1) ** 2 + (y[1] - 1) ** 2]
        for _ in range(K):
            ks = [ff(x)]
            for s in range(1, 6):
                y = [x[c] + h * sum(A[s][m] * ks[m][c] for m in range(s)) for c in range(3)]
                ks.append(ff(y))
            x = [x[c] + h * sum(A[6][m] * ks[m][c] for m in range(6)) for c in range(3)]
    ref = np.array([float(v) for v in x])
    got = r["xend"][:, 0]
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-13


@pytest.mark.parametrize("name", sorted(__import__("problems").BENCH_TABLE))
def test_benchmark_problem_derivatives_match_central_differences(name):
    """every right-hand side of lib/OptimalControlBenchmarks restated in the oracle: the dual-number derivatives of the
    last sample agree with central differences of the values (pins the hand-written tangent arithmetic)"""
    import problems as P
    from oracle import pyoracle as po
    spec, row, scale = P.bench_spec(name, 3 if name == "ductedfan" else 4)
    O = po.OracleLayer(spec)
    rng = np.random.default_rng(1)
    x = O.default_ps() * (1 + scale * rng.standard_normal(O.nvars)) + scale * rng.standard_normal(O.nvars)
    r = O.eval(x, K=3)
    assert np.isfinite(r["xend"]).all() and np.isfinite(r["defk"]).all()
    _, _, Jl = O.jacobians(x, K=3)
    for v in rng.choice(O.nvars, 4, replace=False):
        h = 1e-6 * max(1.0, abs(x[v]))
        xp, xm = x.copy(), x.copy()
        xp[v] += h
        xm[v] -= h
        fd = (O.eval(xp, K=3)["last"] - O.eval(xm, K=3)["last"]) / (2 * h)
        assert np.max(np.abs(fd - Jl[:, v])) <= 1e-6 * max(1.0, np.max(np.abs(Jl[:, v]))), (name, v)
