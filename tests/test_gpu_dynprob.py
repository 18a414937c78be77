"""The NLP assembly of src/dynprob.jl on top of the device path: the reference's example tests
(test/examples/lotka_oc.jl, test/examples/lotka_ms.jl) with SciPy's SLSQP standing in for Ipopt."""
import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu


def _lotka_layer(cb, shooting_points, K, **kw):
    spec = P.lotka_spec(shooting_points=shooting_points)
    spec["control_names"] = ["fishing"]
    spec["control_bounds"] = [(0.0, 1.0)]
    return P.product_layer(spec, cb.Tsit5(K), bounds_p=([1.0, 1.0], [1.0, 1.0]), **kw)


def test_lotka_ms_callbacks_reference_goldens():
    """test/examples/lotka_ms.jl:45-48: objective and constraint vector at the initial point"""
    import corleone_b200 as cb
    lay = _lotka_layer(cb, [0.0, 3.0, 6.0, 9.0], 8, bounds_ic=([0.1, 0.1, 0.0], [100.0, 100.0, 100.0]))
    prob = cb.CorleoneDynamicOptProblem(lay, "x₃")
    assert len(prob.u0) == lay.parameterlength() == 137
    assert abs(prob.objective(prob.u0) - 1.2417260108009376) < 1e-9          # reference atol 1e-4
    gold = np.array([1.3757549609694821, 0.2235735751355118, 1.24172601080094, 1.375754960969481, 0.22357357513551102,
                     1.2417260108009385, 1.3757549609694824, 0.2235735751355129, 1.2417260108009414] + [0.0] * 6)
    c = prob.constraints(prob.u0)
    assert c.shape == (15,) and np.allclose(c, gold, atol=1e-10, rtol=0)
    assert np.array_equal(prob.lcons, np.zeros(15)) and np.array_equal(prob.ucons, np.zeros(15))
    # bounds as the reference test checks them
    assert np.all(prob.lb[-30:] == 0.0) and np.all(prob.ub[-30:] == 1.0)


def test_lotka_single_shooting_optimum():
    """test/examples/lotka_oc.jl:80-94: f(u0) = 6.062277454291031, optimum 1.344336 (atol 1e-4)"""
    import corleone_b200 as cb
    lay = _lotka_layer(cb, None, 2)
    prob = cb.CorleoneDynamicOptProblem(lay, "x₃")
    assert abs(prob.objective(prob.u0) - 6.062277454291031) < 1e-6
    res = cb.solve_slsqp(prob, maxiter=400, ftol=1e-12)
    assert res.success, res.message
    assert abs(res.fun - 1.344336) < 1e-4
    assert np.all(res.x[:2] == 1.0)                                           # p_opt.p == p0[2:end]
    assert np.all((res.x[2:] >= -1e-12) & (res.x[2:] <= 1 + 1e-12))


def test_lotka_multiple_shooting_optimum():
    """test/examples/lotka_ms.jl:50-56: the same optimum through 15 matching constraints"""
    import corleone_b200 as cb
    lay = _lotka_layer(cb, [0.0, 3.0, 6.0, 9.0], 2, bounds_ic=([0.1, 0.1, 0.0], [100.0, 100.0, 100.0]))
    prob = cb.CorleoneDynamicOptProblem(lay, "x₃")
    res = cb.solve_slsqp(prob, maxiter=600, ftol=1e-12)
    # the parameter matchings p_i - p_{i+1} = 0 are redundant with lb.p == ub.p: SLSQP may stop at the optimum with
    # "positive directional derivative" (status 8) on that degenerate set instead of status 0
    assert res.status in (0, 8), res.message
    assert abs(res.fun - 1.344336) < 1e-4
    assert np.max(np.abs(prob.constraints(res.x))) < 1e-5       # the reference solves with Ipopt tol = 1e-3


def test_point_constraints_and_their_jacobian():
    """point constraints (src/dynprob.jl:89-103) stacked before the shooting constraints; Jacobian vs central differences"""
    import corleone_b200 as cb
    rng = np.random.default_rng(61)
    spec = P.lotka_spec(shooting_points=[0.0, 3.0, 6.0, 9.0], controls=rng.uniform(0, 1, 120), quadrature=True)
    spec["control_names"] = ["fishing"]
    lay = P.product_layer(spec, cb.Tsit5(2))
    tcon = [1.0, 3.0, 4.5, 6.0, 11.9]
    prob = cb.CorleoneDynamicOptProblem(lay, "x₃", ("x₁", dict(t=tcon, bounds=(0.4, np.inf))),
                                        ("x₃", dict(t=[9.0, 12.0], bounds=(0.0, 5.0))))
    x = prob.u0 + 0.05 * rng.standard_normal(len(prob.u0))
    f, g, c, J = prob.value_and_derivatives(x)
    assert prob.n_point == len(c) - prob.n_shoot and J.shape == (len(c), len(x))
    assert np.all(prob.lcons[:5] == 0.4) and np.all(np.isinf(prob.ucons[:5]))
    h = 1e-6
    for v in rng.choice(len(x), 10, replace=False):
        xp, xm = x.copy(), x.copy()
        xp[v] += h
        xm[v] -= h
        fd_c = (prob.constraints(xp) - prob.constraints(xm)) / (2 * h)
        fd_f = (prob.objective(xp) - prob.objective(xm)) / (2 * h)
        assert np.allclose(J[:, v], fd_c, rtol=1e-6, atol=2e-7), v
        assert abs(g[v] - fd_f) < 2e-7 * max(1.0, abs(fd_f))
    # batch axis: the same numbers for several candidates at once
    X = x[None, :] + 0.01 * rng.standard_normal((5, len(x)))
    fo, co, go = prob.evaluate_batch(X)
    assert abs(fo[2] - prob.objective(X[2])) < 1e-13 and np.allclose(co[4], prob.constraints(X[4]), rtol=0, atol=1e-13)


def test_lotka_oed_design_optimum():
    """lib/CorleoneOED/test/core/lotka_oed.jl:105-145 (single shooting): sampling sums [12, 12] at the start, A-criterion
    0.05352869250783344; optimum under sampling sums <= 4: A = 0.03707508955468313 and the golden Fisher matrix"""
    import corleone_b200 as cb
    spec = P.lotka_oed_spec()
    spec["control_bounds"] = [(0.0, 1.0)] * 3
    lay = P.product_layer(spec, cb.Tsit5(4), bounds_p=([1.0, 1.0], [1.0, 1.0]))
    ps_t, st = cb.setup(None, lay)
    st["_fim"] = (7, 2)
    N = lay._native(st, ps_t)
    x0 = cb.to_vec(ps_t)
    A = cb.sampling_sum_matrix(lay, st, ps_t, [1, 2])
    assert np.allclose(A @ x0, [12.0, 12.0], atol=1e-12)                       # lotka_oed.jl:61-64
    lo, hi = cb.get_bounds(lay)
    assert np.array_equal(cb.to_vec(lo), np.r_[np.ones(2), np.zeros(48 + 24 + 79)])     # :58-59
    assert np.array_equal(cb.to_vec(hi), np.ones(2 + 48 + 24 + 79))
    res = cb.solve_oed_slsqp(lay, st, ps_t, N, criterion=0, M=[4.0, 4.0], measurement_rows=(1, 2), maxiter=600)
    assert res.success and abs(res.fun / 0.03707508955468313 - 1) < 1e-5, (res.fun, res.message)   # measured: -1.3e-7
    Fgold = np.array([[38.58695777364362, 5.304118558316029], [5.304118558316029, 92.03122059902915]])
    assert np.allclose(res.fim, Fgold, rtol=1e-5), res.fim                                        # measured: 1e-7
    assert np.all(A @ res.x <= 4.0 + 1e-8)


def test_reference_example_callbacks_in_the_default_adaptive_mode():
    """test/examples/lotka_ms.jl:45-48,76-79 and lotka_oc.jl:80 with the integration mode the reference runs them in
    (adaptive Tsit5, abstol 1e-8 / reltol 1e-6): optprob.f and optprob.f.cons at the initial point to all printed digits,
    and the optimum of the single-shooting example"""
    import corleone_b200 as cb

    def layer(shooting_points, **kw):
        spec = dict(P.lotka_spec(shooting_points=shooting_points), kwargs=dict(abstol=1.0e-8, reltol=1.0e-6))
        spec["control_names"] = ["fishing"]
        spec["control_bounds"] = [(0.0, 1.0)]
        return P.product_layer(spec, cb.Tsit5(adaptive=True), bounds_p=([1.0, 1.0], [1.0, 1.0]), **kw)

    bic = ([0.1, 0.1, 0.0], [100.0, 100.0, 100.0])
    prob = cb.CorleoneDynamicOptProblem(layer([0.0, 3.0, 6.0, 9.0], bounds_ic=bic), "x₃")
    assert abs(prob.objective(prob.u0) - 1.2417260108009376) <= 2e-14
    gold = np.array([1.3757549609694821, 0.2235735751355118, 1.24172601080094, 1.375754960969481, 0.22357357513551102,
                     1.2417260108009385, 1.3757549609694824, 0.2235735751355129, 1.2417260108009414] + [0.0] * 6)
    assert np.max(np.abs(prob.constraints(prob.u0) - gold)) <= 2e-14
    probq = cb.CorleoneDynamicOptProblem(layer([0.0, 3.0, 6.0, 9.0], bounds_ic=bic, quadrature_indices=[3]), "x₃")
    assert abs(probq.objective(probq.u0) - 4 * 1.2417260108009376) <= 5e-14
    goldq = np.array([1.3757549609694821, 0.2235735751355118, 1.375754960969481, 0.22357357513551102,
                      1.3757549609694824, 0.2235735751355129] + [0.0] * 6)
    assert np.max(np.abs(probq.constraints(probq.u0) - goldq)) <= 2e-14
    ss = cb.CorleoneDynamicOptProblem(layer(None), "x₃")
    assert abs(ss.objective(ss.u0) - 6.062277454291031) <= 5e-14
    res = cb.solve_slsqp(ss, maxiter=400, ftol=1e-12)
    assert res.success and abs(res.fun - 1.344336) < 1e-4


def test_lotka_oed_design_optimum_in_the_default_adaptive_mode():
    """the same design problem integrated as the reference integrates it (adaptive Tsit5 at the default tolerances):
    the initial A-criterion to 1e-13 and the optimum `_uopt.objective ≈ 0.03707508955468313` (lotka_oed.jl:124,140-145)"""
    import corleone_b200 as cb
    spec = P.lotka_oed_spec()
    spec["control_bounds"] = [(0.0, 1.0)] * 3
    lay = P.product_layer(spec, cb.Tsit5(adaptive=True), bounds_p=([1.0, 1.0], [1.0, 1.0]))
    ps_t, st = cb.setup(None, lay)
    st["_fim"] = (7, 2)
    N = lay._native(st, ps_t)
    crit, _ = cb.batched_criteria(N, cb.to_vec(ps_t)[None, :])
    assert abs(crit[0, 0] / 0.05352869250783344 - 1) <= 1e-13
    A = cb.sampling_sum_matrix(lay, st, ps_t, [1, 2])
    res = cb.solve_oed_slsqp(lay, st, ps_t, N, criterion=0, M=[4.0, 4.0], measurement_rows=(1, 2), maxiter=600)
    dev = res.fun / 0.03707508955468313 - 1
    print("adaptive design optimum: relative deviation from the reference's", dev)
    assert abs(dev) < 1e-6, (res.fun, res.message)
    Fgold = np.array([[38.58695777364362, 5.304118558316029], [5.304118558316029, 92.03122059902915]])
    assert np.allclose(res.fim, Fgold, rtol=1e-5), res.fim
    assert np.all(A @ res.x <= 4.0 + 1e-8)


def test_a_failed_integration_is_reported_not_handed_to_the_optimiser():
    """the reference never inspects the solver's return code: a diverging candidate would reach the NLP solver as NaN
    objective / constraints / Jacobian.  The host mirror checks the device's status flags and refuses."""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    lay = _lotka_layer(cb, [0.0, 3.0, 6.0, 9.0], 8, bounds_ic=([0.1, 0.1, 0.0], [100.0, 100.0, 100.0]))
    prob = cb.CorleoneDynamicOptProblem(lay, "x₃")
    assert np.isfinite(prob.objective(prob.u0))
    bad = prob.u0.copy()
    bad[3] = 1e200                      # a node value: x' ~ -x y overflows at once
    with pytest.raises(nat.IntegrationFailure, match="candidate"):
        prob.objective(bad)
    with pytest.raises(nat.IntegrationFailure):
        prob.constraints(bad)
    # the OED read-out checks the same flags
    from corleone_b200 import oed as oedm
    base = cb.SingleShootingLayer(cb.ODEProblem("lotka2", [0.5, 0.7], (0.0, 12.0), [0.0, 1.0, 1.0]), cb.Tsit5(4),
                                  controls=[(1, cb.ControlParameter(0.25 * np.arange(48), name="fishing", bounds=(0.0, 1.0)))],
                                  bounds_p=([1.0, 1.0],) * 2)
    o = oedm.OEDLayer(base, False)
    ps, st = o.setup(None)
    X = np.tile(cb.to_vec(ps), (3, 1))
    assert np.isfinite(o.fisher_information(X, ps, st)).all()
    X[1, 0] = np.nan
    with pytest.raises(nat.IntegrationFailure) as e:
        o.fisher_information(X, ps, st)
    assert list(e.value.candidates) == [1]
