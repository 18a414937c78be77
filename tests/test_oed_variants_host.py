"""Host tables of the OED layer variants (lib/CorleoneOED/src/oed.jl) against the oracle's literal restatement, and
the oracle's own pins for the variants that have no golden value in the reference's tests (CPU only)."""
import numpy as np
import pytest

import corleone_b200 as cb
from corleone_b200 import oed as oedm
from oracle import pyoracle

CG = 0.25 * np.arange(48)
T1 = 0.5 * np.arange(24)
T2 = 3.0 * np.arange(79) / 20.0


def lotka_base(shooting=None, fixed=False, K=4):
    prob = cb.ODEProblem("lotka2", [0.5, 0.7], (0.0, 12.0), [0.0, 1.0, 1.0])
    ctr = [] if fixed else [(1, cb.ControlParameter(CG, name="fishing", bounds=(0.0, 1.0)))]
    bp = ([0.0, 1.0, 1.0],) * 2 if fixed else ([1.0, 1.0],) * 2
    if shooting is None:
        return cb.SingleShootingLayer(prob, cb.Tsit5(K), controls=ctr, bounds_p=bp)
    return cb.MultipleShootingLayer(prob, cb.Tsit5(K), *shooting, controls=ctr, bounds_p=bp)


def measurements(w=1.0):
    return [cb.ControlParameter(T1, controls=np.full(len(T1), w), bounds=(0.0, 1.0)),
            cb.ControlParameter(T2, controls=np.full(len(T2), w), bounds=(0.0, 1.0))]


def test_variant_flags_and_models():
    """OEDLayer{DISCRETE, SAMPLED, FIXED} (oed.jl:87-91,128-132) -> compiled-in augmentation"""
    cases = [(False, True, False, "lotka2_oed", 0), (False, False, False, "lotka2_oed_cu", 0),
             (False, True, True, "lotka2_oed_cf", 3), (True, True, False, "lotka2_oed_ds", 2),
             (True, False, False, "lotka2_oed_du", 1), (True, True, True, "lotka2_oed_ds", 2)]
    for discrete, sampled, fixed, model, mode in cases:
        o = oedm.OEDLayer(lotka_base(fixed=fixed), discrete, measurements() if sampled else ())
        assert (o.DISCRETE, o.SAMPLED, o.FIXED, o.model, o.fim_mode) == (discrete, sampled, fixed, model, mode)
        assert o.n_observed() == (2 if sampled else 0)
    with pytest.raises(NotImplementedError):
        oedm.OEDLayer(lotka_base(shooting=(0.0, 3.0, 6.0, 9.0), fixed=True), False, measurements())


def test_initial_state_and_bounds_layout():
    """lib/CorleoneOED/test/core/lotka_oed.jl:56,58: sol.u[1] == [u0; zeros(4 + 3); ...] and the variable layout"""
    o = oedm.OEDLayer(lotka_base(), False, measurements())
    ps, st = o.setup(None)
    assert np.array_equal(o.layer.problem.u0, np.r_[0.5, 0.7, np.zeros(7)])
    assert np.array_equal(o.layer.problem.p, [0.0, 1.0, 1.0, 1.0, 1.0])
    lb, ub = cb.get_bounds(o.layer)
    assert np.array_equal(cb.to_vec(lb), np.r_[np.ones(2), np.zeros(len(CG) + len(T1) + len(T2))])
    assert np.array_equal(st["F_init"], np.zeros((2, 2)))
    assert st["_fim"]["offset"] == 7 and st["_fim"]["mode"] == 0


@pytest.mark.parametrize("shooting", [None, (0.0, 3.0, 6.0, 9.0)])
def test_weighting_grid_matches_the_oracle_restatement(shooting):
    """WeightedObservation.grid: numpy builder of the host mirror == loop restatement of oed.jl:291-311 / :341-357"""
    o = oedm.OEDLayer(lotka_base(shooting), True, measurements())
    ps, st = o.setup(None)
    grids = [CG, T1, T2]
    sts = [st] if shooting is None else list(st.values())
    for sti in sts:
        tsp = oedm._flatten_chunks(sti["tspans"])
        tspan = (0.0, 12.0) if shooting is None else (tsp[0][0], tsp[-1][1])
        ig = pyoracle.build_index_grid(grids, *tspan)[0]
        ref = pyoracle.weighting_grid(grids, [2, 3], ig, tspan, shooting is not None)
        assert np.array_equal(sti["observation_grid"], np.asarray(ref, dtype=np.int64))
        # every row belongs to a saved sample of the interval: its start, the measurement times inside its pieces (saveat)
        # and the piece ends (+ the final time, single shooting): the union of the control and the measurement grids
        n_inside = sum(1 for a, b in tsp for t in np.union1d(T1, T2) if a < t < b)
        assert len(ref) == len(tsp) + n_inside + (1 if shooting is None else 0)
    # the final time carries no measurement (t0 <= t < t1, src/local_controls.jl:53-57)
    if shooting is None:
        assert not st["observation_grid"][-1].any()
        # measurement k acts at exactly its own grid points, in order
        for k, g in enumerate((T1, T2)):
            rows = np.nonzero(st["observation_grid"][:, k])[0]
            assert len(rows) == len(g)
            assert np.array_equal(st["observation_grid"][rows, k], len(CG) + (len(T1) if k else 0) + 1 + np.arange(len(g)))


def test_sampling_sums():
    """get_sampling_sums (lotka_oed.jl:61-64: [12, 12] for the continuous layer; the discrete layer counts measurements,
    oed.jl:573-580; the fixed continuous one integrates w dt over its own pieces, :582-589)"""
    for shooting in (None, (0.0, 3.0, 6.0, 9.0)):
        o = oedm.OEDLayer(lotka_base(shooting), False, measurements())
        ps, st = o.setup(None)
        assert np.allclose(o.get_sampling_sums(ps, st), [12.0, 12.0], rtol=0, atol=1e-12)
        o = oedm.OEDLayer(lotka_base(shooting), True, measurements())
        ps, st = o.setup(None)
        assert np.array_equal(o.get_sampling_sums(ps, st), [float(len(T1)), float(len(T2))])
    assert len(oedm.OEDLayer(lotka_base(), True).get_sampling_sums(None, None)) == 0
    m = [cb.ControlParameter(T1, controls=np.ones(len(T1)), bounds=(0.0, 1.0)) for _ in range(2)]
    o = oedm.OEDLayer(lotka_base(fixed=True), False, m)
    ps, st = o.setup(None)
    assert np.allclose(o.get_sampling_sums(ps, st), [12.0, 12.0], rtol=0, atol=1e-12)


def _oracle_spec(o, ps):
    lay = o.layer if o.single else o.layer.layer
    ctr = [(ci, c.t, np.asarray(c.controls, dtype=np.float64) if not callable(c.controls) else None)
           for ci, c in zip(lay.control_indices, lay.controls)]
    return dict(model=o.model, u0=lay.problem.u0, p0=lay.problem.p, tspan=lay.problem.tspan, controls=ctr,
                tunable_ic=lay.tunable_ic, quadrature_indices=[], saveat=lay.problem.kwargs.get("saveat"),
                shooting_points=None if o.single else [t[0] for t in o.layer.shooting_intervals])


def test_discrete_information_of_the_1d_problem_is_the_closed_form():
    """lib/CorleoneOED/test/core/1d_oed.jl:11-31: x' = p x, x0 = 1, p = -2 => G(t) = t e^{pt}.  The discrete layer sums
    (w_i G(t_i))^2 over the measurement times (oed.jl:2-16,409-415) -- pins the oracle's read-out and augmentation."""
    t = np.arange(100) / 100.0
    rng = np.random.default_rng(5)
    w = rng.uniform(0, 1, 100)
    prob = cb.ODEProblem("lin1d", [1.0], (0.0, 1.0), [-2.0])
    # the layer has no control, so the horizon is ONE piece and all 99 inner measurement times are saveat samples from the
    # dense output of the 256 steps (4th-order interpolant: 3e-15 at this step size, 1e-5 at K = 4)
    o = oedm.OEDLayer(cb.SingleShootingLayer(prob, cb.Tsit5(256)), True, [cb.ControlParameter(t, controls=w, bounds=(0.0, 1.0))])
    ps, st = o.setup(None)
    assert o.FIXED and o.model == "lin1d_oed_ds"
    assert np.array_equal(o.layer.problem.kwargs["saveat"], t) and np.array_equal(o.layer.problem.p, [-2.0])
    orc = pyoracle.OracleLayer(_oracle_spec(o, ps))
    x = cb.to_vec(ps)
    r = orc.eval(x, K=256)
    assert np.array_equal(r["t"], np.r_[t, 1.0])
    F = pyoracle.fisher_discrete_sampled(ps["controls"], st["observation_grid"].tolist(), r["u"], 0, 1, 1, 1)
    exact = np.sum((w * t * np.exp(-2.0 * t)) ** 2)
    assert abs(F[0, 0] - exact) <= 1e-12 * exact
    # the fixed continuous layer integrates the same sensitivities: F = sum_j w_j int_{t_j}^{t_j+1} t^2 e^{2pt} dt
    o2 = oedm.OEDLayer(cb.SingleShootingLayer(prob, cb.Tsit5(4)), False, [cb.ControlParameter(t, controls=w, bounds=(0.0, 1.0))])
    ps2, st2 = o2.setup(None)
    assert o2.model == "lin1d_oed_cf" and o2.fim_mode == 3
    r2 = pyoracle.OracleLayer(_oracle_spec(o2, ps2)).eval(cb.to_vec(ps2), K=4)
    F2 = pyoracle.fisher_fixed_continuous(ps2["controls"], st2["active_controls"], r2["u"], 1, 1, 1)
    prim = lambda s: np.exp(-4.0 * s) * (-(s * s) / 4.0 - s / 8.0 - 1.0 / 32.0)     # antiderivative of s^2 e^{-4 s}
    edges = np.r_[t, 1.0]
    exact2 = np.sum(w * (prim(edges[1:]) - prim(edges[:-1])))
    assert abs(F2[0, 0] - exact2) <= 1e-11 * exact2


def test_saveat_samples_on_step_ends_equal_the_cut_pieces():
    """The discrete layers sample inside a piece (`saveat`, oed.jl:106-113) instead of cutting it.  When the measurement
    times fall on step ends the dense output at theta = 1 is the step's own result: the samples equal those of a layer
    whose control grid cuts the pieces there (same steps: k1 re-evaluated at a cut equals the FSAL k7)."""
    prob = cb.ODEProblem("lotka2", [0.5, 0.7], (0.0, 12.0), [0.0, 1.0, 1.0])
    vals = np.linspace(0.1, 0.9, 12)
    fishing = cb.ControlParameter(1.0 * np.arange(12), name="fishing", controls=vals, bounds=(0.0, 1.0))
    fine = cb.ControlParameter(0.25 * np.arange(48), name="fishing", controls=np.repeat(vals, 4), bounds=(0.0, 1.0))
    meas = [cb.ControlParameter(0.25 * np.arange(48), controls=np.ones(48), bounds=(0.0, 1.0)) for _ in range(2)]
    sampled = oedm.OEDLayer(cb.SingleShootingLayer(prob, cb.Tsit5(8), controls=[(1, fishing)]), True, meas)
    cut = oedm.OEDLayer(cb.SingleShootingLayer(prob, cb.Tsit5(2), controls=[(1, fine)]), True)
    ps_s, st_s = sampled.setup(None)
    ps_c, _ = cut.setup(None)
    assert len(oedm._flatten_chunks(st_s["tspans"])) == 12          # the measurement grids do not cut the pieces
    a = pyoracle.OracleLayer(_oracle_spec(sampled, ps_s)).eval(cb.to_vec(ps_s), K=8)
    b = pyoracle.OracleLayer(_oracle_spec(cut, ps_c)).eval(cb.to_vec(ps_c), K=2)
    assert a["u"].shape[1] == b["u"].shape[1] == 49 and np.array_equal(a["t"], b["t"])
    assert np.max(np.abs(a["u"][:6] - b["u"][:6])) <= 1e-14
    assert np.array_equal(a["u"][:6, ::4], b["u"][:6, ::4])          # the piece ends themselves: bit for bit


def test_multi_experiment_layer_structure():
    """lib/CorleoneOED/test/core/lotka_oed.jl:154-247 (the variant with one parameter set for all experiments): grouping
    by experiment, block structure, bounds, sampling sums per experiment"""
    for shooting in (None, (0.0, 3.0, 6.0, 9.0)):
        for discrete in (False, True):
            m = cb.MultiExperimentLayer(oedm.OEDLayer(lotka_base(shooting), discrete, measurements()), 2)
            ps, st = m.setup(None)
            assert list(ps) == list(st) == ["experiment_1", "experiment_2"]
            nv = len(cb.to_vec(ps["experiment_1"]))
            assert len(cb.to_vec(ps)) == 2 * nv
            b1 = cb.get_block_structure(m.layers.layer)
            assert np.array_equal(m.get_block_structure(), np.r_[b1, b1[1:] + b1[-1]])
            lo, hi = m.get_bounds()
            assert np.array_equal(cb.to_vec(lo), np.tile(cb.to_vec(cb.get_bounds(m.layers.layer)[0]), 2))
            sums = m.get_sampling_sums(ps, st)
            assert np.allclose(sums, [24.0, 79.0, 24.0, 79.0] if discrete else [12.0] * 4, rtol=0, atol=1e-12)   # :230
            assert m.n_observed() == 4
            # per experiment: 3 nodes x (states + 2 tunable parameters), states = 9 ([x; G; F]) or 6 ([x; G])
            assert m.get_number_of_shooting_constraints(ps, st) == (0 if shooting is None else 2 * 3 * ((6 if discrete else 9) + 2))


@pytest.mark.parametrize("discrete", [False, True])
@pytest.mark.parametrize("fixed", [False, True])
def test_multiexperiments_case_1_of_the_reference_test(discrete, fixed):
    """lib/CorleoneOED/test/core/lotka_oed.jl:154-247, the branch `split = false, shooting = false`: two experiments over
    the same parameter set; sampling sums, bounds and block structure exactly as the reference asserts them (:228-240)"""
    num_exp = 2
    g = 0.25 * np.arange(48)
    meas = [cb.ControlParameter(g, controls=np.ones(48), bounds=(0.0, 1.0)) for _ in range(2)]
    multi = cb.MultiExperimentLayer(oedm.OEDLayer(lotka_base(None, fixed=fixed), discrete, meas), num_exp)
    ps, st = multi.setup(None)
    res = multi.get_sampling_sums(ps, st)
    assert np.array_equal(res, [48.0] * 4) if discrete else np.allclose(res, [12.0] * 4, rtol=0, atol=1e-12)
    lb, ub = multi.get_bounds()
    head = [0.0] if fixed else []
    assert np.array_equal(cb.to_vec(lb), np.tile(np.r_[head, np.ones(2), np.zeros(2 * 48 if fixed else 3 * 48)], num_exp))
    assert np.array_equal(cb.to_vec(ub), np.tile(np.r_[head, np.ones(2), np.ones(2 * 48 if fixed else 3 * 48)], num_exp))
    blocks = multi.get_block_structure()
    per = 3 + 48 * 2 if fixed else 2 + 48 * 3
    assert np.array_equal(blocks, np.r_[0, np.cumsum([per] * num_exp)])
