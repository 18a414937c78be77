"""Host-side mirror (corleone_b200) vs the oracle's independent restatement: every index table bit-exact,
plus the reference's own unit tests for ControlParameter / layer structure / initialisers (CPU only)."""
import numpy as np
import pytest

import corleone_b200 as cb
import problems as P
from corleone_b200 import controls as cc
from oracle import pyoracle as po


def _specs():
    rng = np.random.default_rng(0)
    return {
        "lotka_ms4": P.lotka_spec(),
        "lotka_ms4_quad": P.lotka_spec(quadrature=True),
        "lotka_ss_chunked": P.lotka_spec(shooting_points=None, tunable_ic=[2, 1]),
        "lotka_ms24": P.lotka_ms24_spec(rng),
        "lqr_ms100": P.lqr_spec(100, rng),
        "lotka_oed_ss": P.lotka_oed_spec(),
        "lotka_oed_ms4": P.lotka_oed_spec(shooting_points=[0.0, 3.0, 6.0, 9.0]),
        "oed_cfg3": P.config3_lotka_oed(8, rng),
        "lin1d_oed": P.lin1d_oed_spec(),
        "egerstedt": P.egerstedt_spec(),
        "dense20": P.dense20_spec(4, rng),
        # shooting nodes that do NOT coincide with control grid points -> control matchings
        "lotka_offgrid": P.lotka_spec(shooting_points=[0.0, 2.95, 6.0, 9.05]),
    }


@pytest.mark.parametrize("name", list(_specs().keys()))
def test_tables_bit_exact_vs_oracle(name):
    spec = _specs()[name]
    O = po.OracleLayer(spec)
    lay = P.product_layer(spec, cb.Tsit5(1))
    ps, st = cb.setup(None, lay)
    sts = [st] if "tspans" in st else list(st.values())
    pss = [ps] if "u0" in ps else list(ps.values())
    assert len(sts) == O.I
    assert np.array_equal(cb.get_block_structure(lay), O.block_structure())
    assert np.array_equal(cb.to_vec(ps), O.default_ps())
    from corleone_b200.single_shooting import _flatten_chunks
    for i, (s, p) in enumerate(zip(sts, pss)):
        it = O.interval(i)
        tsp = _flatten_chunks(s["tspans"])
        grid = np.array([tsp[0][0]] + [b for _, b in tsp])
        assert np.array_equal(grid, it["grid"]), (name, i)                       # bit-exact float grid
        ig = _flatten_chunks(s["index_grid"])
        if it["index_grid"].size:
            assert ig.dtype == np.int64 and np.array_equal(ig, it["index_grid"]), (name, i)
        assert np.array_equal(s["shooting_indices"], it["shooting_indices"]), (name, i)
        assert np.array_equal(s["initial_condition"].sorting, it["sorting"])
        assert np.array_equal(s["initial_condition"].constants, it["constants"])
        assert (len(p["u0"]), len(p["p"]), len(p["controls"])) == (it["n_u0"], it["n_p"], it["n_c"])
    if O.I > 1:
        assert cb.get_number_of_shooting_constraints(lay, ps, st) == O.ndef


def test_offgrid_nodes_give_control_matchings():
    spec = _specs()["lotka_offgrid"]
    lay = P.product_layer(spec, cb.Tsit5(1))
    ps, st = cb.setup(None, lay)
    assert list(st["interval_2"]["shooting_indices"]) == [1, 2, 3, 4]     # 2.95 is not a control grid point
    assert list(st["interval_3"]["shooting_indices"]) == [1, 2, 3]        # 6.0 is
    assert cb.get_number_of_control_matchings(lay, ps, st) == 2
    # the first piece of interval 2 is (2.95, 3.0) and uses local control 1 (clamped searchsortedlast = 0 -> 1)
    assert st["interval_2"]["tspans"][0] == (2.95, 3.0)
    assert list(st["interval_2"]["index_grid"][0, :3]) == [1, 1, 2]


def test_index_grid_function_vs_oracle_random_grids():
    rng = np.random.default_rng(42)
    for trial in range(200):
        nc = int(rng.integers(1, 4))
        ts = [np.unique(np.round(rng.uniform(0, 10, int(rng.integers(0, 12))), 1)) for _ in range(nc)]
        t0 = float(np.round(rng.uniform(0, 5), 1))
        t1 = t0 + float(np.round(rng.uniform(0.1, 5), 1))
        ctrls = [cb.ControlParameter(t, controls=np.zeros(len(t))) for t in ts]
        idx = cc.build_index_grid(*ctrls, tspan=(t0, t1))
        tsp = cc.collect_tspans(*ctrls, tspan=(t0, t1))
        oidx, ogrid = po.build_index_grid(ts, t0, t1)
        assert np.array_equal(idx, oidx), (trial, ts, t0, t1)
        assert np.array_equal(np.array([tsp[0][0]] + [b for _, b in tsp]), ogrid)
        for t in ts:
            assert np.array_equal(cc._restrict(t, (t0, t1)), po.restrict_grid(t, t0, t1))


def test_subvector_indices_and_chunking():
    for M, L in [(0, 5), (5, 5), (6, 5), (120, 100), (250, 100), (99, 100)]:
        assert cc.get_subvector_indices(M, L) == po.subvector_indices(M, L)
    with pytest.raises(ValueError):
        cc.get_subvector_indices(-1, 3)
    lay = P.product_layer(P.lotka_spec(shooting_points=None), cb.Tsit5(1))
    st = lay.initialstates(None)
    assert isinstance(st["index_grid"], tuple) and [g.shape for g in st["index_grid"]] == [(1, 100), (1, 20)]
    assert [len(c) for c in st["tspans"]] == [100, 20]      # Appendix B of SURVEY.md
    assert lay.parameterlength() == 120 + 2


# ---- the reference's own ControlParameter unit tests (test/core/local_controls.jl:9-42)
def test_control_parameter_reference_unit_tests():
    c = cb.ControlParameter(np.arange(101) / 100.0)
    lb, ub = cb.get_control_bounds(c)
    assert c.controls is cb.default_u and c.bounds is cb.default_bounds
    cb.check_consistency(None, c)
    assert np.unique(lb) == [-np.inf] and np.unique(ub) == [np.inf]
    c1 = cb.ControlParameter(np.arange(1.0, 11.0), bounds=(0.0, 1.0))
    lb1, ub1 = cb.get_control_bounds(c1)
    assert np.unique(lb1) == [0.0] and np.unique(ub1) == [1.0]
    cb.check_consistency(None, c1)
    c2 = cb.ControlParameter(np.arange(1.0, 11.0), bounds=(-np.ones(10), np.ones(10)))
    cb.check_consistency(None, c2)
    c3 = cb.ControlParameter(np.arange(1.0, 11.0), bounds=(-np.ones(10), np.ones(10)), controls=np.arange(10) / 10.0)
    assert np.array_equal(cb.get_controls(None, c3), np.arange(10) / 10.0)
    cb.check_consistency(None, c3)
    c4 = cb.ControlParameter(np.arange(1.0, 11.0), bounds=(-np.ones(11), np.ones(10)), controls=np.arange(10) / 10.0)
    with pytest.raises(ValueError, match="Incompatible control bound definition"):
        cb.check_consistency(None, c4)
    c5 = cb.ControlParameter(np.arange(1.0, 11.0), bounds=(-np.ones(10), np.ones(10)), controls=np.arange(11) / 10.0)
    with pytest.raises(AssertionError, match="Sizes are inconsistent"):
        cb.check_consistency(None, c5)
    c6 = cb.ControlParameter(np.arange(1.0, 11.0), bounds=(np.ones(10), -np.ones(10)), controls=np.arange(11) / 10.0)
    with pytest.raises(AssertionError, match="Bounds are inconsistent"):
        cb.check_consistency(None, c6)


def test_layer_structure_reference_unit_tests():
    """test/core/multiple_shooting.jl:43-51, test/examples/lotka_oc.jl:68-73"""
    lay = P.product_layer(P.lotka_spec(), cb.Tsit5(1))
    ps, st = cb.setup(None, lay)
    assert len(ps) == 4 and len(ps["interval_1"]["u0"]) == 0
    assert all(len(ps[f"interval_{i}"]["u0"]) == 3 for i in (2, 3, 4))
    blocks = np.cumsum([sum(len(v) for v in p.values()) for p in ps.values()])
    assert np.array_equal(cb.get_block_structure(lay), np.r_[0, blocks])
    assert lay.parameterlength() == 137 == len(cb.to_vec(ps))
    lb, ub = cb.get_bounds(lay)
    assert np.all(np.isinf(lb["interval_2"]["u0"]))
    ss = P.product_layer(P.lotka_spec(shooting_points=None), cb.Tsit5(1),
                         bounds_p=([1.0, 1.0], [1.0, 1.0]))
    lb, ub = ss.get_bounds()
    assert np.array_equal(lb["p"], [1.0, 1.0]) and np.array_equal(ub["p"], [1.0, 1.0])
    back = cb.from_vec(cb.to_vec(ps), ps)
    assert all(np.array_equal(back[k][f], ps[k][f]) for k in ps for f in ("u0", "p", "controls"))


def test_cheap_initialisers_reference_unit_tests():
    """test/core/multiple_shooting.jl:153-201 (constant, linear, custom: exact values)"""
    u0 = np.array([0.5, 0.7, 0.0])
    mk = lambda init: P.product_layer(P.lotka_spec(), cb.Tsit5(1), initialization=init)
    ps = mk(lambda *a: cb.constant_initialization(*a, u0=[0.9, 1.1, 3.0])).initialparameters(None)
    assert len(ps["interval_1"]["u0"]) == 0
    assert all(np.array_equal(ps[f"interval_{i}"]["u0"], [0.9, 1.1, 3.0]) for i in (2, 3, 4))
    ps = mk(lambda *a: cb.linear_initialization(*a, u_infinity=[2.0, 1.0, 1.34])).initialparameters(None)
    for i, t in ((2, 3), (3, 6), (4, 9)):
        assert np.allclose(ps[f"interval_{i}"]["u0"], u0 + (np.array([2.0, 1.0, 1.34]) - u0) * t / 12, atol=1e-7)
    ps = mk(lambda *a: cb.custom_initialization(*a, u0s=[u0] + [[1.0, 1.1, 1.34]] * 3)).initialparameters(None)
    assert all(np.array_equal(ps[f"interval_{i}"]["u0"], [1.0, 1.1, 1.34]) for i in (2, 3, 4))
    rng = np.random.default_rng(0)
    lay = P.product_layer(P.lotka_spec(), cb.Tsit5(1), bounds_ic=([0.1, 0.1, 0.0], [100.0, 100.0, 100.0]),
                          initialization=cb.random_initialization)
    ps = lay.initialparameters(rng)
    assert all(np.all((ps[f"interval_{i}"]["u0"] >= [0.1, 0.1, 0.0]) & (ps[f"interval_{i}"]["u0"] <= 100.0)) for i in (2, 3, 4))


def test_symbol_cache_permutation():
    """test/core/local_controls.jl:45-79: names follow the sample rows"""
    lay = P.product_layer(P.egerstedt_spec(), cb.Tsit5(1))
    st = lay.initialstates(None)
    v = st["symcache"].variables
    assert (v["con2"], v["con3"], v["con1"]) == (3, 4, 5) and v["x₁"] == 0
    assert st["symcache"].parameters == {}
