"""Interval sharding (SURVEY section 8e, B < #GPUs): cb200_set_interval_range restricts an evaluation to a contiguous
range of shooting intervals; the outputs of disjoint ranges assemble by summation -- bit-identical to the full
evaluation, because every number is produced by the same thread arithmetic either way."""
import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu


def _setup(spec, alg):
    import corleone_b200 as cb
    lay = P.product_layer(spec, alg)
    ps_t, st = cb.setup(None, lay)
    return lay, ps_t, st, lay._native(st, ps_t)


def _assemble(N, X, want, row, ranges):
    tot = None
    for first, last in ranges:
        N.set_interval_range(first, last)
        r = N.eval_host(X, want, objective_row=row)
        r = {k: v for k, v in r.items() if isinstance(v, np.ndarray)}
        tot = r if tot is None else {k: tot[k] + r[k] for k in tot}
    N.set_interval_range(1, N.I)
    return tot


@pytest.mark.parametrize("B", [1, 3, 40])
def test_ranges_sum_to_the_full_evaluation_lotka(B):
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(7)
    spec = P.lotka_spec(shooting_points=[0.0, 2.0, 4.0, 6.0, 8.0, 10.0], controls=rng.uniform(0, 1, 120), quadrature=True)
    lay, ps_t, st, N = _setup(spec, cb.Tsit5(2))
    X = cb.to_vec(ps_t)[None, :] + 0.05 * rng.standard_normal((B, N.nvars))
    want = (nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC
            | nat.WANT_GRAD_COMPACT | nat.WANT_TRAJ_SENS)
    full = N.eval_host(X, want, objective_row=3)
    for ranges in ([(1, 3), (4, 6)], [(1, 1), (2, 2), (3, 5), (6, 6)], [(1, 6), (4, 3)]):   # (4, 3): an empty shard
        tot = _assemble(N, X, want, 3, ranges)
        for k in ("xend", "sens", "defects_kind", "defects_stage", "grad", "jac_vals", "grad_compact", "traj_sens"):
            assert np.array_equal(tot[k], full[k]), (k, ranges)
        # the quadrature objective is a sum over the intervals: the partial sums associate differently
        assert np.allclose(tot["objective"], full["objective"], rtol=4e-16 * N.I, atol=0)
        assert not tot["status"].any()
    # a single range leaves the other intervals' entries exactly zero
    N.set_interval_range(2, 2)
    r = N.eval_host(X, nat.WANT_XEND | nat.WANT_DEFECTS, objective_row=3)
    N.set_interval_range(1, N.I)
    xe = r["xend"].reshape(B, N.I, 3)
    assert np.array_equal(xe[:, 1], full["xend"].reshape(B, N.I, 3)[:, 1]) and not xe[:, [0, 2, 3, 4, 5]].any()
    assert np.count_nonzero(r["defects_kind"][0]) <= N.n_defects // (N.I - 1)
    # back to the default: identical to the first full evaluation
    again = N.eval_host(X, want, objective_row=3)
    assert all(np.array_equal(again[k], full[k]) for k in full)


def test_ranges_sum_to_the_full_evaluation_dense20_and_lqr():
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(8)
    for spec, alg, row, B in ((P.dense20_spec(8, rng), cb.Tsit5(2), 1, 7), (P.lqr_spec(12, rng), cb.Tsit5(2), 2, 5)):
        lay, ps_t, st, N = _setup(spec, alg)
        X = cb.to_vec(ps_t)[None, :] + 0.05 * rng.standard_normal((B, N.nvars))
        want = nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD | nat.WANT_JAC
        full = N.eval_host(X, want, objective_row=row)
        half = N.I // 2
        tot = _assemble(N, X, want, row, [(1, half), (half + 1, N.I)])
        for k in ("xend", "sens", "defects_kind", "defects_stage", "grad", "jac_vals"):
            assert np.array_equal(tot[k], full[k]), (spec["model"], k)
        assert np.allclose(tot["objective"], full["objective"], rtol=4e-16 * N.I, atol=0)


def test_control_row_objective_and_rejections():
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(9)
    spec = P.lotka_spec(controls=rng.uniform(0, 1, 120), quadrature=True)
    lay, ps_t, st, N = _setup(spec, cb.Tsit5(1))
    X = cb.to_vec(ps_t)[None, :]
    want = nat.WANT_OBJ | nat.WANT_GRAD
    full = N.eval_host(X, want, objective_row=4)           # row 4 = the control active at the end
    tot = _assemble(N, X, want, 4, [(1, 2), (3, 4)])
    assert np.array_equal(tot["objective"], full["objective"]) and np.array_equal(tot["grad"], full["grad"])
    N.set_interval_range(1, 2)
    with pytest.raises(nat.NativeError, match="quadrature running sums"):
        N.eval_host(X, nat.WANT_TRAJ)
    with pytest.raises(nat.NativeError, match="outside"):
        N.set_interval_range(0, 2)
    with pytest.raises(nat.NativeError, match="outside"):
        N.set_interval_range(3, 5)
    N.set_interval_range(1, N.I)
    assert N.eval_host(X, nat.WANT_TRAJ)["traj"].shape == (1, N.n_samples * N.n_row)


def test_more_than_65535_intervals_flat_launch_equals_ranged_grid_launches():
    """a long horizon: 66 000 shooting intervals exceed gridDim.y, so the kernel runs with a flat thread index; two
    interval ranges of 33 000 each run with the regular (candidate tiles, interval, direction group) grid -- their
    sum is bit-identical to the flat evaluation"""
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(10)
    lay, ps_t, st, N = _setup(P.lqr_spec(66000, rng), cb.Tsit5(2))
    assert N.I == 66000 and N.nvars == 527998
    B = 40
    X = np.tile(cb.to_vec(ps_t), (B, 1))
    X += 0.05 * rng.standard_normal(X.shape)
    want = nat.WANT_XEND | nat.WANT_DEFECTS | nat.WANT_OBJ | nat.WANT_GRAD
    full = N.eval_host(X, want, objective_row=2)
    assert np.isfinite(full["xend"]).all() and not full["status"].any()
    tot = _assemble(N, X, want, 2, [(1, 33000), (33001, 66000)])
    for k in ("xend", "defects_kind", "defects_stage", "grad"):
        assert np.array_equal(tot[k], full[k]), k
    assert np.allclose(tot["objective"], full["objective"], rtol=1e-12, atol=0)
    # the closed form of the LQR state over one interval pins the values themselves: x' = a x + b u, piecewise constant u
    ps = cb.from_vec(X[0], ps_t)
    i = 41234
    blk = ps[f"interval_{i + 1}"]
    a, b = blk["p"]
    x = blk["u0"][0]
    for u in blk["controls"]:
        x = np.exp(a * 0.03) * x + b * u * (np.exp(a * 0.03) - 1.0) / a
    assert abs(full["xend"][0].reshape(N.I, 2)[i, 0] - x) <= 1e-9 * max(1.0, abs(x))
