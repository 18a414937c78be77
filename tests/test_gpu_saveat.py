"""`saveat` inside control pieces on the device (cb200_desc.saveat: Tsit5's dense output on the step that contains the
time) against the CPU checker: samples, their sensitivities, and everything downstream; fixed step and adaptive.
Reference: problem.kwargs saveat, lib/CorleoneOED/src/oed.jl:106-113; piece solves src/single_shooting.jl:443-453."""
import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu


def lotka_spec_saveat(rng, shooting, saveat, model="lotka3"):
    grid = np.sort(np.r_[0.0, rng.uniform(0.0, 6.0, 9)])
    spec = dict(model=model, u0=[0.5, 0.7, 0.0] if model == "lotka3" else [0.5, 0.7], p0=[0.0, 1.0, 1.0], tspan=(0.0, 6.0),
                controls=[(1, grid, rng.uniform(0, 1, len(grid)))], tunable_ic=[], saveat=list(saveat),
                quadrature_indices=[3] if model == "lotka3" else [], shooting_points=shooting)
    return spec


@pytest.mark.parametrize("alg", ["fixed", "adaptive"])
@pytest.mark.parametrize("shooting", [None, [0.0, 1.5, 3.1, 4.4]])
def test_saveat_samples_and_sample_sensitivities_match_the_checker(alg, shooting):
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    from oracle import pyoracle as po
    rng = np.random.default_rng(3)
    saveat = np.sort(np.r_[rng.uniform(0.0, 6.0, 23), 1.5, 3.0])      # some inside pieces, one on a node, one wherever
    spec = lotka_spec_saveat(rng, shooting, saveat)
    if alg == "adaptive":
        spec["kwargs"] = dict(abstol=1e-9, reltol=1e-8)
    O = po.OracleLayer(spec)
    A = cb.Tsit5(3) if alg == "fixed" else cb.Tsit5()
    lay, ps_t, st, N = P.native_from_spec(spec, A)
    assert N.n_samples == O.nsamples
    B = 5
    ps = O.default_ps()[None, :] + 0.03 * rng.standard_normal((B, O.nvars))
    r = N.eval_host(ps, nat.WANT_TRAJ | nat.WANT_TRAJ_SENS | nat.WANT_SENS | nat.WANT_XEND | nat.WANT_DEFECTS)
    kw = dict(method="tsit5", K=3) if alg == "fixed" else dict(method="tsit5_adaptive")
    if alg == "adaptive":
        import ctypes
        po.lib().co_set_tolerances(ctypes.c_double(1e-9), ctypes.c_double(1e-8))
    try:
        for b in range(B):
            o = O.eval(ps[b], **kw)
            assert np.array_equal(N.sample_times(), o["t"])
            u = r["traj"][b].reshape(N.n_samples, N.n_row).T
            assert np.max(np.abs(u - o["u"])) <= 1e-10
            assert np.max(np.abs(r["xend"][b].reshape(O.I, O.nx).T - o["xend"])) <= 1e-10
            if O.ndef:
                assert np.max(np.abs(r["defects_kind"][b] - o["defk"])) <= 1e-10
            if b < 2:
                from corleone_b200.multiple_shooting import trajectory_jacobian
                J = trajectory_jacobian(lay, st, ps_t, r, b)
                _, Jo = O.traj_jacobian(ps[b], **kw)
                tol = 1e-10 if alg == "fixed" else 1e-8    # adaptive tangents: the controller's Float32 step quantum
                assert np.max(np.abs(J - Jo)) <= tol * max(1.0, np.max(np.abs(Jo)))
    finally:
        if alg == "adaptive":
            po.lib().co_set_tolerances(ctypes.c_double(1e-6), ctypes.c_double(1e-3))


def test_saveat_needs_tsit5_and_is_refused_elsewhere():
    import corleone_b200 as cb
    from corleone_b200 import native as nat
    rng = np.random.default_rng(1)
    spec = lotka_spec_saveat(rng, None, [0.01, 2.2])
    with pytest.raises(nat.NativeError, match="dense output"):
        P.native_from_spec(spec, cb.RK4(2))
    # a model without the sampling instantiation reports it instead of skipping the samples
    spec = P.lqr_spec(4, rng)
    spec["saveat"] = [0.013]
    lay, ps_t, st, N = P.native_from_spec(spec, cb.Tsit5(2))
    with pytest.raises(nat.NativeError, match="no kernel"):
        N.eval_host(cb.to_vec(ps_t)[None, :], nat.WANT_TRAJ)
