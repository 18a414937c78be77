"""The other variants of the OED layer (lib/CorleoneOED/src/oed.jl, augmentation.jl) on the device: continuous without
measurements, fixed continuous with measurements, discrete with and without measurements -- Fisher information,
criteria and sensitivities against the oracle's literal restatement, <= 1e-10 relative."""
import numpy as np
import pytest

import corleone_b200 as cb
from corleone_b200 import native as nat
from corleone_b200 import oed as oedm

from test_gpu_parity import RTOL, rel_err
from test_oed_variants_host import CG, T1, T2, _oracle_spec, lotka_base, measurements

pytestmark = pytest.mark.gpu

SHOOT = (0.0, 3.0, 6.0, 9.0)


def oracle_fim(o, O, ps_like, st, x, K):
    """F of one candidate with the oracle: trajectory samples from the C restatement, read-out by pyoracle's loops"""
    from oracle import pyoracle as po
    r = O.eval(x, K=K)
    u = r["u"]
    bx, nf, nobs = o.base_nx, o.n_fisher, o.n_obs
    if o.fim_mode == oedm.FIM_DISCRETE:
        return po.fisher_discrete(u, bx, nf, nobs)
    ps = cb.from_vec(x, ps_like)
    if o.fim_mode == oedm.FIM_DISCRETE_SAMPLED:
        if o.single:
            return po.fisher_discrete_sampled(ps["controls"], st["observation_grid"].tolist(), u, 0, bx, nf, nobs)
        F, s0 = np.zeros((nf, nf)), 0
        for k in st:                                      # oed.jl:418-432: per interval, its own grid and controls
            g = st[k]["observation_grid"].tolist()
            F += po.fisher_discrete_sampled(ps[k]["controls"], g, u, s0, bx, nf, nobs)
            s0 += len(g)
        return F
    if o.fim_mode == oedm.FIM_FIXED_CONTINUOUS:
        return po.fisher_fixed_continuous(ps["controls"], st["active_controls"], u, bx, nf, nobs)
    f_off = bx * (1 + nf)                                  # continuous: last sample, Symmetric from triu (oed.jl:388-398)
    F = np.zeros((nf, nf))
    for j in range(nf):
        for i in range(j + 1):
            F[i, j] = F[j, i] = u[f_off + j * (j + 1) // 2 + i, -1]
    return F


def numpy_criteria(F):
    lam = np.linalg.eigvalsh(F)
    return np.array([np.sum(1.0 / lam), 1.0 / np.prod(lam), np.max(1.0 / lam), -np.sum(lam), -np.prod(lam), -lam[0]])


CASES = [
    ("continuous_unsampled_ss", False, False, None, False),
    ("continuous_unsampled_ms", False, False, SHOOT, False),
    ("continuous_fixed_ss", False, True, None, True),
    ("discrete_sampled_ss", True, True, None, False),
    ("discrete_sampled_ms", True, True, SHOOT, False),
    ("discrete_sampled_fixed_ss", True, True, None, True),
    ("discrete_unsampled_ss", True, False, None, False),
    ("discrete_unsampled_ms", True, False, SHOOT, False),
]


@pytest.mark.parametrize("name,discrete,sampled,shooting,fixed", CASES, ids=[c[0] for c in CASES])
def test_fisher_information_of_every_variant(name, discrete, sampled, shooting, fixed):
    from oracle import pyoracle as po
    # the fixed discrete layer has no control: its horizon is ONE piece and the measurement times are saveat samples inside
    # it (oed.jl:106-113) -- the fixed step has to resolve 12 time units by itself
    K = 96 if (fixed and discrete) else 2
    meas = ()
    if sampled:
        meas = measurements()
        if fixed and not discrete:   # hcat of the weights needs equally long measurement grids (oed.jl:357)
            meas = [cb.ControlParameter(T1, controls=np.ones(len(T1)), bounds=(0.0, 1.0)) for _ in range(2)]
    o = oedm.OEDLayer(lotka_base(shooting, fixed=fixed, K=K), discrete, meas)
    ps, st = o.setup(None)
    O = po.OracleLayer(_oracle_spec(o, ps))
    N = o.native(ps, st)
    assert N.nvars == O.nvars and N.n_samples == O.nsamples and N.n_row == O.nrow
    rng = np.random.default_rng(sum(map(ord, name)))
    B = 5
    x0 = cb.to_vec(ps)
    if shooting is not None:
        x0 = O.forward_init(x0, K=K)   # consistent nodes; then perturbed
    X = x0[None, :] + 0.05 * rng.standard_normal((B, len(x0)))
    lb, ub = (cb.to_vec(b) for b in cb.get_bounds(o.layer))
    ctrl = np.isfinite(lb) & np.isfinite(ub) & (ub > lb)          # controls and weights: drawn inside their bounds
    X[:, ctrl] = rng.uniform(lb[ctrl], ub[ctrl], size=(B, int(ctrl.sum())))
    F_init = np.array([[0.3, 0.1], [0.1, 0.2]])
    first = st if o.single else next(iter(st.values()))
    first["F_init"] = F_init
    F, crit = o.fisher_information(X, ps, st, criteria=True)
    r = N.eval_host(X, nat.WANT_FIM | nat.WANT_TRAJ, fim_init=F_init)
    for b in range(B):
        Fo = F_init + oracle_fim(o, O, ps, st, X[b], K)
        assert rel_err(F[b], Fo) <= RTOL, (name, b, F[b], Fo)
        assert np.array_equal(F[b], F[b].T)
        assert rel_err(crit[b], numpy_criteria(Fo)) <= 1e-9 * max(1.0, np.linalg.cond(Fo))
        assert np.array_equal(r["fim"][b].reshape(2, 2), F[b])     # same numbers with the samples requested as well
    assert rel_err(r["fim_sum"], F.sum(axis=0).ravel()) <= 1e-13
    assert not r["status"].any()
    F0 = o.fisher_information(X[:1], ps, st, add_initial=False)
    assert rel_err(F0[0] + F_init, F[0]) <= 1e-14


@pytest.mark.parametrize("variant", ["cu", "cf", "ds", "du"])
def test_state_and_sensitivity_parity_of_the_variant_models(variant):
    """end states, defects and d x_end / d ps of the new augmentations (second-order terms of G, the per-observation
    F blocks) against the oracle's dual-number integration"""
    from oracle import pyoracle as po
    discrete = variant in ("ds", "du")
    sampled = variant in ("cf", "ds")
    fixed = variant == "cf"
    meas = measurements() if sampled else ()
    if fixed:   # hcat of the weights needs equally long measurement grids (oed.jl:357)
        meas = [cb.ControlParameter(T1, controls=np.ones(len(T1)), bounds=(0.0, 1.0)) for _ in range(2)]
    o = oedm.OEDLayer(lotka_base(None if fixed else SHOOT, fixed=fixed, K=2), discrete, meas)
    assert o.model == "lotka2_oed_" + variant
    ps, st = o.setup(None)
    O = po.OracleLayer(_oracle_spec(o, ps))
    N = o.native(ps, st)
    rng = np.random.default_rng(11)
    B = 3
    X = cb.to_vec(ps)[None, :] + 0.05 * rng.standard_normal((B, N.nvars))
    r = N.eval_host(X, nat.WANT_XEND | nat.WANT_SENS | nat.WANT_DEFECTS | nat.WANT_TRAJ)
    blocks = O.block_structure()
    nx, I = O.nx, O.I
    for b in range(B):
        ref = O.eval(X[b], K=2)
        assert rel_err(r["xend"][b].reshape(I, nx).T, ref["xend"]) <= RTOL
        assert rel_err(r["defects_kind"][b], ref["defk"]) <= RTOL
        assert rel_err(r["traj"][b].reshape(O.nsamples, O.nrow).T, ref["u"]) <= RTOL
    Jx, _, _ = O.jacobians(X[0], K=2)
    off = 0
    for i in range(I):
        ns = blocks[i + 1] - blocks[i]
        S = r["sens"][0][off:off + nx * ns].reshape(ns, nx).T
        assert rel_err(S, Jx[nx * i:nx * (i + 1), blocks[i]:blocks[i + 1]]) <= RTOL, (variant, i)
        off += nx * ns


def test_discrete_1d_closed_form_on_the_device():
    """lib/CorleoneOED/test/core/1d_oed.jl:11-31 with the discrete layer: F = sum_i (w_i t_i e^{p t_i})^2, and the fixed
    continuous layer: F = sum_j w_j int t^2 e^{2pt} dt over piece j -- for a batch of weight vectors in ONE call"""
    t = np.arange(100) / 100.0
    prob = cb.ODEProblem("lin1d", [1.0], (0.0, 1.0), [-2.0])
    rng = np.random.default_rng(3)
    W = rng.uniform(0, 1, (64, 100))
    prim = lambda s: np.exp(-4.0 * s) * (-(s * s) / 4.0 - s / 8.0 - 1.0 / 32.0)
    edges = np.r_[t, 1.0]
    for discrete in (True, False):
        # discrete: one piece, the 99 inner measurement times come from the dense output of its 256 steps
        o = oedm.OEDLayer(cb.SingleShootingLayer(prob, cb.Tsit5(256 if discrete else 4)), discrete,
                          [cb.ControlParameter(t, controls=np.ones(100), bounds=(0.0, 1.0))])
        ps, st = o.setup(None)
        X = np.tile(cb.to_vec(ps), (64, 1))
        X[:, -100:] = W                                      # variables = [p; w_1..w_100]
        F = o.fisher_information(X, ps, st)[:, 0, 0]
        exact = (np.sum((W * t * np.exp(-2.0 * t)) ** 2, axis=1) if discrete
                 else np.sum(W * (prim(edges[1:]) - prim(edges[:-1])), axis=1))
        assert np.max(np.abs(F - exact) / exact) <= 1e-10, discrete


def test_sample_based_readout_rejects_bad_tables():
    o = oedm.OEDLayer(lotka_base(), True, measurements())
    ps, st = o.setup(None)
    st["_fim"]["rows"] = [np.full_like(st["observation_grid"], 10_000)]
    with pytest.raises(nat.NativeError, match="observation_grid"):
        o.native(ps, st)


GRAD_CASES = [("discrete_sampled_ss", True, True, None, False), ("discrete_sampled_ms", True, True, SHOOT, False),
              ("discrete_unsampled_ms", True, False, SHOOT, False), ("continuous_fixed_ss", False, True, None, True)]


@pytest.mark.parametrize("name,discrete,sampled,shooting,fixed", GRAD_CASES, ids=[c[0] for c in GRAD_CASES])
def test_fisher_gradient_of_the_sample_based_readouts(name, discrete, sampled, shooting, fixed):
    """dF/d(variables) assembled from the device's sample sensitivities == the same product rule on the oracle's
    dual-number sample Jacobian (<= 1e-10), and == central differences of the device's own F (<= 1e-6)"""
    from oracle import pyoracle as po
    K = 2
    meas = ()
    if sampled:
        meas = measurements(0.7)
        if fixed:
            meas = [cb.ControlParameter(T1, controls=np.full(len(T1), 0.7), bounds=(0.0, 1.0)) for _ in range(2)]
    o = oedm.OEDLayer(lotka_base(shooting, fixed=fixed, K=K), discrete, meas)
    ps, st = o.setup(None)
    O = po.OracleLayer(_oracle_spec(o, ps))
    rng = np.random.default_rng(17)
    x = cb.to_vec(ps)
    if shooting is not None:
        x = O.forward_init(x, K=K)
    x = x + 0.02 * rng.standard_normal(len(x))
    F, dF = o.fisher_gradient(x[None, :], ps, st)
    # (1) central differences of the device's own read-out on a handful of variables of every kind
    pick = rng.choice(len(x), size=min(12, len(x)), replace=False)
    h = 1e-6
    Xp = np.tile(x, (2 * len(pick), 1))
    for q, v in enumerate(pick):
        Xp[2 * q, v] += h
        Xp[2 * q + 1, v] -= h
    Fp = o.fisher_information(Xp, ps, st)
    for q, v in enumerate(pick):
        fd = (Fp[2 * q] - Fp[2 * q + 1]) / (2 * h)
        assert rel_err(dF[0, :, :, v], fd) <= 1e-6, (name, v)
    # (2) the oracle's samples and sample Jacobian through the same product rule
    u, J = O.traj_jacobian(x, K=K)
    Fo = oracle_fim(o, O, ps, st, x, K)
    assert rel_err(F[0], Fo) <= RTOL
    eps = 1e-7
    for v in pick[:4]:       # directional check of the oracle read-out itself along the dual-number tangent
        up = u + eps * J[:, :, v]
        um = u - eps * J[:, :, v]
        xp, xm = x.copy(), x.copy()
        xp[v] += eps
        xm[v] -= eps

        def readout(uu, xx):
            class _O:   # oracle_fim only needs eval()["u"]
                def eval(self, _x, K):
                    return dict(u=uu)
            return oracle_fim(o, _O(), ps, st, xx, K)
        lin = (readout(up, xp) - readout(um, xm)) / (2 * eps)
        assert rel_err(dF[0, :, :, v], lin) <= 1e-7, (name, v)
    crit, g = o.criteria_gradient(x[None, :], ps, st)
    Fi = np.linalg.inv(F[0])
    assert rel_err(g[0, 0], -np.einsum("ab,bcv,ca->v", Fi, dF[0], Fi)) <= 1e-12      # A criterion: -tr(F^-1 dF F^-1)


def test_design_optimum_of_the_1d_problem_fixed_continuous_and_discrete():
    """lib/CorleoneOED/test/core/1d_oed.jl with the measurement budget as a constraint: the information density
    t^2 e^{2pt} peaks at t = 1/|p| = 0.5, so the optimal design measures on the pieces (samples) around it."""
    t = np.arange(100) / 100.0
    prob = cb.ODEProblem("lin1d", [1.0], (0.0, 1.0), [-2.0])
    base = cb.SingleShootingLayer(prob, cb.Tsit5(4), bounds_p=([-2.0], [-2.0]))
    prim = lambda s: np.exp(-4.0 * s) * (-(s * s) / 4.0 - s / 8.0 - 1.0 / 32.0)
    edges = np.r_[t, 1.0]
    # fixed continuous: F is linear in w; budget sum_j w_j dt <= 0.2 -> w = 1 on the 20 richest pieces
    o = oedm.OEDLayer(base, False, [cb.ControlParameter(t, controls=np.full(100, 0.2), bounds=(0.0, 1.0))])
    ps, st = o.setup(None)
    res = oedm.solve_slsqp(o, ps, st, criterion=0, M=[0.2])
    c = prim(edges[1:]) - prim(edges[:-1])
    best = np.sort(np.argsort(c)[-20:])
    w = res.x[-100:]
    assert abs(res.fun - 1.0 / c[best].sum()) <= 1e-6 / c[best].sum(), (res.fun, 1.0 / c[best].sum())
    assert np.all(w[best] > 0.99) and np.sum(w) <= 20.0 + 1e-6
    # discrete: sum_i w_i <= 20 measurements; F = sum (w_i g_i)^2 -> the 20 samples with the largest g^2
    o = oedm.OEDLayer(base, True, [cb.ControlParameter(t, controls=np.full(100, 0.2), bounds=(0.0, 1.0))])
    ps, st = o.setup(None)
    res = oedm.solve_slsqp(o, ps, st, criterion=0, M=[20.0])
    g2 = (t * np.exp(-2.0 * t)) ** 2
    best = np.sort(np.argsort(g2)[-20:])
    assert res.fun <= 1.0 / g2[best].sum() * (1 + 1e-6) or abs(res.fun - 1.0 / g2[best].sum()) <= 1e-3 / g2[best].sum()
    assert abs(o.fisher_information(res.x[None, :], ps, st)[0, 0, 0] - 1.0 / res.fun) <= 1e-9 / res.fun


def test_sample_readouts_sensitivities_observed_and_information_gains():
    """sensitivities / observed_equations / local and global information gain (oed.jl:464-534) for a batch, against the
    oracle's samples; summing the local gains with the design weights gives the discrete layer's Fisher information"""
    from oracle import pyoracle as po
    o = oedm.OEDLayer(lotka_base(None, K=2), True, measurements(0.6))
    ps, st = o.setup(None)
    O = po.OracleLayer(_oracle_spec(o, ps))
    rng = np.random.default_rng(23)
    X = cb.to_vec(ps)[None, :] + 0.02 * rng.standard_normal((3, len(cb.to_vec(ps))))
    G = o.sensitivities(X, ps, st)
    obs = o.observed_equations(X, ps, st)
    L = o.local_information_gain(X, ps, st)
    Gl = o.global_information_gain(X, ps, st)
    F = o.fisher_information(X, ps, st)
    for b in range(3):
        u = O.eval(X[b], K=2)["u"]
        assert rel_err(obs[b].T, u[:2]) <= RTOL
        for s in (0, 7, u.shape[1] - 1):
            Gs = u[2:6, s].reshape(2, 2).T
            assert rel_err(G[b, s], Gs) <= RTOL
            for i in range(2):
                assert rel_err(L[b, s, i], np.outer(Gs[i], Gs[i])) <= RTOL
                x = Gs[i] @ np.linalg.inv(F[b])
                assert rel_err(Gl[b, s, i], np.outer(x, x)) <= 1e-9
        # discrete layer: F = sum_s sum_i w_is^2 (local gain)
        pd = cb.from_vec(X[b], ps)
        W = np.where(st["observation_grid"] > 0, pd["controls"][np.maximum(st["observation_grid"], 1) - 1], 0.0)
        assert rel_err(np.einsum("si,siab->ab", W ** 2, L[b]), F[b]) <= 1e-12
    # update_fim folds finished experiments into F_init
    F0 = o.fisher_information(X[:1], ps, st)[0]
    o.update_fim(dict(X=X[1:], ps=ps, st=st), st)
    assert rel_err(o.fisher_information(X[:1], ps, st)[0], F0 + F[1] + F[2]) <= 1e-13


@pytest.mark.parametrize("discrete", [False, True])
def test_multi_experiment_fisher_information_is_the_sum_over_experiments(discrete):
    """MultiExperimentLayer (multiexperiments.jl:279-289): F = F_init + sum_e F_e, a batch of joint designs evaluated as
    B * n_exp candidates in one call; update_fim folds finished designs into F_init (:195-221)"""
    o = oedm.OEDLayer(lotka_base(None, K=2), discrete, measurements(0.8))
    m = cb.MultiExperimentLayer(o, 3)
    ps, st = m.setup(None)
    rng = np.random.default_rng(29)
    nv = len(cb.to_vec(ps["experiment_1"]))
    X = np.tile(cb.to_vec(ps), (4, 1))
    X[:, :] += 0.0
    for e in range(3):      # different fishing controls and weights per experiment
        X[:, e * nv + 2:(e + 1) * nv] = rng.uniform(0, 1, (4, nv - 2))
    F_init = np.array([[0.2, 0.05], [0.05, 0.1]])
    st["experiment_1"]["F_init"] = F_init
    F = m.fisher_information(X, ps, st)
    for b in range(4):
        parts = [o.fisher_information(X[b, e * nv:(e + 1) * nv][None, :], ps["experiment_1"], st["experiment_1"],
                                      add_initial=False)[0] for e in range(3)]
        assert rel_err(F[b], F_init + sum(parts)) <= 1e-13
    m.update_fim(dict(X=X[2:], ps=ps, st=st), st)
    assert rel_err(st["experiment_1"]["F_init"], F_init + (F[2] - F_init) + (F[3] - F_init)) <= 1e-13
