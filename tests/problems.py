"""Problem specs shared by the oracle (oracle/pyoracle.py) and the product layers (corleone_b200).

A spec is a plain dict: model, u0, p0, tspan, controls=[(p_index_1based, t, u)], tunable_ic, quadrature_indices,
shooting_points, model_data.  `product_layer(spec, alg)` builds the corleone_b200 layer; `OracleLayer(spec)` the oracle.
"""
import numpy as np


def lotka_spec(shooting_points=(0.0, 3.0, 6.0, 9.0), controls=None, quadrature=False, tunable_ic=()):
    """test/core/multiple_shooting.jl:11-37 / test/examples/lotka_ms.jl"""
    cgrid = np.arange(120) / 10.0
    return dict(model="lotka3", u0=[0.5, 0.7, 0.0], p0=[0.0, 1.0, 1.0], tspan=(0.0, 12.0),
                controls=[(1, cgrid, np.zeros(120) if controls is None else np.asarray(controls, float))],
                tunable_ic=list(tunable_ic), quadrature_indices=[3] if quadrature else [],
                shooting_points=None if shooting_points is None else list(shooting_points))


def lotka_ms24_spec(rng):
    """BASELINE config 1: 24 intervals of 0.5, control grid k/10, controls U[0,1]."""
    s = lotka_spec(shooting_points=[0.5 * k for k in range(24)], controls=rng.uniform(0, 1, 120))
    return s


def lqr_spec(n_intervals=100, rng=None):
    """BASELINE config 2: LQR, 100 intervals of 0.12, control step 0.03 (400 values)."""
    nctl = 4 * n_intervals
    t = 3.0 * np.arange(nctl) / 100.0          # k * 0.03, correctly rounded
    u = np.zeros(nctl) if rng is None else rng.standard_normal(nctl)
    pts = [12.0 * k / n_intervals for k in range(n_intervals)]
    return dict(model="lqr", u0=[1.0, 0.0], p0=[-1.0, 1.0, 0.0], tspan=(0.0, 12.0 * (n_intervals / 100.0)) if n_intervals != 100 else (0.0, 12.0),
                controls=[(3, t, u)], tunable_ic=[], quadrature_indices=[],
                shooting_points=[3.0 * 4 * k / 100.0 for k in range(n_intervals)])


def lotka_oed_spec(shooting_points=None, fishing=None, w=1.0, grids=None):
    """lib/CorleoneOED/test/core/lotka_oed.jl:15-48: augmented [x; G; F_triu], controls fishing, w1, w2."""
    if grids is None:
        cg = 0.25 * np.arange(48)
        g1 = 0.5 * np.arange(24)
        g2 = 3.0 * np.arange(79) / 20.0          # 0.0:0.15:11.75 -> k*0.15
    else:
        cg, g1, g2 = grids
    return dict(model="lotka2_oed", u0=[0.5, 0.7] + [0.0] * 7, p0=[0.0, 1.0, 1.0, 1.0, 1.0], tspan=(0.0, 12.0),
                controls=[(1, cg, np.zeros(len(cg)) if fishing is None else np.asarray(fishing, float)),
                          (4, g1, np.full(len(g1), w)), (5, g2, np.full(len(g2), w))],
                tunable_ic=[], quadrature_indices=[],
                shooting_points=None if shooting_points is None else list(shooting_points))


def lin1d_oed_spec():
    """lib/CorleoneOED/test/core/1d_oed.jl:11-31"""
    t = np.arange(100) / 100.0
    return dict(model="lin1d_oed", u0=[1.0, 0.0, 0.0], p0=[-2.0, 1.0], tspan=(0.0, 1.0),
                controls=[(2, t, np.ones(100))], tunable_ic=[], quadrature_indices=[], shooting_points=None)


def egerstedt_spec():
    """test/core/local_controls.jl:45-79: three controls passed in the order [2, 3, 1]"""
    N = 20
    cgrid = np.linspace(0.0, 1.0, N + 1)[:-1]
    c1, c2, c3 = np.linspace(0.0, 0.2, N), np.linspace(0.3, 0.5, N), np.linspace(0.6, 0.8, N)
    return dict(model="egerstedt", u0=[0.5, 0.5, 0.0], p0=[1 / 3] * 3, tspan=(0.0, 1.0),
                controls=[(2, cgrid, c2), (3, cgrid, c3), (1, cgrid, c1)], control_names=["con2", "con3", "con1"],
                tunable_ic=[], quadrature_indices=[], shooting_points=None)


def dense20_data(seed=20261023):
    rng = np.random.default_rng(seed)
    W = 0.5 * rng.standard_normal((20, 20)) / np.sqrt(20.0)
    b = rng.standard_normal(20)
    return np.concatenate([W.ravel(), b])


def dense20_spec(n_intervals=8, rng=None):
    """BASELINE config 4 (synthetic): x' = W x - x^3 + b u, M = 4 pieces per interval."""
    nctl = 4 * n_intervals
    T = float(n_intervals) * 0.5
    t = T * np.arange(nctl) / nctl
    u = np.zeros(nctl) if rng is None else rng.uniform(-1, 1, nctl)
    x0 = np.zeros(20) if rng is None else rng.uniform(-1, 1, 20)
    return dict(model="dense20", u0=list(x0), p0=[0.0], tspan=(0.0, T), controls=[(1, t, u)], tunable_ic=[],
                quadrature_indices=[], shooting_points=[T * k / n_intervals for k in range(n_intervals)],
                model_data=dense20_data())


def product_layer(spec, alg, **kw):
    import corleone_b200 as cb
    prob = cb.ODEProblem(spec["model"], spec["u0"], spec["tspan"], spec["p0"], model_data=spec.get("model_data"))
    names = spec.get("control_names") or [f"u{k + 1}" for k in range(len(spec["controls"]))]
    ctr = [(ci, cb.ControlParameter(t, name=n, controls=(np.zeros(len(t)) if u is None else u)))
           for (ci, t, u), n in zip(spec["controls"], names)]
    args = dict(controls=ctr, tunable_ic=spec.get("tunable_ic", []), quadrature_indices=spec.get("quadrature_indices", []))
    args.update(kw)
    if spec.get("shooting_points") is None:
        return cb.SingleShootingLayer(prob, alg, **args)
    return cb.MultipleShootingLayer(prob, alg, *spec["shooting_points"], **args)
