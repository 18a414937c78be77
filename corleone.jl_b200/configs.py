"""The named BASELINE.json configs as concrete synthetic problems.

A spec is plain data (model, u0, p0, tspan, controls=[(p_index_1based, grid, values)], tunable_ic,
quadrature_indices, shooting_points, model_data) so that the product layers and the test oracle are built
from the same description.  Seeds: numpy default_rng(20261019 + config number).
"""
from __future__ import annotations

import numpy as np

from .controls import ControlParameter
from .multiple_shooting import MultipleShootingLayer
from .problem import ODEProblem
from .single_shooting import SingleShootingLayer

SEED0 = 20261019


def lotka_spec(shooting_points=(0.0, 3.0, 6.0, 9.0), controls=None, quadrature=False, tunable_ic=()):
    """Lotka fishing problem of the reference tests (test/core/multiple_shooting.jl:11-37)."""
    cgrid = np.arange(120) / 10.0   # 0.0:0.1:11.9 as Julia's range produces it (k/10 correctly rounded)
    return dict(model="lotka3", u0=[0.5, 0.7, 0.0], p0=[0.0, 1.0, 1.0], tspan=(0.0, 12.0),
                controls=[(1, cgrid, np.zeros(120) if controls is None else np.asarray(controls, float))],
                tunable_ic=list(tunable_ic), quadrature_indices=[3] if quadrature else [],
                shooting_points=None if shooting_points is None else list(shooting_points))


def config1_lotka_ms24(rng=None):
    """config 1: Lotka fishing, 24 intervals of 0.5, control grid k/10 (M = 5), controls U[0,1]."""
    rng = np.random.default_rng(SEED0 + 1) if rng is None else rng
    return lotka_spec(shooting_points=[0.5 * k for k in range(24)], controls=rng.uniform(0, 1, 120))


def config2_lqr(n_intervals=100, rng=None):
    """config 2: LQR (examples/linear_quadratic/main.jl:45-68), 100 intervals of 0.12, control step 0.03
    (M = 4 pieces per interval, 400 control values N(0,1))."""
    rng = np.random.default_rng(SEED0 + 2) if rng is None else rng
    nctl = 4 * n_intervals
    t = 3.0 * np.arange(nctl) / 100.0                     # k * 0.03 correctly rounded
    return dict(model="lqr", u0=[1.0, 0.0], p0=[-1.0, 1.0, 0.0], tspan=(0.0, 3.0 * nctl / 100.0),
                controls=[(3, t, rng.standard_normal(nctl))], tunable_ic=[], quadrature_indices=[],
                shooting_points=[3.0 * (4 * k) / 100.0 for k in range(n_intervals)])


def lotka_oed_spec(shooting_points=None, fishing=None, w=1.0, grids=None):
    """Augmented Lotka OED [x; G; F_triu] with controls (fishing, w1, w2)
    (lib/CorleoneOED/test/core/lotka_oed.jl:15-48 by default)."""
    if grids is None:
        cg, g1, g2 = 0.25 * np.arange(48), 0.5 * np.arange(24), 3.0 * np.arange(79) / 20.0
    else:
        cg, g1, g2 = grids
    return dict(model="lotka2_oed", u0=[0.5, 0.7] + [0.0] * 7, p0=[0.0, 1.0, 1.0, 1.0, 1.0], tspan=(0.0, 12.0),
                controls=[(1, cg, np.zeros(len(cg)) if fishing is None else np.asarray(fishing, float)),
                          (4, g1, np.full(len(g1), w)), (5, g2, np.full(len(g2), w))],
                tunable_ic=[], quadrature_indices=[],
                shooting_points=None if shooting_points is None else list(shooting_points), fim=(7, 2))


def config3_lotka_oed(n_intervals=64, rng=None):
    """config 3: 64 intervals of 3/16, every grid with step 3/32 (M = 2), fishing U[0,1], w = 1."""
    rng = np.random.default_rng(SEED0 + 3) if rng is None else rng
    n = 2 * n_intervals
    g = 3.0 * np.arange(n) / 32.0
    s = lotka_oed_spec(shooting_points=[3.0 * (2 * k) / 32.0 for k in range(n_intervals)],
                       fishing=rng.uniform(0, 1, n), grids=(g, g, g))
    s["tspan"] = (0.0, 3.0 * n / 32.0)
    return s


def dense20_data(seed=SEED0 + 4, nx=20):
    rng = np.random.default_rng(seed)
    W = 0.5 * rng.standard_normal((nx, nx)) / np.sqrt(float(nx))
    return np.concatenate([W.ravel(), rng.standard_normal(nx)])


def config4_dense20(n_intervals=256, rng=None, nx=20):
    """config 4 (synthetic): x' = W x - x^3 + b u, 20 states, M = 4 pieces per interval of length 0.5
    (nx = 16, 24, 32: the same family at the other sizes of the tensor-path kernel)."""
    rng = np.random.default_rng(SEED0 + 4) if rng is None else rng
    nctl = 4 * n_intervals
    t = np.arange(nctl) / 8.0
    return dict(model=f"dense{nx}", u0=list(rng.uniform(-1, 1, nx)), p0=[0.0], tspan=(0.0, nctl / 8.0),
                controls=[(1, t, rng.uniform(-1, 1, nctl))], tunable_ic=[], quadrature_indices=[],
                shooting_points=[(4 * k) / 8.0 for k in range(n_intervals)], model_data=dense20_data(nx=nx))


def lin1d_oed_spec():
    """lib/CorleoneOED/test/core/1d_oed.jl:11-31"""
    t = np.arange(100) / 100.0
    return dict(model="lin1d_oed", u0=[1.0, 0.0, 0.0], p0=[-2.0, 1.0], tspan=(0.0, 1.0),
                controls=[(2, t, np.ones(100))], tunable_ic=[], quadrature_indices=[], shooting_points=None,
                fim=(3, 1))


def egerstedt_spec():
    """test/core/local_controls.jl:45-79: three controls passed in the order [2, 3, 1]"""
    N = 20
    cgrid = np.linspace(0.0, 1.0, N + 1)[:-1]
    c1, c2, c3 = np.linspace(0.0, 0.2, N), np.linspace(0.3, 0.5, N), np.linspace(0.6, 0.8, N)
    return dict(model="egerstedt", u0=[0.5, 0.5, 0.0], p0=[1 / 3] * 3, tspan=(0.0, 1.0),
                controls=[(2, cgrid, c2), (3, cgrid, c3), (1, cgrid, c1)], control_names=["con2", "con3", "con1"],
                tunable_ic=[], quadrature_indices=[], shooting_points=None)


def vdp_spec(n_intervals=5, rng=None):
    """Van der Pol with Lagrange state (lib/OptimalControlBenchmarks/src/problems/van_der_pol.jl), 40 control values"""
    rng = np.random.default_rng(SEED0 + 8) if rng is None else rng
    t = np.arange(40) / 4.0
    return dict(model="vdp3", u0=[0.0, 1.0, 0.0], p0=[0.0], tspan=(0.0, 10.0),
                controls=[(1, t, rng.uniform(-1, 1, 40))], tunable_ic=[], quadrature_indices=[3],
                shooting_points=[10.0 * k / n_intervals for k in range(n_intervals)])


def cartpole_spec(n_intervals=4, rng=None):
    """cart pendulum with Lagrange state (.../cart_pendulum.jl), 32 control values on (0, 4)"""
    rng = np.random.default_rng(SEED0 + 9) if rng is None else rng
    t = np.arange(32) / 8.0
    return dict(model="cartpole5", u0=[0.0, 0.0, 0.0, 0.0, 0.0], p0=[0.0], tspan=(0.0, 4.0),
                controls=[(1, t, rng.uniform(-3, 3, 32))], tunable_ic=[], quadrature_indices=[],
                shooting_points=[4.0 * k / n_intervals for k in range(n_intervals)])


def rocket_spec(n_intervals=4, rng=None):
    """Goddard rocket with free final time T as tunable parameter (.../goddarts_rocket.jl), normalised time (0, 1)"""
    rng = np.random.default_rng(SEED0 + 10) if rng is None else rng
    t = np.arange(20) / 20.0
    return dict(model="rocket3", u0=[1.0, 0.0, 1.0], p0=[0.2, 0.5], tspan=(0.0, 1.0),
                controls=[(2, t, rng.uniform(0.2, 0.8, 20))], tunable_ic=[], quadrature_indices=[],
                shooting_points=[k / n_intervals for k in range(n_intervals)])


def quadrotor_spec(n_intervals=3, rng=None):
    """quadrotor with Lagrange state (.../quadrotor.jl): four controls (w1, w2, w3, u) on different grids"""
    rng = np.random.default_rng(SEED0 + 11) if rng is None else rng
    tw, tu = np.arange(15) / 2.0, np.arange(30) / 4.0
    return dict(model="quadrotor7", u0=[0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0], p0=[1 / 3, 1 / 3, 1 / 3, 0.0],
                tspan=(0.0, 7.5),
                controls=[(1, tw, rng.uniform(0, 1, 15)), (2, tw, rng.uniform(0, 1, 15)), (3, tw, rng.uniform(0, 1, 15)),
                          (4, tu, rng.uniform(0, 1, 30))],
                tunable_ic=[], quadrature_indices=[7], shooting_points=[7.5 * k / n_intervals for k in range(n_intervals)])


def threetank_spec(n_intervals=4, rng=None):
    """three-tank system with Lagrange state (.../three_tank.jl): sqrt right-hand side, three controls"""
    rng = np.random.default_rng(SEED0 + 12) if rng is None else rng
    t = np.arange(24) / 4.0
    return dict(model="threetank4", u0=[2.0, 2.0, 2.0, 0.0], p0=[1 / 3, 1 / 3, 1 / 3], tspan=(0.0, 6.0),
                controls=[(1, t, rng.uniform(0.4, 0.7, 24)), (2, t, rng.uniform(0.4, 0.7, 24)),
                          (3, t, rng.uniform(0.1, 0.3, 24))],
                tunable_ic=[], quadrature_indices=[], shooting_points=[6.0 * k / n_intervals for k in range(n_intervals)])


def batchreactor_spec(n_intervals=4, rng=None):
    """batch reactor (.../batch_reactor.jl): Arrhenius terms exp(-c / u), temperature control in [298, 398]"""
    rng = np.random.default_rng(SEED0 + 13) if rng is None else rng
    t = np.arange(20) / 20.0
    return dict(model="batchreactor2", u0=[1.0, 0.0], p0=[298.0], tspan=(0.0, 1.0),
                controls=[(1, t, rng.uniform(298, 398, 20))], tunable_ic=[], quadrature_indices=[],
                shooting_points=[k / n_intervals for k in range(n_intervals)])


def lotkacomp_spec(n_intervals=4, rng=None):
    """competitive Lotka-Volterra with Lagrange state (.../lotka_competitive.jl)"""
    rng = np.random.default_rng(SEED0 + 14) if rng is None else rng
    t = np.arange(40) / 4.0
    return dict(model="lotkacomp3", u0=[0.5, 1.5, 0.0], p0=[0.5], tspan=(0.0, 10.0),
                controls=[(1, t, rng.uniform(0, 1, 40))], tunable_ic=[], quadrature_indices=[3],
                shooting_points=[10.0 * k / n_intervals for k in range(n_intervals)])


def f8_spec(n_intervals=4, rng=None):
    """F-8 aircraft, time-optimal: final time T tunable, control u (.../f8.jl), normalised time (0, 1)"""
    rng = np.random.default_rng(SEED0 + 15) if rng is None else rng
    t = np.arange(24) / 24.0
    return dict(model="f8aircraft4", u0=[0.4655, 0.0, 0.0, 0.0], p0=[0.5, 0.5], tspan=(0.0, 1.0),
                controls=[(2, t, rng.uniform(0, 1, 24))], tunable_ic=[], quadrature_indices=[],
                shooting_points=[k / n_intervals for k in range(n_intervals)])


def doubleosc_spec(n_intervals=3, rng=None):
    """double oscillator (.../double_oscillator.jl); explicit time dependence carried by the state tau' = 1"""
    rng = np.random.default_rng(SEED0 + 16) if rng is None else rng
    t = np.arange(30) / 5.0
    return dict(model="doubleosc6", u0=[0.0, 0.0, 0.0, 0.0, 0.0, 0.0], p0=[1.0], tspan=(0.0, 6.0),
                controls=[(1, t, rng.uniform(-1, 1, 30))], tunable_ic=[3, 4], quadrature_indices=[6],
                shooting_points=[6.0 * k / n_intervals for k in range(n_intervals)])


def layer_from_spec(spec, alg, **kw):
    """Build the SingleShootingLayer / MultipleShootingLayer a spec describes."""
    prob = ODEProblem(spec["model"], spec["u0"], spec["tspan"], spec["p0"], model_data=spec.get("model_data"),
                      kwargs=dict(spec.get("kwargs", {}),   # abstol / reltol of the adaptive mode ride here, as in the reference
                                  **({"saveat": list(spec["saveat"])} if spec.get("saveat") is not None else {})))
    names = spec.get("control_names") or [f"u{k + 1}" for k in range(len(spec["controls"]))]
    cb = spec.get("control_bounds") or [None] * len(spec["controls"])
    ctr = [(ci, ControlParameter(t, name=n, controls=(np.zeros(len(t)) if u is None else u),
                                 **({} if b is None else {"bounds": b})))
           for (ci, t, u), n, b in zip(spec["controls"], names, cb)]
    args = dict(controls=ctr, tunable_ic=spec.get("tunable_ic", []),
                quadrature_indices=spec.get("quadrature_indices", []))
    args.update(kw)
    if spec.get("shooting_points") is None:
        return SingleShootingLayer(prob, alg, **args)
    return MultipleShootingLayer(prob, alg, *spec["shooting_points"], **args)


def native_from_spec(spec, alg, device=0):
    """(layer, ps template, st, NativeLayer) for a spec; the Fisher block, if any, is registered."""
    lay = layer_from_spec(spec, alg, device=device)
    ps, st = lay.initialparameters(None), lay.initialstates(None)
    first = st if "tspans" in st else next(iter(st.values()))
    if spec.get("fim"):
        first["_fim"] = tuple(spec["fim"])
    return lay, ps, st, lay._native(st, ps)


# ---- the rest of lib/OptimalControlBenchmarks/src/problems: (model, u0, p0, tspan, [(p index, lo, hi) per control],
#      number of control values, quadrature_indices, objective row (1-based), perturbation scale of the parity tests)
#      initial values and horizons as in the problem files; control values are drawn uniformly inside (a part of) the
#      bounds the files state
BENCH_TABLE = {
    "brysondenham": ("brysondenham3", [0.0, 1.0, 0.0], [0.0], (0.0, 1.0), [(1, -2.0, 2.0)], 20, [3], 3, 0.05),
    "catalystmixing": ("catalystmixing2", [1.0, 0.0], [0.0], (0.0, 1.0), [(1, 0.0, 1.0)], 20, [], 1, 0.05),
    "cushioned": ("cushioned3", [2.0, 5.0, 0.0], [1.0, 0.0], (0.0, 10.0), [(2, -5.0, 5.0)], 20, [3], 3, 0.05),
    "denbigh": ("denbigh3", [1.0, 0.0, 0.0], [300.0], (0.0, 1000.0), [(1, 300.0, 330.0)], 100, [], 3, 0.005),   # fixed steps: rates <= 0.25
    "dielectrophoretic": ("dielectrophoretic3", [1.0, 0.0, 0.0], [1.0, 1.0], (0.0, 5.0), [(2, -1.0, 1.0)], 20, [3], 1, 0.05),
    "electriccar": ("electriccar4", [0.0, 0.0, 0.0, 0.0], [0.0], (0.0, 10.0), [(1, 0.0, 0.1)], 20, [4], 4, 0.01),
    "fuller": ("fuller3", [0.01, 0.0, 0.0], [0.5], (0.0, 1.0), [(1, 0.0, 1.0)], 20, [3], 3, 0.05),
    "jackson": ("jackson3", [1.0, 0.0, 0.0], [0.5], (0.0, 1.0), [(1, 0.0, 1.0)], 20, [], 3, 0.05),
    "mountaincar": ("mountaincar3", [-0.5, 0.0, 0.0], [1.0, 0.5], (0.0, 100.0), [(2, -1.0, 1.0)], 20, [3], 1, 0.02),
    "particlesteering": ("particlesteering5", [0.0, 0.0, 0.0, 0.0, 0.0], [3.0e-2, 0.0], (0.0, 10.0),
                         [(2, -1.5, 1.5)], 20, [5], 1, 0.02),
    "raomease": ("raomease2", [1.0, 0.0], [0.0], (0.0, 10.0), [(1, -1.0, 1.0)], 20, [2], 2, 0.05),
    "robbins": ("robbins4", [1.0, -2.0, 0.0, 0.0], [0.0], (0.0, 10.0), [(1, -1.0, 1.0)], 20, [4], 4, 0.05),
    "moonlanding": ("moonlanding3", [1.0, -0.783, 1.0], [1.0, 0.5], (0.0, 1.0), [(2, 0.0, 1.227)], 20, [], 3, 0.02),
    "lotkashared": ("lotkashared4", [1.5, 1.0, 0.5, 0.0], [0.5], (0.0, 40.0), [(1, 0.0, 1.0)], 40, [4], 4, 0.02),
    "hangingchain": ("hangingchain3", [1.0, 0.0, 0.0], [0.5], (0.0, 1.0), [(1, -1.0, 2.0)], 20, [], 2, 0.05),
    "tubularreactor": ("tubularreactor2", [1.0, 0.0], [0.0], (0.0, 1.0), [(1, 0.0, 5.0)], 20, [], 2, 0.05),
    "ocean": ("ocean4", [2.0e3, 1.0e4, 0.0, 0.0], [7.0, 7.0], (0.0, 400.0), [(1, 0.0, 40.0), (2, 0.0, 40.0)], 20, [4], 4, 0.02),
    "ductedfan": ("ductedfan7", [0.0] * 7, [0.5, 0.0, 8.5], (0.0, 10.0), [(2, -1.0, 1.0), (3, 0.0, 10.0)], 20, [7], 7, 0.02),
    "hangglider": ("hangglider4", [0.0, 1.0e3, 13.23, -1.288], [0.5, 0.5], (0.0, 100.0), [(2, 0.2, 1.2)], 20, [], 1, 0.01),
}


def bench_spec(name, n_intervals=4, rng=None):
    """(spec, objective row, perturbation scale) of one of the BENCH_TABLE problems on `n_intervals` shooting intervals"""
    model, u0, p0, tspan, ctr, n, quad, row, scale = BENCH_TABLE[name]
    rng = np.random.default_rng(SEED0 + 100 + sorted(BENCH_TABLE).index(name)) if rng is None else rng
    t = tspan[0] + (tspan[1] - tspan[0]) * np.arange(n) / n
    spec = dict(model=model, u0=list(u0), p0=list(p0), tspan=tspan,
                controls=[(ci, t, rng.uniform(lo, hi, n)) for ci, lo, hi in ctr],
                control_bounds=[(lo, hi) for _, lo, hi in ctr], tunable_ic=[], quadrature_indices=list(quad),
                shooting_points=[tspan[0] + (tspan[1] - tspan[0]) * k / n_intervals for k in range(n_intervals)])
    return spec, row, scale
