"""SingleShootingLayer -- host-side mirror of src/single_shooting.jl.

Static description and one-time table construction stay on the host (as in the reference); the hot loop
(_sequential_solve, :421-471) runs on the GPU through `native.NativeLayer`.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import native
from .controls import (ControlParameter, build_index_grid, collect_local_control_bounds, collect_local_controls,
                       collect_tspans, control_length, find_shooting_indices)
from .problem import MODEL_REGISTRY, ODEProblem
from .trajectory import Trajectory, retrieve_symbol_cache


def default_u0(rng, problem: ODEProblem, tunables, bounds):
    """single_shooting.jl:40-44 (tunables 1-based)"""
    lb, ub = bounds
    t = np.asarray(tunables, dtype=np.int64) - 1
    return np.clip(problem.u0[t], lb[t], ub[t])


def default_p0(rng, problem: ODEProblem, parameters, bounds):
    """single_shooting.jl:46-51"""
    t = np.asarray(parameters, dtype=np.int64) - 1
    return np.clip(problem.p[t], bounds[0], bounds[1])


class InitialConditionRemaker:
    """single_shooting.jl:255-270: u0_full = vcat(u0_fixed[constants], u_tunable)[sorting] (1-based tables)"""

    def __init__(self, sorting, constants):
        self.sorting = np.asarray(sorting, dtype=np.int64)
        self.constants = np.asarray(constants, dtype=np.int64)

    def __call__(self, u, u0):
        u = np.asarray(u, dtype=np.float64)
        if u.size == 0:
            return np.asarray(u0, dtype=np.float64)
        return np.concatenate([np.asarray(u0)[self.constants - 1], u])[self.sorting - 1]

    def statelength(self):
        return len(self.sorting)


def _setdiff(n, drop):
    drop = set(int(x) for x in drop)
    return [k for k in range(1, n + 1) if k not in drop]


def _flatten_chunks(x):
    """index_grid / tspans may be tuples of <=100-piece chunks (local_controls.jl:173-176); flatten."""
    if isinstance(x, tuple) and len(x) and isinstance(x[0], np.ndarray):
        return np.concatenate(x, axis=1)
    if isinstance(x, tuple) and len(x) and isinstance(x[0], tuple) and len(x[0]) and isinstance(x[0][0], tuple):
        return tuple(p for chunk in x for p in chunk)
    return x


class SingleShootingLayer:
    """single_shooting.jl:15-38, ctor :91-132.

    controls: sequence of (p_index_1based, ControlParameter) pairs -- `controls = (1 => control,)`.
    """

    def __init__(self, prob: ODEProblem, alg, *, controls: Sequence = (), tunable_ic: Sequence[int] = (),
                 bounds_ic=None, state_initialization=default_u0, bounds_p=None,
                 parameter_initialization=default_p0, quadrature_indices: Sequence[int] = (), device: int = 0,
                 **kwargs):
        self.problem = prob.remake(**kwargs) if kwargs else prob
        self.algorithm = alg
        controls = list(controls)
        self.control_indices = [int(c[0]) for c in controls]
        self.controls = [c[1] for c in controls]
        self.tunable_ic = [int(i) for i in tunable_ic]
        nx, npar = len(prob.u0), len(prob.p)
        self.tunable_p = [i for i in range(1, npar + 1) if i not in self.control_indices]   # setdiff (:109)
        p_vec = prob.p[np.asarray(self.tunable_p, dtype=np.int64) - 1]
        self.bounds_ic = ((np.full(nx, -np.inf), np.full(nx, np.inf)) if bounds_ic is None
                          else tuple(np.asarray(b, dtype=np.float64) for b in bounds_ic))
        self.bounds_p = ((np.full(len(p_vec), -np.inf), np.full(len(p_vec), np.inf)) if bounds_p is None
                         else tuple(np.asarray(b, dtype=np.float64) for b in bounds_p))
        self.state_initialization = state_initialization
        self.parameter_initialization = parameter_initialization
        self.quadrature_indices = [int(q) for q in quadrature_indices]
        self.device = device
        assert self.bounds_ic[0].shape == self.bounds_ic[1].shape == prob.u0.shape, \
            "The size of the initial states and its bounds is inconsistent."
        assert self.bounds_p[0].shape == self.bounds_p[1].shape == p_vec.shape, \
            "The size of the initial parameter vector and its bounds is inconsistent."
        assert all(1 <= q <= nx for q in self.quadrature_indices), \
            "Some quadrature indices are inconsistent with the state dimension"

    def __repr__(self):
        return f"SingleShootingLayer with {len(self.controls)} controls.\nUnderlying problem: {self.problem}"

    # ---- accessors (:134-144)
    def get_problem(self): return self.problem
    def get_controls(self): return self.controls, self.control_indices
    def get_tspan(self): return self.problem.tspan
    def get_tunable(self): return self.tunable_ic
    def get_params(self): return _setdiff(len(self.problem.p), self.control_indices)
    def get_quadrature_indices(self): return self.quadrature_indices

    # ---- bounds (:146-159)
    def get_bounds(self, shooting: bool = False, tspan=None):
        nx = len(self.problem.u0)
        state_indices = np.asarray(_setdiff(nx, self.quadrature_indices), dtype=np.int64) - 1
        sel = state_indices if shooting else np.asarray(self.tunable_ic, dtype=np.int64) - 1
        b_ic = tuple(b[sel] for b in self.bounds_ic)
        if self.controls:
            clb, cub = collect_local_control_bounds(*self.controls,
                                                    tspan=self.problem.tspan if tspan is None else tspan)
        else:
            clb = cub = np.zeros(0)
        return (dict(u0=b_ic[0], p=self.bounds_p[0], controls=clb),
                dict(u0=b_ic[1], p=self.bounds_p[1], controls=cub))

    # ---- LuxCore triple
    def initialparameters(self, rng, *, tspan=None, u0=None, shooting_layer: bool = False):
        """__initialparameters (:170-201): (u0, p, controls) in this field order."""
        tspan = self.problem.tspan if tspan is None else tspan
        problem = self.problem.remake(tspan=tspan, **({} if u0 is None else {"u0": u0}))
        nx = len(problem.u0)
        tun = _setdiff(nx, self.quadrature_indices) if shooting_layer else self.tunable_ic
        return dict(
            u0=np.asarray(self.state_initialization(rng, problem, tun, self.bounds_ic), dtype=np.float64),
            p=np.asarray(self.parameter_initialization(rng, problem, self.tunable_p, self.bounds_p), dtype=np.float64),
            controls=(collect_local_controls(rng, *self.controls, tspan=tspan) if self.controls
                      else np.zeros(0)),
        )

    def parameterlength(self, *, tspan=None, shooting_layer: bool = False) -> int:
        """__parameterlength (:203-215)"""
        tspan = self.problem.tspan if tspan is None else tspan
        nx = len(self.problem.u0)
        N = nx - len(self.quadrature_indices) if shooting_layer else len(self.tunable_ic)
        N += len(self.tunable_p)
        N += sum(control_length(c, tspan=tspan) for c in self.controls)
        return N

    def initialstates(self, rng, *, tspan=None, shooting_layer: bool = False):
        """__initialstates (:274-352): the static tables the hot loop consumes."""
        tspan = self.problem.tspan if tspan is None else tspan
        prob = self.problem
        nx, npar = len(prob.u0), len(prob.p)
        if not shooting_layer:
            constants = _setdiff(nx, self.tunable_ic)
            sorting = np.argsort(np.asarray(constants + self.tunable_ic), kind="stable") + 1   # sortperm
        else:
            constants = list(self.quadrature_indices)
            sorting = np.argsort(np.asarray(constants + _setdiff(nx, constants)), kind="stable") + 1
        initial_condition = InitialConditionRemaker(sorting, constants)
        # controls which do not act on the dynamics are filtered (:301-304)
        active = [ci <= npar for ci in self.control_indices]
        control_indices = [ci for ci, a in zip(self.control_indices, active) if a]
        controls = [c for c, a in zip(self.controls, active) if a]
        # scatter matrices A, B (:306-318): column ids handed out in ascending p index
        A = np.zeros((npar, npar - len(control_indices)), dtype=bool)
        Bm = np.zeros((npar, len(control_indices)), dtype=bool)
        pid = cid = 0
        for i in range(1, npar + 1):
            if i in control_indices:
                Bm[i - 1, cid] = True
                cid += 1
            else:
                A[i - 1, pid] = True
                pid += 1
        if controls:
            grid = build_index_grid(*controls, tspan=tspan, subdivide=100)
            tspans = collect_tspans(*controls, tspan=tspan, subdivide=100)
        else:
            grid = np.asarray(control_indices, dtype=np.int64)
            tspans = (prob.tspan,)   # (:331) the problem's span, also on sub-intervals -- reference behaviour kept
        sidx = np.zeros(nx + len(controls), dtype=bool)
        if shooting_layer:
            sidx[np.asarray(_setdiff(nx, self.quadrature_indices), dtype=np.int64) - 1] = True
            first = _flatten_chunks(tspans)[0]
            for i, c in enumerate(controls):
                sidx[nx + i] = not find_shooting_indices(first, c)
        shooting_indices = np.nonzero(sidx)[0].astype(np.int64) + 1
        order = np.argsort(np.asarray(control_indices), kind="stable") if control_indices else []
        control_names = [controls[i].name for i in order]
        symcache = retrieve_symbol_cache(nx, npar, control_indices, control_names)
        if isinstance(grid, tuple):
            act = sorted(set(int(v) for g in grid for row in g for v in np.unique(row)))
        else:
            act = [np.unique(row) for row in np.atleast_2d(grid)] if len(controls) else []
        return dict(initial_condition=initial_condition, index_grid=grid, tspans=tspans,
                    parameter_matrices=(A, Bm), symcache=symcache, shooting_indices=shooting_indices,
                    active_controls=act, _control_indices=control_indices, _shooting_layer=shooting_layer,
                    _tspan=tspan)

    # ---- native handle
    def _interval_desc(self, st, ps_lengths, is_shooting=None):
        grid = _flatten_chunks(st["index_grid"])
        tspans = _flatten_chunks(st["tspans"])
        nact = len(st["_control_indices"])
        ig = np.asarray(grid, dtype=np.int64).reshape(nact, -1) if nact else np.zeros((0, len(tspans)), dtype=np.int64)
        return dict(tspans=tspans, index_grid=ig, n_u0=ps_lengths[0], n_c=ps_lengths[2],
                    is_shooting=st["_shooting_layer"] if is_shooting is None else is_shooting, sorting=st["initial_condition"].sorting,
                    constants=st["initial_condition"].constants, shooting_indices=st["shooting_indices"])

    def _make_native(self, interval_sts, interval_lengths, fim=(0, 0), is_shooting=None):
        prob = self.problem
        model_id, nx, npar = MODEL_REGISTRY[prob.model]
        cidx = interval_sts[0]["_control_indices"]
        return native.NativeLayer(
            model_id=model_id, method=self.algorithm.method_id, steps_per_piece=self.algorithm.steps_per_piece,
            nx=nx, np_all=npar, u0=prob.u0, tunable_p=[i for i in range(1, npar + 1) if i not in cidx],
            control_indices=cidx, quadrature_indices=self.quadrature_indices,
            intervals=[self._interval_desc(s, l, is_shooting) for s, l in zip(interval_sts, interval_lengths)],
            device=self.device, model_data=prob.model_data,
            abstol=float(prob.kwargs.get("abstol", 1e-6)), reltol=float(prob.kwargs.get("reltol", 1e-3)),
            saveat=prob.kwargs.get("saveat"),
            **(dict(fim_offset=fim["offset"], fim_n=fim["n"], fim_mode=fim["mode"], fim_base_nx=fim["base_nx"],
                    n_observed=fim["n_observed"], observation_rows=fim["rows"]) if isinstance(fim, dict)
               else dict(fim_offset=fim[0], fim_n=fim[1])))

    def _native(self, st, ps, shooting_layer: bool = False):
        # fixval is problem.u0 unless the caller passes shooting_layer = true (:359); one handle per flag
        key = "_native_ss1" if shooting_layer else "_native_ss0"
        if key not in st:
            lengths = (len(ps["u0"]), len(ps["p"]), len(ps["controls"]))
            st[key] = self._make_native([st], [lengths], fim=st.get("_fim", (0, 0)), is_shooting=shooting_layer)
        return st[key]

    # ---- evaluation (:357-374)
    def __call__(self, u0, ps, st, shooting_layer: bool = False):
        """layer(nothing, ps, st) -> (Trajectory, st).  Passing an explicit u0 overrides the tunable
        initial condition as in the reference (:364); it must then be the full state."""
        if u0 is not None:
            raise NotImplementedError("explicit u0 override: put the value into ps.u0 / tunable_ic instead")
        nat = self._native(st, ps, shooting_layer)
        vec = np.concatenate([np.asarray(ps["u0"], float), np.asarray(ps["p"], float), np.asarray(ps["controls"], float)])
        r = nat.eval_host(vec[None, :], native.WANT_TRAJ)
        t = nat.sample_times()   # piece starts / saveat times / piece ends (`saveat` in problem.kwargs, oed.jl:106-113)
        u = r["traj"][0].reshape(nat.n_samples, nat.n_row)
        return Trajectory(st["symcache"], u, np.asarray(ps["p"], float), t, {}, np.zeros(0, dtype=np.int64)), st

    def get_block_structure(self, tspan=None):
        """single_shooting.jl:502-506"""
        return np.array([0, self.parameterlength(tspan=tspan)], dtype=np.int64)


def to_vec(ps) -> np.ndarray:
    """ComponentArray(ps) |> collect (ext/CorleoneComponentArraysExtension.jl:15-23): intervals in order,
    fields (u0, p, controls) in that order."""
    if "u0" in ps:
        return np.concatenate([np.asarray(ps[k], dtype=np.float64).ravel() for k in ("u0", "p", "controls")])
    return np.concatenate([to_vec(v) for v in ps.values()])


def from_vec(vec, like):
    """Inverse of to_vec given a template ps (the ComponentArray axes)."""
    vec = np.asarray(vec, dtype=np.float64)
    off = 0

    def take(tpl):
        nonlocal off
        if "u0" in tpl:
            out = {}
            for k in ("u0", "p", "controls"):
                n = len(tpl[k])
                out[k] = vec[off:off + n].copy()
                off += n
            return out
        return {k: take(v) for k, v in tpl.items()}

    return take(like)
