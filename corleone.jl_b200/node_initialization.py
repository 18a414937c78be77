"""Node initialisation schemes -- mirror of src/node_initialization.jl:13-255.

All take (rng, shooting; ps=default, ...) and return the nested ps {interval_i: {u0, p, controls}}.
`forward_initialization` is a consumer of the hot path: it integrates interval after interval on the GPU.
"""
from __future__ import annotations

import copy

import numpy as np

from .multiple_shooting import MultipleShootingLayer, default_initialization, get_bounds


def _random_value(rng, lb, ub):
    """src/Corleone.jl:37"""
    lb, ub = np.asarray(lb, float), np.asarray(ub, float)
    return lb + rng.random(lb.shape) * (ub - lb)


def random_initialization(rng, shooting: MultipleShootingLayer, ps=None, **kw):
    """node_initialization.jl:13-33"""
    ps = default_initialization(rng, shooting) if ps is None else ps
    tic = np.asarray(shooting.layer.tunable_ic, dtype=np.int64) - 1
    keys = list(ps.keys())
    if len(ps[keys[-1]]["u0"]) == 0:
        return ps
    lb, ub = get_bounds(shooting)
    out = {}
    for i, k in enumerate(keys):
        if i == 0:
            # the reference indexes the already restricted bounds with tunable_ic again (:26)
            unew = _random_value(rng, lb[k]["u0"][tic], ub[k]["u0"][tic])
        else:
            unew = _random_value(rng, lb[k]["u0"], ub[k]["u0"])
        out[k] = dict(ps[k], u0=unew)
    return out


def linear_initialization(rng, shooting: MultipleShootingLayer, ps=None, u_infinity=None, **kw):
    """node_initialization.jl:48-69; note the division by t_inf, not (t_inf - t0) (:65)"""
    ps = default_initialization(rng, shooting) if ps is None else ps
    keys = list(ps.keys())
    u_infinity = ps[keys[-1]]["u0"] if u_infinity is None else np.asarray(u_infinity, float)
    if len(u_infinity) == 0:
        return ps
    problem = shooting.layer.problem
    tic = np.asarray(shooting.layer.tunable_ic, dtype=np.int64) - 1
    u0 = ps[keys[0]]["u0"]
    if len(u0) == 0:
        u0 = problem.u0
    t0, tinf = problem.tspan
    for i, k in enumerate(keys):
        t = shooting.shooting_intervals[i][0]
        u0new = (u_infinity - u0) * (t - t0) / tinf + u0
        ps[k]["u0"][...] = u0new if i > 0 else u0new[tic]
    return ps


def _setu0(u0, unew, tunable_ics):
    """__setu0! (:71-88); tunable_ics 1-based"""
    if isinstance(unew, dict):
        return np.array([unew.get(i, u0[i - 1]) for i in tunable_ics], dtype=np.float64)
    if isinstance(unew, (list, tuple)) and len(unew) and isinstance(unew[0], tuple):
        d = dict(unew)
        return np.array([d[i] for i in tunable_ics], dtype=np.float64)   # missing index errors, as in the reference
    if isinstance(unew, (np.ndarray, list, tuple)):
        return np.asarray(unew, dtype=np.float64)[np.asarray(tunable_ics, dtype=np.int64) - 1]
    return u0


def custom_initialization(rng, shooting: MultipleShootingLayer, ps=None, u0s=(), **kw):
    """node_initialization.jl:121-137"""
    ps = default_initialization(rng, shooting) if ps is None else ps
    if len(u0s) == 0:
        return ps
    tic = shooting.layer.tunable_ic
    for i, k in enumerate(ps.keys()):
        if len(u0s) > i:
            sel = list(range(1, len(ps[k]["u0"]) + 1)) if i > 0 else tic
            ps[k]["u0"][...] = _setu0(ps[k]["u0"], u0s[i], sel)
    return ps


def forward_initialization(rng, shooting: MultipleShootingLayer, ps=None, fixed_indices=(), **kw):
    """node_initialization.jl:152-170: sequential forward sweep through the intervals; intervals listed in
    fixed_indices (1-based) keep their node.  The sweep runs on the GPU (cb200_forward_init): one thread per
    candidate walks the intervals, calling the layer without the shooting flag as the reference does (:165)."""
    from .single_shooting import from_vec, to_vec
    ps = default_initialization(rng, shooting) if ps is None else ps
    st = shooting.initialstates(rng)
    nat = shooting._native(st, ps)
    new = from_vec(nat.forward_init(to_vec(ps)[None, :], fixed_indices)[0], ps)
    for k in ps:   # in place, like the reference (sps.u0 .= ...)
        ps[k]["u0"][...] = new[k]["u0"]
    return ps


def forward_initialization_batch(shooting: MultipleShootingLayer, ps_batch, ps_like, st=None, fixed_indices=()):
    """The batch axis the reference does not have: ps_batch is (B, nvars), one flat candidate per row; returns the
    forward-initialised copy (only node values change)."""
    st = shooting.initialstates(None) if st is None else st
    return shooting._native(st, ps_like).forward_init(ps_batch, fixed_indices)


def constant_initialization(rng, shooting: MultipleShootingLayer, ps=None, u0=None, **kw):
    """node_initialization.jl:202-216"""
    ps = default_initialization(rng, shooting) if ps is None else ps
    if u0 is None or len(u0) == 0:
        return ps
    tic = shooting.layer.tunable_ic
    for i, k in enumerate(ps.keys()):
        sel = list(range(1, len(ps[k]["u0"]) + 1)) if i > 0 else tic
        ps[k]["u0"][...] = _setu0(ps[k]["u0"], u0, sel)
    return ps


def hybrid_initialization(rng, shooting: MultipleShootingLayer, *f, ps=None, **kw):
    """node_initialization.jl:234-255: f = (intervals, method) pairs applied in sequence; only the listed
    intervals (1-based) take the method's result."""
    ps = default_initialization(rng, shooting) if ps is None else ps
    for intervals, method in f:
        intervals = [intervals] if np.isscalar(intervals) else list(intervals)
        pnew = method(rng, shooting, ps=copy.deepcopy(ps), **kw)
        keys = list(ps.keys())
        # NamedTuple{names}(pnew) takes the FIRST length(intervals) entries of pnew under the listed names (:246-247)
        pvals = list(pnew.values())
        ps = dict(ps)
        for n, iv in enumerate(intervals):
            ps[keys[iv - 1]] = pvals[n]
    return ps
