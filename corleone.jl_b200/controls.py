"""Piecewise-constant control parameterisation -- host-side mirror of src/local_controls.jl.

Setup-time integer/table logic only (runs once per layer, on the host, as in the reference); the tables it
produces are uploaded to the GPU by `native.NativeLayer`.  Index tables are kept exactly as Julia holds
them: int64, 1-based.
"""
from __future__ import annotations

import itertools
from typing import Callable, Sequence

import numpy as np

_gensym = itertools.count(1)
INT64_MAX = np.iinfo(np.int64).max


def default_u(rng, t, bounds):
    """local_controls.jl:19"""
    return np.zeros(np.shape(t), dtype=np.float64)


def default_bounds(t):
    """local_controls.jl:20"""
    t = np.asarray(t, dtype=np.float64)
    return np.full(t.shape, -np.inf), np.full(t.shape, np.inf)


class ControlParameter:
    """Piecewise constant control discretisation (local_controls.jl:8-17, ctor :49-51).

    t        time points at which discretised variables are introduced
    controls initial values: a vector, or a function (rng, t, bounds) -> u
    bounds   (lb, ub) scalars, (lb, ub) vectors, or a function t -> (lb, ub)
    """

    def __init__(self, t, name: str | None = None, controls=default_u, bounds=default_bounds):
        self.name = name if name is not None else f"##w#{next(_gensym)}"
        self.t = np.array(t, dtype=np.float64).ravel()
        self.controls = controls if callable(controls) else np.array(controls, dtype=np.float64).ravel()
        if callable(bounds):
            self.bounds = bounds
        else:
            lb, ub = bounds
            self.bounds = (np.atleast_1d(np.asarray(lb, dtype=np.float64)), np.atleast_1d(np.asarray(ub, dtype=np.float64)))

    def __repr__(self):
        return f"ControlParameter({self.name!r}, {len(self.t)} points)"


_SORTED = {}   # id(array) -> (array, is it non-decreasing?): a layer asks for the same grid once per shooting interval


def _is_sorted(t: np.ndarray) -> bool:
    hit = _SORTED.get(id(t))
    if hit is None or hit[0] is not t:
        if len(_SORTED) > 256:
            _SORTED.clear()
        hit = _SORTED[id(t)] = (t, bool(np.all(t[1:] >= t[:-1])) and not np.isnan(t).any())
    return hit[1]


def _restrict(t: np.ndarray, tspan) -> np.ndarray:
    """findall(tspan[1] .<= t .< tspan[2])  (0-based positions; local_controls.jl:55).  On a sorted grid the positions
    are one contiguous run found by bisection -- a layer with I intervals would otherwise scan the grid I times."""
    if tspan is None:
        return np.arange(len(t))
    if len(t) > 64 and _is_sorted(t):
        return np.arange(np.searchsorted(t, tspan[0], side="left"), np.searchsorted(t, tspan[1], side="left"))
    return np.nonzero((tspan[0] <= t) & (t < tspan[1]))[0]


def get_timegrid(c: ControlParameter, tspan=(-np.inf, np.inf)) -> np.ndarray:
    """local_controls.jl:53-57"""
    return c.t[_restrict(c.t, tspan)]


def control_length(c: ControlParameter, tspan=None) -> int:
    """local_controls.jl:64-68"""
    return int(len(_restrict(c.t, tspan)))


def get_bounds(c: ControlParameter, tspan=None):
    """local_controls.jl:94-107"""
    if callable(c.bounds):
        return c.bounds(get_timegrid(c, (-np.inf, np.inf) if tspan is None else tspan))
    idx = _restrict(c.t, tspan)
    lb, ub = c.bounds
    if len(lb) == len(ub) == 1:
        return np.repeat(lb, len(idx)), np.repeat(ub, len(idx))
    if len(lb) == len(ub) == len(c.t):
        return lb[idx], ub[idx]
    raise ValueError(f"Incompatible control bound definition. Got {len(lb)} elements, expected {len(c.t)}.")


def get_controls(rng, c: ControlParameter, raw: bool = False, tspan=None) -> np.ndarray:
    """local_controls.jl:75-87"""
    if callable(c.controls):
        bounds = get_bounds(c, tspan=tspan)
        return np.asarray(c.controls(rng, c.t[_restrict(c.t, tspan)], bounds), dtype=np.float64)
    if raw:
        return c.controls
    return c.controls[_restrict(c.t, tspan)]


def check_consistency(rng, c: ControlParameter) -> None:
    """local_controls.jl:109-118 (same messages)"""
    grid = get_timegrid(c)
    u = get_controls(rng, c, raw=True)
    lb, ub = get_bounds(c)
    assert np.all(np.diff(grid) >= 0), "Time grid is not sorted."
    assert len(np.unique(grid)) == len(grid), "Time grid is not unique."
    if np.shape(lb) == np.shape(ub):
        assert np.all(lb <= ub), "Bounds are inconsistent"
    assert np.shape(lb) == np.shape(ub) == np.shape(u) == np.shape(grid), "Sizes are inconsistent"
    assert np.all((lb <= u) & (u <= ub)), "Initial values are inconsistent"


def get_subvector_indices(M: int, L: int):
    """local_controls.jl:120-147 -> list of 1-based inclusive (start, stop) ranges"""
    if M < 0 or L <= 0:
        raise ValueError("M must be a non-negative integer and L must be a positive integer.")
    N = M // L
    out = [((i - 1) * L + 1, i * L) for i in range(1, N + 1)]
    if M - N * L > 0:
        out.append((N * L + 1, M))
    return out


def _merged_grid(controls: Sequence[ControlParameter], tspan):
    ts = [get_timegrid(c, tspan) for c in controls]
    allp = np.concatenate(ts + [np.array(tspan, dtype=np.float64)])
    grid = np.unique(allp)  # sort! |> unique!
    return ts, grid[np.isfinite(grid)]


def build_index_grid(*controls: ControlParameter, offset: bool = True, tspan=(-np.inf, np.inf),
                     subdivide: int = INT64_MAX):
    """local_controls.jl:149-178.  Returns an int64 (n_controls, M) matrix of 1-based indices into the
    interval's flat local control vector, or a tuple of column chunks when M > subdivide."""
    ts, grid = _merged_grid(controls, tspan)
    M = len(grid) - 1
    idx = np.zeros((len(ts), M), dtype=np.int64)
    for i, ti in enumerate(ts):
        k = np.searchsorted(ti, grid[:M], side="right").astype(np.int64)  # searchsortedlast
        lo, hi = 1, len(ti)
        idx[i] = np.where(k > hi, hi, np.where(k < lo, lo, k))             # Julia clamp(x, lo, hi)
    if offset:
        for i in range(1, len(ts)):
            idx[i] += idx[i - 1].max() if M > 0 else 0
    if idx.size:
        idx -= idx.min() - 1
    if M > subdivide:
        return tuple(idx[:, a - 1:b] for a, b in get_subvector_indices(M, subdivide))
    return idx


def find_shooting_indices(tspan, control: ControlParameter) -> bool:
    """local_controls.jl:180: any(tspan[1] .== control.t) -- by bisection on a sorted grid"""
    t = control.t
    if len(t) > 64 and _is_sorted(t):
        k = int(np.searchsorted(t, tspan[0], side="left"))
        return bool(k < len(t) and t[k] == tspan[0])
    return bool(np.any(tspan[0] == t))


def collect_tspans(*controls: ControlParameter, tspan=(-np.inf, np.inf), subdivide: int = INT64_MAX):
    """local_controls.jl:183-195 -> tuple of (t0, t1) pairs, or a tuple of such tuples when chunked"""
    _, grid = _merged_grid(controls, tspan)
    full = tuple((float(a), float(b)) for a, b in zip(grid[:-1], grid[1:]))
    if len(full) > subdivide:
        return tuple(full[a - 1:b] for a, b in get_subvector_indices(len(full), subdivide))
    return full


def collect_local_controls(rng, *controls: ControlParameter, **kw) -> np.ndarray:
    """local_controls.jl:201-207"""
    return np.concatenate([get_controls(rng, c, **kw) for c in controls]) if controls else np.zeros(0)


def collect_local_control_bounds(*controls: ControlParameter, **kw):
    """local_controls.jl:213-220"""
    b = [get_bounds(c, **kw) for c in controls]
    return np.concatenate([x[0] for x in b]), np.concatenate([x[1] for x in b])
