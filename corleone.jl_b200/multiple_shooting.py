"""MultipleShootingLayer -- host-side mirror of src/multiple_shooting.jl.

The per-interval map (mythreadmap over intervals, src/Corleone.jl:22-24, multiple_shooting.jl:118-130), the
merge with shooting defects and quadrature sums (:139-182) and both constraint orderings (:229-310) all run
on the GPU in one cb200_eval; this module keeps the reference's layer API around that call and adds the
batch axis the reference does not have (`evaluate`).
"""
from __future__ import annotations

import numpy as np

from . import native
from .single_shooting import SingleShootingLayer, _flatten_chunks, to_vec
from .trajectory import Trajectory


def default_initialization(rng, shooting: "MultipleShootingLayer"):
    """multiple_shooting.jl:45-55"""
    return {f"interval_{i + 1}": shooting.layer.initialparameters(rng, tspan=ts, shooting_layer=(i != 0))
            for i, ts in enumerate(shooting.shooting_intervals)}


class MultipleShootingLayer:
    """multiple_shooting.jl:13-22, ctors :63-96.

    MultipleShootingLayer(prob, alg, t1, t2, ...; controls=..., initialization=...) or
    MultipleShootingLayer(layer, t1, t2, ...)
    """

    def __init__(self, prob_or_layer, *args, ensemble_alg=None, initialization=default_initialization, **kwargs):
        if isinstance(prob_or_layer, SingleShootingLayer):
            layer, tpoints = prob_or_layer, args
        else:
            alg, tpoints = args[0], args[1:]
            layer = SingleShootingLayer(prob_or_layer, alg, **kwargs)
        if len(tpoints) == 1 and np.ndim(tpoints[0]) == 1:
            tpoints = tuple(tpoints[0])
        pts = np.unique(np.concatenate([np.asarray(tpoints, dtype=np.float64), np.asarray(layer.problem.tspan)]))
        self.layer = layer
        self.shooting_intervals = tuple((float(a), float(b)) for a, b in zip(pts[:-1], pts[1:]))   # (:86-90)
        self.ensemble_alg = ensemble_alg   # kept for API parity; every interval runs in one GPU launch
        self.initialization = initialization

    def __repr__(self):
        return (f"MultipleShootingLayer with {len(self.shooting_intervals)} shooting intervals and "
                f"{len(self.layer.controls)} controls.\nUnderlying problem: {self.layer!r}")

    def get_quadrature_indices(self):
        return self.layer.get_quadrature_indices()

    # ---- LuxCore triple (:98-116)
    def initialparameters(self, rng):
        return self.initialization(rng, self)

    def parameterlength(self) -> int:
        return int(get_block_structure(self)[-1])

    def initialstates(self, rng):
        return {f"interval_{i + 1}": self.layer.initialstates(rng, tspan=ts, shooting_layer=(i != 0))
                for i, ts in enumerate(self.shooting_intervals)}

    # ---- native handle, cached in the (mutable) state of the first interval
    def _native(self, st, ps) -> native.NativeLayer:
        first = next(iter(st.values()))
        if "_native" not in first:
            keys = list(st.keys())
            lengths = [(len(ps[k]["u0"]), len(ps[k]["p"]), len(ps[k]["controls"])) for k in keys]
            first["_native"] = self.layer._make_native([st[k] for k in keys], lengths, fim=first.get("_fim", (0, 0)))
        return first["_native"]

    def _times(self, st):
        keys = list(st.keys())
        t = []
        for n, k in enumerate(keys):
            tsp = _flatten_chunks(st[k]["tspans"])
            ti = [tsp[0][0]] + [b for _, b in tsp]
            t += ti if n == len(keys) - 1 else ti[:-1]   # (:145-147)
        return np.array(t)

    # ---- evaluation (:132-182)
    def __call__(self, u0, ps, st):
        """shooting(nothing, ps, st) -> (Trajectory, st) for one candidate."""
        if len(st) == 1:
            k = next(iter(st))
            return self.layer(u0, ps[k], st[k])   # size(u, 1) == 1 && return only(u)  (:140)
        nat = self._native(st, ps)
        r = nat.eval_host(to_vec(ps)[None, :], native.WANT_TRAJ | native.WANT_DEFECTS)
        u = r["traj"][0].reshape(nat.n_samples, nat.n_row)
        keys = list(st.keys())
        nx = len(self.layer.problem.u0)
        # split the stage-ordered defects back into the (u0, p, controls) NamedTuple per stage (:149-169)
        shooting = {keys[0]: dict(u0=np.zeros(0), p=np.zeros(0), controls=np.zeros(0))}
        ds, o = r["defects_stage"][0], 0
        for k in keys[1:]:
            sidx = st[k]["shooting_indices"]
            nu, nc, npar = int(np.sum(sidx <= nx)), int(np.sum(sidx > nx)), len(ps[k]["p"])
            shooting[k] = dict(u0=ds[o:o + nu], p=ds[o + nu:o + nu + npar],
                               controls=ds[o + nu + npar:o + nu + npar + nc])
            o += nu + npar + nc
        # offsets = cumsum(lastindex(us[i])) of the *unmerged* pieces (:148)
        offsets = np.cumsum([len(_flatten_chunks(st[k]["tspans"])) + 1 for k in keys[:-1]]).astype(np.int64)
        traj = Trajectory(st[keys[0]]["symcache"], u, np.asarray(ps[keys[0]]["p"], float), nat.sample_times(),
                          shooting, offsets)
        return traj, st

    def evaluate(self, ps_batch, st, ps_like, want=native.WANT_DEFECTS | native.WANT_OBJ, objective_row=1,
                 fim_init=None):
        """Batched evaluation (new axis): ps_batch is (B, nvars), one flat candidate per row, in
        ComponentArray order.  Returns a dict of (B, n) arrays."""
        nat = self._native(st, ps_like)
        return nat.eval_host(ps_batch, want, objective_row, fim_init)


def sample_pieces(layer, sts):
    """Per interval: the control piece (0-based) every sample of the MERGED trajectory belongs to -- the start sample and
    the samples of piece 0 carry piece 0's controls, saveat samples inside a piece that piece's (single_shooting.jl:456;
    `saveat` in problem.kwargs, lib/CorleoneOED/src/oed.jl:106-113); the end sample of all but the last interval is dropped."""
    saveat = layer.problem.kwargs.get("saveat") or []
    out = []
    for i, sti in enumerate(sts):
        pieces = [0]
        for j, (a, bnd) in enumerate(_flatten_chunks(sti["tspans"])):
            pieces += [j] * (1 + sum(1 for s in saveat if a < s < bnd))
        out.append(np.asarray(pieces if i == len(sts) - 1 else pieces[:-1], dtype=np.int64))
    return out


def trajectory_jacobian(shooting, st, ps_like, result, b: int = 0) -> np.ndarray:
    """Dense d(merged trajectory samples)/d(flat variables) of candidate `b`, shape (n_row, n_samples, nvars) --
    what ForwardDiff through `traj` yields and the point constraints of src/dynprob.jl:30-58,89-103 differentiate.

    Assembled from one cb200_eval with WANT_TRAJ_SENS | WANT_SENS: `traj_sens` holds the per-interval blocks at
    every saved sample; quadrature rows additionally carry the running sums of the earlier intervals
    (src/multiple_shooting.jl:171-177), whose derivatives are the end-of-interval blocks of `sens`; a control row
    is the control value of the piece that ends at the sample (src/single_shooting.jl:456)."""
    single = isinstance(shooting, SingleShootingLayer)
    layer = shooting if single else shooting.layer
    sts = [st] if single else list(st.values())
    pss = [ps_like] if single else list(ps_like.values())
    nat = shooting._native(st, ps_like)
    nx, nrow, I = nat.nx, nat.n_row, nat.I
    blocks = nat.block_structure()
    quad = [q - 1 for q in layer.quadrature_indices]
    ts, se = result["traj_sens"][b], result["sens"][b]
    J = np.zeros((nrow, nat.n_samples, nat.nvars))
    so = toff = soff = 0
    qacc = np.zeros((len(quad), nat.nvars))          # d(running quadrature sum)/d(variables of earlier intervals)
    pieces = sample_pieces(layer, sts)
    for i in range(I):
        ns = int(blocks[i + 1] - blocks[i])
        grid = np.asarray(_flatten_chunks(sts[i]["index_grid"]), dtype=np.int64)
        nact = len(sts[i]["_control_indices"])
        grid = grid.reshape(nact, -1) if nact else np.zeros((0, len(_flatten_chunks(sts[i]["tspans"]))), dtype=np.int64)
        nsmp = len(pieces[i])
        c0 = blocks[i] + len(pss[i]["u0"]) + len(pss[i]["p"])
        blk = ts[toff:toff + nsmp * ns * nx].reshape(nsmp, ns, nx)
        for j in range(nsmp):
            J[:nx, so + j, blocks[i]:blocks[i + 1]] = blk[j].T
            for qi, q in enumerate(quad):
                J[q, so + j, :] += qacc[qi]
            piece = int(pieces[i][j])
            for r in range(nact):
                J[nx + r, so + j, c0 + grid[r, piece] - 1] = 1.0
        send = se[soff:soff + nx * ns].reshape(ns, nx)
        for qi, q in enumerate(quad):
            qacc[qi, blocks[i]:blocks[i + 1]] += send[:, q]
        so += nsmp - (1 if i == I - 1 else 0)
        toff += nsmp * ns * nx
        soff += nx * ns
    return J


# ---- constraint counts (:184-218)
def get_number_of_state_matchings(shooting, ps=None, st=None, rng=None):
    st = shooting.initialstates(rng) if st is None else st
    return int(sum(np.sum(s["shooting_indices"] <= s["initial_condition"].statelength())
                   for s in list(st.values())[1:]))


def get_number_of_parameter_matchings(shooting, ps=None, st=None, rng=None):
    ps = shooting.initialparameters(rng) if ps is None else ps
    return int(sum(len(p["p"]) for p in list(ps.values())[:-1]))


def get_number_of_control_matchings(shooting, ps=None, st=None, rng=None):
    st = shooting.initialstates(rng) if st is None else st
    return int(sum(np.sum(s["shooting_indices"] > s["initial_condition"].statelength())
                   for s in list(st.values())[1:]))


def get_number_of_shooting_constraints(layer, ps=None, st=None, rng=None):
    if isinstance(layer, SingleShootingLayer):
        return 0
    return (get_number_of_state_matchings(layer, ps, st, rng) + get_number_of_control_matchings(layer, ps, st, rng)
            + get_number_of_parameter_matchings(layer, ps, st, rng))


# ---- constraint orderings (:220-310)
def stage_ordered_shooting_constraints(traj: Trajectory) -> np.ndarray:
    """sorted by shooting stage, per stage states - parameters - controls (:229)"""
    parts = [v[k] for v in traj.shooting.values() for k in ("u0", "p", "controls")]
    return np.concatenate(parts) if parts else np.zeros(0)


def stage_ordered_shooting_constraints_(res, traj: Trajectory):
    v = stage_ordered_shooting_constraints(traj)
    res[:len(v)] = v
    return res


def _matchings(traj: Trajectory, kind: str) -> np.ndarray:
    parts = [v[kind] for v in traj.shooting.values()]
    return np.concatenate(parts) if parts else np.zeros(0)


def state_matchings(traj): return _matchings(traj, "u0")
def parameter_matchings(traj): return _matchings(traj, "p")
def control_matchings(traj): return _matchings(traj, "controls")


def shooting_constraints(traj: Trajectory) -> np.ndarray:
    """sorted by kind states - parameters - controls, per kind by stage (:281)"""
    if not traj.shooting:
        return np.zeros(0)
    return np.concatenate([_matchings(traj, k) for k in ("u0", "p", "controls")])


def shooting_constraints_(res, traj: Trajectory):
    """in-place shooting_constraints! (:289-295); writes the leading entries of res"""
    v = shooting_constraints(traj)
    res[:len(v)] = v
    return res


# ---- structure / bounds (:320-341)
def get_block_structure(layer) -> np.ndarray:
    if isinstance(layer, SingleShootingLayer):
        return layer.get_block_structure()
    lens = [layer.layer.parameterlength(tspan=ts, shooting_layer=(i > 0))
            for i, ts in enumerate(layer.shooting_intervals)]
    return np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)


def get_bounds(layer):
    if isinstance(layer, SingleShootingLayer):
        return layer.get_bounds()
    b = [layer.layer.get_bounds(tspan=ts, shooting=(i > 0)) for i, ts in enumerate(layer.shooting_intervals)]
    names = [f"interval_{i + 1}" for i in range(len(b))]
    return ({n: x[0] for n, x in zip(names, b)}, {n: x[1] for n, x in zip(names, b)})
