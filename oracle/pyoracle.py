"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODELS = {
    "lotka3": 0,
    "lqr": 1,
    "lotka2": 2,
    "lotka2_oed": 3,
    "lin1d": 4,
    "lin1d_oed": 5,
    "dense20": 6,
    "egerstedt": 7,
    "vdp3": 8,
    "cartpole5": 9,
    "rocket3": 10,
    "quadrotor7": 11,
    "threetank4": 12,
    "batchreactor2": 13,
    "lotkacomp3": 14,
    "f8aircraft4": 15,
    "doubleosc6": 16,
    "lotka2_oed_cu": 17,
    "lotka2_oed_cf": 18,
    "lotka2_oed_ds": 19,
    "lotka2_oed_du": 20,
    "lin1d_oed_cf": 21,
    "lin1d_oed_ds": 22,
    "brysondenham3": 23,
    "catalystmixing2": 24,
    "cushioned3": 25,
    "denbigh3": 26,
    "dielectrophoretic3": 27,
    "electriccar4": 28,
    "fuller3": 29,
    "jackson3": 30,
    "mountaincar3": 31,
    "particlesteering5": 32,
    "raomease2": 33,
    "robbins4": 34,
    "moonlanding3": 35,
    "lotkashared4": 36,
    "hangingchain3": 37,
    "tubularreactor2": 38,
    "ocean4": 39,
    "ductedfan7": 40,
    "hangglider4": 41,
    "dense16": 42,
    "dense24": 43,
    "dense32": 44,
}
METHODS = {"tsit5": 0, "rk4": 1, "tsit5_adaptive": 2}
CO_ND = 16


class _Problem(C.Structure):
    _fields_ = [
        ("model", C.c_int),
        ("nx", C.c_int),
        ("np_all", C.c_int),
        ("u0", C.POINTER(C.c_double)),
        ("p0", C.POINTER(C.c_double)),
        ("tspan", C.c_double * 2),
        ("ncontrols", C.c_int),
        ("control_indices", C.POINTER(C.c_int)),
        ("ctrl_t", C.POINTER(C.POINTER(C.c_double))),
        ("ctrl_u", C.POINTER(C.POINTER(C.c_double))),
        ("ctrl_nt", C.POINTER(C.c_int)),
        ("n_tunable_ic", C.c_int),
        ("tunable_ic", C.POINTER(C.c_int)),
        ("nquad", C.c_int),
        ("quad_idx", C.POINTER(C.c_int)),
        ("n_shoot", C.c_int),
        ("shoot_pts", C.POINTER(C.c_double)),
        ("model_data", C.POINTER(C.c_double)),
        ("n_saveat", C.c_int),
        ("saveat", C.POINTER(C.c_double)),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "corleone_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp, ip, i64p = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int64)
        L.co_layer_create.restype = C.c_void_p
        L.co_layer_create.argtypes = [C.POINTER(_Problem)]
        L.co_layer_destroy.argtypes = [C.c_void_p]
        for f in ("co_n_intervals", "co_nvars", "co_nrow", "co_n_defects", "co_n_samples"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.co_block_structure.argtypes = [C.c_void_p, ip]
        L.co_interval_info.argtypes = [C.c_void_p, C.c_int, dp, ip, ip, ip, ip, ip]
        L.co_interval_tables.argtypes = [C.c_void_p, C.c_int, i64p, dp, ip, ip, ip, ip]
        L.co_default_ps.argtypes = [C.c_void_p, dp]
        L.co_eval.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int] + [dp] * 10
        L.co_eval.restype = C.c_int
        L.co_eval_traj.argtypes = [C.c_void_p, dp, dp, C.c_int, C.c_int, C.c_int, dp, dp]
        L.co_eval_traj.restype = C.c_int
        L.co_forward_init.argtypes = [C.c_void_p, dp, ip, C.c_int, C.c_int, C.c_int]
        L.co_forward_init.restype = C.c_int
        L.co_bench_units.argtypes = [C.c_void_p, dp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                     C.c_int, C.c_int, dp]
        L.co_bench_units.restype = C.c_double
        L.co_max_threads.restype = C.c_int
        L.co_build_index_grid.argtypes = [C.c_int, C.POINTER(dp), ip, C.c_double, C.c_double, i64p, dp, C.c_int]
        L.co_build_index_grid.restype = C.c_int
        L.co_restrict_grid.argtypes = [dp, C.c_int, C.c_double, C.c_double, ip]
        L.co_restrict_grid.restype = C.c_int
        L.co_subvector_indices.argtypes = [C.c_int, C.c_int, ip]
        L.co_subvector_indices.restype = C.c_int
        _LIB = L
    return _LIB


def set_tolerances(abstol: float, reltol: float):
    """abstol / reltol of method "tsit5_adaptive" (the reference passes them through the ODEProblem's kwargs)"""
    lib().co_set_tolerances(C.c_double(abstol), C.c_double(reltol))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


def build_index_grid(ts, t0, t1):
    """src/local_controls.jl:149-195 -> (idx [nc x M] int64 1-based, grid [M+1])."""
    ts = [np.ascontiguousarray(t, dtype=np.float64) for t in ts]
    nc = len(ts)
    cap = 2 + sum(len(t) for t in ts)
    arr = (C.POINTER(C.c_double) * max(nc, 1))(*[_dp(t) for t in ts])
    nt = np.array([len(t) for t in ts] or [0], dtype=np.int32)
    idx = np.zeros(max(nc, 1) * cap, dtype=np.int64)
    grid = np.zeros(cap + 1)
    M = lib().co_build_index_grid(nc, arr, _ip(nt), float(t0), float(t1),
                                  idx.ctypes.data_as(C.POINTER(C.c_int64)), _dp(grid), cap)
    assert M >= 0
    return idx[: nc * M].reshape((M, nc)).T.copy(), grid[: M + 1].copy()


def restrict_grid(t, t0, t1):
    t = np.ascontiguousarray(t, dtype=np.float64)
    pos = np.zeros(len(t), dtype=np.int32)
    n = lib().co_restrict_grid(_dp(t), len(t), float(t0), float(t1), _ip(pos))
    return pos[:n].copy()


def subvector_indices(M, L):
    ss = np.zeros(2 * (M // L + 2), dtype=np.int32)
    n = lib().co_subvector_indices(M, L, _ip(ss))
    return [(int(ss[2 * k]), int(ss[2 * k + 1])) for k in range(n)]


class OracleLayer:
    """A SingleShootingLayer (no shooting_points) or MultipleShootingLayer restated on the CPU.

    spec keys: model, u0, p0, tspan, controls=[(p_index_1based, t, u_or_None), ...], tunable_ic=[...1-based],
    quadrature_indices=[...1-based], shooting_points=[...] or None, model_data (dense20: W.ravel() ++ b),
    saveat=[sorted times] or None (problem.kwargs saveat: samples inside pieces from Tsit5's dense output).
    """

    def __init__(self, spec: dict):
        self._keep = []
        P = _Problem()
        P.model = MODELS[spec["model"]]
        u0 = np.ascontiguousarray(spec["u0"], dtype=np.float64)
        p0 = np.ascontiguousarray(spec["p0"], dtype=np.float64)
        P.nx, P.np_all = len(u0), len(p0)
        P.u0, P.p0 = _dp(u0), _dp(p0)
        P.tspan[0], P.tspan[1] = spec["tspan"]
        ctr = spec.get("controls", [])
        nc = len(ctr)
        cidx = np.array([c[0] for c in ctr] or [0], dtype=np.int32)
        ts = [np.ascontiguousarray(c[1], dtype=np.float64) for c in ctr]
        us = [None if c[2] is None else np.ascontiguousarray(c[2], dtype=np.float64) for c in ctr]
        nt = np.array([len(t) for t in ts] or [0], dtype=np.int32)
        tarr = (C.POINTER(C.c_double) * max(nc, 1))(*[_dp(t) for t in ts])
        uarr = (C.POINTER(C.c_double) * max(nc, 1))(*[_dp(u) if u is not None else None for u in us])
        P.ncontrols = nc
        P.control_indices = _ip(cidx)
        P.ctrl_t, P.ctrl_u, P.ctrl_nt = tarr, uarr, _ip(nt)
        tic = np.array(list(spec.get("tunable_ic", [])) or [0], dtype=np.int32)
        P.n_tunable_ic = len(spec.get("tunable_ic", []))
        P.tunable_ic = _ip(tic)
        quad = np.array(list(spec.get("quadrature_indices", [])) or [0], dtype=np.int32)
        P.nquad = len(spec.get("quadrature_indices", []))
        P.quad_idx = _ip(quad)
        sp = spec.get("shooting_points")
        spa = np.ascontiguousarray(sp if sp is not None else [0.0], dtype=np.float64)
        P.n_shoot = 0 if sp is None else len(sp)
        P.shoot_pts = _dp(spa)
        md = spec.get("model_data")
        mda = None if md is None else np.ascontiguousarray(md, dtype=np.float64)
        P.model_data = _dp(mda) if mda is not None else None
        sv = spec.get("saveat")
        sva = np.ascontiguousarray(sv if sv is not None and len(sv) else [0.0], dtype=np.float64)
        P.n_saveat = 0 if sv is None else len(sv)
        P.saveat = _dp(sva)
        self._keep += [u0, p0, cidx, ts, us, nt, tarr, uarr, tic, quad, spa, mda, sva, P]
        self.h = lib().co_layer_create(C.byref(P))
        L = lib()
        self.nx = P.nx
        self.I = L.co_n_intervals(self.h)
        self.nvars = L.co_nvars(self.h)
        self.nrow = L.co_nrow(self.h)
        self.ndef = L.co_n_defects(self.h)
        self.nsamples = L.co_n_samples(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().co_layer_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def block_structure(self):
        b = np.zeros(self.I + 1, dtype=np.int32)
        lib().co_block_structure(self.h, _ip(b))
        return b.astype(np.int64)

    def interval(self, i):
        ts = np.zeros(2)
        M, nu0, npp, ncc, nsi = (C.c_int() for _ in range(5))
        lib().co_interval_info(self.h, i, _dp(ts), C.byref(M), C.byref(nu0), C.byref(npp), C.byref(ncc),
                               C.byref(nsi))
        nact = self.nrow - self.nx
        idx = np.zeros(max(nact * M.value, 1), dtype=np.int64)
        grid = np.zeros(M.value + 1)
        sidx = np.zeros(max(nsi.value, 1), dtype=np.int32)
        sorting = np.zeros(self.nx, dtype=np.int32)
        consts = np.zeros(self.nx, dtype=np.int32)
        ncon = C.c_int()
        lib().co_interval_tables(self.h, i, idx.ctypes.data_as(C.POINTER(C.c_int64)), _dp(grid), _ip(sidx),
                                 _ip(sorting), _ip(consts), C.byref(ncon))
        return dict(tspan=(ts[0], ts[1]), M=M.value, n_u0=nu0.value, n_p=npp.value, n_c=ncc.value,
                    index_grid=idx[: nact * M.value].reshape((M.value, nact)).T.copy(), grid=grid,
                    shooting_indices=sidx[: nsi.value].astype(np.int64), sorting=sorting.astype(np.int64),
                    constants=consts[: ncon.value].astype(np.int64))

    def default_ps(self):
        ps = np.zeros(self.nvars)
        lib().co_default_ps(self.h, _dp(ps))
        return ps

    def forward_init(self, ps, fixed_indices=(), method="tsit5", K=1):
        ps = np.ascontiguousarray(ps, dtype=np.float64).copy()
        fx = np.array(list(fixed_indices) or [0], dtype=np.int32)
        rc = lib().co_forward_init(self.h, _dp(ps), _ip(fx), len(fixed_indices), METHODS[method], K)
        assert rc == 0, rc
        return ps

    def eval(self, ps, dps=None, method="tsit5", K=1):
        ps = np.ascontiguousarray(ps, dtype=np.float64)
        assert ps.shape == (self.nvars,)
        nd = 0 if dps is None else dps.shape[1]
        assert nd <= CO_ND
        dpsf = None if dps is None else np.asfortranarray(dps, dtype=np.float64)
        nx, I, nrow, ns, ndef = self.nx, self.I, self.nrow, self.nsamples, self.ndef
        out = dict(t=np.zeros(ns), u=np.zeros((nrow, ns), order="F"), xend=np.zeros((nx, I), order="F"),
                   defk=np.zeros(max(ndef, 1)), defs=np.zeros(max(ndef, 1)), last=np.zeros(nrow))
        d = dict(dxend=None, ddefk=None, ddefs=None, dlast=None)
        if nd:
            d = dict(dxend=np.zeros((nx * I, nd), order="F"), ddefk=np.zeros((max(ndef, 1), nd), order="F"),
                     ddefs=np.zeros((max(ndef, 1), nd), order="F"), dlast=np.zeros((nrow, nd), order="F"))
        rc = lib().co_eval(self.h, _dp(ps), _dp(dpsf) if nd else None, nd, METHODS[method], K, _dp(out["t"]),
                           _dp(out["u"]), _dp(out["xend"]), _dp(d["dxend"]), _dp(out["defk"]), _dp(d["ddefk"]),
                           _dp(out["defs"]), _dp(d["ddefs"]), _dp(out["last"]), _dp(d["dlast"]))
        assert rc == 0
        out["defk"], out["defs"] = out["defk"][:ndef], out["defs"][:ndef]
        if nd:
            d["ddefk"], d["ddefs"] = d["ddefk"][:ndef], d["ddefs"][:ndef]
            out.update(d)
        return out

    def jacobians(self, ps, method="tsit5", K=1):
        """Dense d(xend)/d(ps), d(defk)/d(ps), d(last)/d(ps) by chunks of CO_ND identity seeds."""
        n = self.nvars
        Jx = np.zeros((self.nx * self.I, n))
        Jd = np.zeros((self.ndef, n))
        Jl = np.zeros((self.nrow, n))
        for c0 in range(0, n, CO_ND):
            nd = min(CO_ND, n - c0)
            seeds = np.zeros((n, nd), order="F")
            seeds[c0 + np.arange(nd), np.arange(nd)] = 1.0
            r = self.eval(ps, seeds, method, K)
            Jx[:, c0:c0 + nd] = r["dxend"]
            Jd[:, c0:c0 + nd] = r["ddefk"]
            Jl[:, c0:c0 + nd] = r["dlast"]
        return Jx, Jd, Jl

    def traj_jacobian(self, ps, method="tsit5", K=1):
        """Dense d(merged trajectory samples)/d(ps): (nrow, nsamples, nvars), quadrature rows accumulated as in
        multiple_shooting.jl:171-177 (what the point constraints of dynprob.jl differentiate)."""
        ps = np.ascontiguousarray(ps, dtype=np.float64)
        n = self.nvars
        J = np.zeros((self.nrow, self.nsamples, n))
        u = np.zeros((self.nrow, self.nsamples), order="F")
        for c0 in range(0, n, CO_ND):
            nd = min(CO_ND, n - c0)
            seeds = np.zeros((n, nd), order="F")
            seeds[c0 + np.arange(nd), np.arange(nd)] = 1.0
            du = np.zeros((self.nrow, self.nsamples, nd), order="F")
            rc = lib().co_eval_traj(self.h, _dp(ps), _dp(seeds), nd, METHODS[method], K, _dp(u), _dp(du))
            assert rc == 0
            J[:, :, c0:c0 + nd] = du
        return u, J

    def bench_units(self, ps_batch, method="tsit5", K=1, with_sens=True, nthreads=0, b_range=None):
        """ps_batch: (B, nvars) C-contiguous == nvars x B column-major. Returns (seconds, units, checksum)."""
        ps_batch = np.ascontiguousarray(ps_batch, dtype=np.float64)
        B = ps_batch.shape[0]
        b0, b1 = b_range if b_range is not None else (0, B)
        chk = C.c_double()
        s = lib().co_bench_units(self.h, _dp(ps_batch), B, b0, b1, METHODS[method], K, int(with_sens),
                                 nthreads, C.byref(chk))
        return s, (b1 - b0) * self.I, chk.value


def max_threads():
    return lib().co_max_threads()


# ------------------------------------------------------------------------------------------------------------------
# Fisher read-outs of lib/CorleoneOED/src/oed.jl, restated literally (plain loops, 1-based tables as Julia holds them).
# `u` is the (nrow, nsamples) sample matrix of OracleLayer.eval; observed = the leading n_obs states, so
# h_x[k, :] G = row k of the bx x nf block G that follows the base states (augmentation.jl:131-159).
def _hxG(u, s, bx, nf, n_obs):
    """[h_1 G; h_2 G; ...] at sample s: (n_obs, nf)"""
    return np.array([[u[bx + k + bx * a, s] for a in range(nf)] for k in range(n_obs)])


def weighting_grid(control_grids, sampling_positions, index_grid, tspan, restrict):
    """WeightedObservation.grid (oed.jl:291-311 for a single-shooting layer, :341-357 per interval of a multiple-shooting
    layer with restrict = True).  control_grids: the full time grid of EVERY control of the layer; sampling_positions:
    1-based positions of the measurement controls among them; index_grid: build_index_grid(controls...; tspan) as an
    (n_controls, M) 1-based matrix.  Returns a list (one entry per point of the overall grid) of lists (one entry per
    measurement control) of 1-based indices into ps.controls, 0 = no measurement there."""
    overall = sorted(set(float(t) for g in control_grids for t in g) | {float(tspan[0]), float(tspan[1])})
    if restrict:
        overall = [t for t in overall if tspan[0] <= t < tspan[1]]
    observed_grid = []
    for s in sampling_positions:
        grid = sorted(set(float(t) for t in control_grids[s - 1]))
        observed_grid.append([i + 1 for i, t in enumerate(overall) if t in grid])          # findall(in(grid), overall)
    measurement_indices = []
    for s in sampling_positions:
        seen = []
        for v in index_grid[s - 1]:
            if int(v) not in seen:
                seen.append(int(v))                                                        # unique(mi)
        measurement_indices.append(seen)
    out = []
    for i in range(1, len(overall) + 1):
        row = []
        for j in range(len(observed_grid)):
            idn = observed_grid[j].index(i) + 1 if i in observed_grid[j] else None         # findfirst(i .== observed_grid[j])
            row.append(0 if idn is None else measurement_indices[j][idn - 1])
        out.append(row)
    return out


def fisher_discrete_sampled(controls, grid, u, sample0, bx, nf, n_obs):
    """WeightedObservation call (oed.jl:2-16): sum_i (psub .* G_i)'(psub .* G_i) over the rows of `grid`; row i
    belongs to sample sample0 + i of `u`.  controls: the (interval's) ps.controls vector."""
    F = np.zeros((nf, nf))
    for i, row in enumerate(grid):
        psub = np.array([0.0 if j == 0 else controls[j - 1] for j in row])
        G = psub[:, None] * _hxG(u, sample0 + i, bx, nf, n_obs)
        F += G.T @ G
    return F


def fisher_discrete(u, bx, nf, n_obs):
    """OEDLayer{true, false} (oed.jl:446-461) with the output expression of augmentation.jl:182-186: the rows of
    h_x G are summed first, then G'G, summed over ALL saved samples."""
    F = np.zeros((nf, nf))
    for s in range(u.shape[1]):
        g = _hxG(u, s, bx, nf, n_obs).sum(axis=0)[None, :]
        F += g.T @ g
    return F


def fisher_fixed_continuous(controls, active_controls, u, bx, nf, n_obs):
    """OEDLayer{false, true, true} (oed.jl:352-362): fim[s] is the (nf, nf, n_obs) array of the per-observation
    Fisher states at sample s (augmentation.jl:262-271), w = eachrow(hcat(controls[active]...)), and
    F = sum over zip(w, diff(fim)) of F[:, :, k] * w[k]."""
    f_off = bx * (1 + nf)
    ntri = nf * (nf + 1) // 2

    def fim_at(s):
        out = np.zeros((nf, nf, n_obs))
        for k in range(n_obs):
            for i in range(nf):
                for j in range(nf):
                    lo, hi = min(i, j), max(i, j)
                    out[i, j, k] = u[f_off + k * ntri + hi * (hi + 1) // 2 + lo, s]
        return out

    fim = [fim_at(s) for s in range(u.shape[1])]
    w = np.stack([np.asarray(controls)[np.asarray(a) - 1] for a in active_controls], axis=1)   # rows = eachrow(hcat)
    diffF = [b - a for a, b in zip(fim[:-1], fim[1:])]
    F = np.zeros((nf, nf))
    for wi, dF in zip(w, diffF):
        for k in range(n_obs):
            F += dF[:, :, k] * wi[k]
    return F
