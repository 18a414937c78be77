"""OED criteria -- mirror of lib/CorleoneOED/src/criteria.jl:33-63.

Each criterion is callable on one symmetric Fisher matrix (host, NumPy: the reference's `crit(F)`), and
`batched_criteria` evaluates all six for a whole batch of candidates on the GPU together with the Fisher
information itself (cb200_eval with CB200_WANT_CRIT: cyclic Jacobi per candidate inside the FIM read-out kernel).
"""
from __future__ import annotations

import numpy as np

from . import native

ORDER = ("ACriterion", "DCriterion", "ECriterion", "FisherACriterion", "FisherDCriterion", "FisherECriterion")


def _sym(F):
    F = np.asarray(F, dtype=np.float64)
    return np.triu(F) + np.triu(F, 1).T          # Symmetric(F) reads the upper triangle


class ACriterion:
    """tr(inv(F)) (:41-43)"""
    def __call__(self, F): return float(np.trace(np.linalg.inv(_sym(F))))


class DCriterion:
    """inv(det(F)) (:45-47)"""
    def __call__(self, F): return float(1.0 / np.linalg.det(_sym(F)))


class ECriterion:
    """maximum(eigvals(inv(F))) (:49-51)"""
    def __call__(self, F): return float(np.max(np.linalg.eigvalsh(np.linalg.inv(_sym(F)))))


class FisherACriterion:
    """-tr(F) (:53-55)"""
    def __call__(self, F): return float(-np.trace(_sym(F)))


class FisherDCriterion:
    """-det(F) (:57-59)"""
    def __call__(self, F): return float(-np.linalg.det(_sym(F)))


class FisherECriterion:
    """-minimum(eigvals(F)) (:61-63)"""
    def __call__(self, F): return float(-np.min(np.linalg.eigvalsh(_sym(F))))


def batched_criteria(nat: "native.NativeLayer", ps_batch, fim_init=None):
    """(B, 6) array, columns in `ORDER`, of F = F_init + F(t_end) for every candidate; also returns the FIMs."""
    r = nat.eval_host(ps_batch, native.WANT_FIM | native.WANT_CRIT, fim_init=fim_init)
    return r["criteria"], r["fim"].reshape(-1, nat.fim_n, nat.fim_n)


def criteria_gradient(nat: "native.NativeLayer", ps_batch, fim_init=None):
    """Gradients of the six criteria w.r.t. the flat variable vector, per candidate: (B, 6, nvars).

    One cb200_eval delivers F and the second-order sensitivities d F(t_end) / d(variables of the last interval)
    (rows of `sens` that belong to the F_triu states); the chain rule through the small n x n matrix functions is
    host arithmetic.  With multiple shooting F is a shooting state, so only the last interval's block is non-zero
    (the other blocks enter through the matching constraints, as in the reference's NLP)."""
    r = nat.eval_host(ps_batch, native.WANT_FIM | native.WANT_SENS, fim_init=fim_init)
    B, n, nx, I = r["fim"].shape[0], nat.fim_n, nat.nx, nat.I
    blocks = nat.block_structure()
    ns_last = int(blocks[I] - blocks[I - 1])
    sens_off = sum(nx * int(blocks[i + 1] - blocks[i]) for i in range(I - 1))
    f_off = nat.fim_offset - 1
    out = np.zeros((B, 6, nat.nvars))
    iu = [(a, b) for b in range(n) for a in range(b + 1)]          # column-major upper triangle (augmentation.jl:211)
    for c in range(B):
        F = r["fim"][c].reshape(n, n)
        S = r["sens"][c][sens_off:sens_off + nx * ns_last].reshape(ns_last, nx)    # [direction, state]
        dF = np.zeros((n, n, ns_last))
        for q, (a, b) in enumerate(iu):
            dF[a, b] = dF[b, a] = S[:, f_off + q]
        out[c, :, blocks[I - 1]:blocks[I]] = criteria_chain(F, dF)
    return out


def criteria_chain(F, dF):
    """d(criteria)/dv from F (n, n) and dF/dv (n, n, nv): (6, nv), rows in `ORDER` (criteria.jl:33-63 differentiated)."""
    lam, V = np.linalg.eigh(F)
    Finv = np.linalg.inv(F)
    det = np.linalg.det(F)
    vmin = V[:, 0]
    t = np.einsum("ab,bav->v", Finv, dF)                 # tr(F^-1 dF)
    q = np.einsum("a,abv,b->v", vmin, dF, vmin)          # v_min' dF v_min
    return np.stack([-np.einsum("ab,bcv,ca->v", Finv, dF, Finv), -t / det, -q / lam[0] ** 2,
                     -np.einsum("aav->v", dF), -det * t, -q])


def _hxG(traj_samples, bx: int, nf: int, nobs: int):
    """h_x G at every trajectory sample for observed = the leading nobs states (unit rows of h_x):
    traj_samples (..., n_samples, n_row) -> (..., n_samples, nobs, nf)."""
    G = np.asarray(traj_samples)[..., bx:bx + bx * nf]
    G = G.reshape(G.shape[:-1] + (nf, bx))            # vec(G) is column-major: [column c][state r]
    return np.swapaxes(G, -1, -2)[..., :nobs, :]


def local_information_gain(traj_samples, bx: int, nf: int, nobs: int):
    """local_information_gain (lib/CorleoneOED/src/oed.jl:500-510): for every sample t_k and observed function i
    the rank-one matrix (h_i G(t_k))'(h_i G(t_k)); result (..., n_samples, nobs, nf, nf).  `traj_samples` are the
    merged trajectory samples of the augmented layer as cb200_eval(WANT_TRAJ) returns them (batched or not)."""
    x = _hxG(traj_samples, bx, nf, nobs)
    return x[..., :, None] * x[..., None, :]


def global_information_gain(traj_samples, F_tf, bx: int, nf: int, nobs: int):
    """global_information_gain (oed.jl:520-534): the same with rows scaled by C = inv(F(t_f))."""
    C = np.linalg.inv(np.asarray(F_tf))
    x = _hxG(traj_samples, bx, nf, nobs)
    x = np.einsum("...koc,...cd->...kod", x, C) if C.ndim > 2 else x @ C
    return x[..., :, None] * x[..., None, :]


def sampling_sum_matrix(layer, st, ps_like, rows):
    """Linear map of the sampling sums (lib/CorleoneOED/src/oed.jl:610-618): for every measurement control (row of the
    index grid, 0-based in `rows`) the integral of its piecewise-constant weight, sum_j controls[index_grid[row, j]]
    * dt_j, as a matrix acting on the flat variable vector; a multiple-shooting layer sums over the intervals (:566-570)."""
    from .multiple_shooting import get_block_structure
    from .single_shooting import SingleShootingLayer, _flatten_chunks
    single = isinstance(layer, SingleShootingLayer)
    sts = [st] if single else list(st.values())
    pss = [ps_like] if single else list(ps_like.values())
    blocks = get_block_structure(layer)
    A = np.zeros((len(rows), int(blocks[-1])))
    for i, (s, p) in enumerate(zip(sts, pss)):
        grid = np.asarray(_flatten_chunks(s["index_grid"]), dtype=np.int64)
        tsp = _flatten_chunks(s["tspans"])
        dts = np.array([b - a for a, b in tsp])
        c0 = int(blocks[i]) + len(p["u0"]) + len(p["p"])
        for k, row in enumerate(rows):
            np.add.at(A[k], c0 + grid[row] - 1, dts)
    return A


def solve_oed_slsqp(layer, st, ps_like, nat, criterion: int = 0, M=None, measurement_rows=(), fim_init=None, x0=None,
                    maxiter=300, ftol=1e-12):
    """Minimise one criterion (index into ORDER) of F = F_init + F(t_end) over the flat variables within their bounds,
    subject to sampling sums <= M (lib/CorleoneOED/ext/CorleoneOEDOptimizationExtension.jl), with SciPy's SLSQP standing
    in for Ipopt.  Value, FIM and second-order sensitivities come from one cb200_eval per iterate; single shooting."""
    from scipy.optimize import Bounds, LinearConstraint, minimize
    from .multiple_shooting import get_bounds
    from .single_shooting import to_vec
    x0 = to_vec(ps_like) if x0 is None else np.asarray(x0, dtype=np.float64)
    lo, hi = get_bounds(layer)
    cache = {}

    def ev(x):
        key = x.tobytes()
        if cache.get("key") != key:
            crit, F = batched_criteria(nat, x[None, :], fim_init=fim_init)
            g = criteria_gradient(nat, x[None, :], fim_init=fim_init)
            cache.update(key=key, val=(float(crit[0, criterion]), g[0, criterion].copy(), F[0]))
        return cache["val"]

    cons = []
    if M is not None and len(measurement_rows):
        cons.append(LinearConstraint(sampling_sum_matrix(layer, st, ps_like, list(measurement_rows)), -np.inf,
                                     np.asarray(M, dtype=np.float64)))
    res = minimize(lambda x: ev(x)[0], x0, jac=lambda x: ev(x)[1], bounds=Bounds(to_vec(lo), to_vec(hi)),
                   constraints=cons, method="SLSQP", options=dict(maxiter=maxiter, ftol=ftol))
    res.fim = ev(res.x)[2]
    return res
