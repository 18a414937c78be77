"""CorleoneDynamicOptProblem -- host-side mirror of src/dynprob.jl.

The reference builds objective / constraint closures that each re-integrate the trajectory (dynprob.jl:68-73,
109-116, 123-129) and lets Optimization.jl differentiate them with ForwardDiff (:143-164).  Here ONE cb200_eval per
iterate delivers f, grad f, the shooting constraints, the nonzeros of their Jacobian and -- for point constraints --
the trajectory samples with their sensitivities; this module only arranges them in the reference's order:

    constraints = [ point constraints (in the order given) ; shooting_constraints (kind-major) ]   (:105-116)

Point constraints are `(symbol, dict(t=[...], bounds=(lo, hi)))`: the symbol's trajectory row at the samples whose
time is in `t` (:89-103).  Deviation, on purpose: the reference takes `findall(in(t), tpoints)` on the concatenation
of every interval's piece boundaries (:75-84), which for a multiple-shooting layer still contains the duplicated
interval boundaries, and applies these positions to the MERGED trajectory (no duplicates) -- off by the number of
earlier interval boundaries and out of range near the end.  Here the positions are taken in the merged time grid,
which is what the reference computes for single shooting and evidently means for multiple shooting.
"""
from __future__ import annotations

import numpy as np

from . import native
from .multiple_shooting import (MultipleShootingLayer, get_bounds, get_number_of_shooting_constraints,
                                trajectory_jacobian)
from .single_shooting import SingleShootingLayer, _flatten_chunks, to_vec


class CorleoneDynamicOptProblem:
    def __init__(self, layer, loss: str, *constraints, rng=None):
        self.layer = layer
        self.ps0 = layer.initialparameters(rng)
        self.st = layer.initialstates(rng)
        single = isinstance(layer, SingleShootingLayer)
        sts = [self.st] if single else list(self.st.values())
        self.sys = sts[0]["symcache"]
        if loss not in self.sys.variables:
            raise KeyError(f"loss {loss!r} is not a trajectory variable: {sorted(self.sys.variables)}")
        self.loss_row = self.sys.variables[loss] + 1            # 1-based row of a trajectory sample
        self.nat = layer._native(self.st, self.ps0)
        # traj.t of the merged trajectory: piece starts of every interval (and saveat times inside pieces) + the final end
        self.tpoints = self.nat.sample_times()
        self.point = []                                          # (row 0-based, sample indices 0-based)
        lb, ub = [], []
        for expr, specs in constraints:
            if not isinstance(expr, str):
                raise TypeError(f"The constraint {expr} is neither a symbol nor an expression!")
            tidx = np.nonzero(np.isin(self.tpoints, np.asarray(specs["t"], dtype=np.float64)))[0]
            if len(tidx):
                self.point.append((self.sys.variables[expr], tidx))
                lo, hi = min(specs["bounds"]), max(specs["bounds"])
                lb.append(np.full(len(tidx), lo))
                ub.append(np.full(len(tidx), hi))
        self.n_point = int(sum(len(t) for _, t in self.point))
        self.n_shoot = 0 if single else get_number_of_shooting_constraints(layer, self.ps0, self.st)
        if self.point or not single:
            lb.append(np.zeros(self.n_shoot))
            ub.append(np.zeros(self.n_shoot))
            self.lcons, self.ucons = np.concatenate(lb), np.concatenate(ub)
        else:
            self.lcons = self.ucons = None
        self.u0 = to_vec(self.ps0)
        blo, bhi = get_bounds(layer)
        self.lb, self.ub = to_vec(blo), to_vec(bhi)
        self._rows, self._cols, self._src = self.nat.jac_structure()
        self._const = np.where(self._src[self.nat.jac_nnz_var:] == -1, 1.0, -1.0)

    # ---- one device evaluation per iterate
    def _eval(self, X, derivatives: bool):
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        want = native.WANT_OBJ | native.WANT_DEFECTS
        if self.point:
            want |= native.WANT_TRAJ
        if derivatives:
            want |= native.WANT_GRAD | native.WANT_JAC
            if self.point:
                want |= native.WANT_TRAJ_SENS | native.WANT_SENS
        r = self.nat.eval_host(X, want, objective_row=self.loss_row)
        native.check_status(r, "CorleoneDynamicOptProblem")   # a failed integration must not reach the optimiser as NaN
        return r

    def _cons_from(self, r, b):
        parts = []
        if self.point:
            traj = r["traj"][b].reshape(self.nat.n_samples, self.nat.n_row)
            parts += [traj[tidx, row] for row, tidx in self.point]
        parts.append(r["defects_kind"][b])
        return np.concatenate(parts)

    def _jac_from(self, r, b):
        nvars = self.nat.nvars
        J = np.zeros((self.n_point + self.n_shoot, nvars))
        if self.point:
            TJ = trajectory_jacobian(self.layer, self.st, self.ps0, r, b)     # (n_row, n_samples, nvars)
            o = 0
            for row, tidx in self.point:
                J[o:o + len(tidx)] = TJ[row, tidx, :]
                o += len(tidx)
        vals = np.concatenate([r["jac_vals"][b], self._const])
        np.add.at(J, (self.n_point + self._rows - 1, self._cols - 1), vals)
        return J

    # ---- the reference's callbacks (one candidate)
    def objective(self, x):
        """optprob.f(u, p): last(getsym(sys, loss)(traj)) (:68-73)"""
        return float(self._eval(x, False)["objective"][0, 0])

    def constraints(self, x):
        """optprob.f.cons(res, u, p) (:105-131)"""
        return self._cons_from(self._eval(x, False), 0)

    def gradient(self, x):
        return self._eval(x, True)["grad"][0].copy()

    def constraint_jacobian(self, x):
        return self._jac_from(self._eval(x, True), 0)

    def value_and_derivatives(self, x):
        """f, grad f, c, dc/dx from ONE evaluation (what an NLP iterate needs)"""
        r = self._eval(x, True)
        return float(r["objective"][0, 0]), r["grad"][0].copy(), self._cons_from(r, 0), self._jac_from(r, 0)

    # ---- the batch axis the reference does not have
    def evaluate_batch(self, X):
        """X: (B, nvars).  Returns objective (B,), constraints (B, n_con), gradient (B, nvars)."""
        r = self._eval(X, True)
        B = r["objective"].shape[0]
        return r["objective"][:, 0].copy(), np.stack([self._cons_from(r, b) for b in range(B)]), r["grad"].copy()


def solve_slsqp(prob: CorleoneDynamicOptProblem, x0=None, maxiter=500, ftol=1e-10):
    """Solve the NLP with SciPy's SLSQP fed by the device-computed derivatives (the reference's tests use Ipopt through
    Optimization.jl; this is the stand-in that lets the end-to-end goldens be checked from Python)."""
    from scipy.optimize import Bounds, minimize
    cache = {}

    def ev(x):
        key = x.tobytes()
        if cache.get("key") != key:
            cache["key"] = key
            cache["val"] = prob.value_and_derivatives(x)
        return cache["val"]

    cons = []
    if prob.lcons is not None and len(prob.lcons):
        eq = prob.lcons == prob.ucons
        if eq.any():
            cons.append(dict(type="eq", fun=lambda x: (ev(x)[2] - prob.lcons)[eq], jac=lambda x: ev(x)[3][eq]))
        if (~eq).any():
            lo, hi = np.isfinite(prob.lcons) & ~eq, np.isfinite(prob.ucons) & ~eq
            if lo.any():
                cons.append(dict(type="ineq", fun=lambda x: (ev(x)[2] - prob.lcons)[lo], jac=lambda x: ev(x)[3][lo]))
            if hi.any():
                cons.append(dict(type="ineq", fun=lambda x: (prob.ucons - ev(x)[2])[hi], jac=lambda x: -ev(x)[3][hi]))
    x0 = prob.u0 if x0 is None else np.asarray(x0, dtype=np.float64)
    return minimize(lambda x: ev(x)[0], x0, jac=lambda x: ev(x)[1], bounds=Bounds(prob.lb, prob.ub), constraints=cons,
                    method="SLSQP", options=dict(maxiter=maxiter, ftol=ftol))
