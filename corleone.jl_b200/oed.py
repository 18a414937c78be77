"""OEDLayer -- host-side mirror of lib/CorleoneOED/src/oed.jl (:30-180 constructors, :286-372 initial states,
:381-470 Fisher read-outs) for the compiled-in base models.

The reference derives the augmented right-hand side symbolically (augmentation.jl); here each (base model, variant)
pair is a compiled-in model of libcorleone_b200.so (csrc/models.cuh `OedAug`):

    OEDLayer{false}, measurements          -> "<base>_oed"      [x; G; F_triu],            p = [base p; w]
    OEDLayer{false}, no measurements       -> "<base>_oed_cu"   [x; G; F_triu],            p = base p
    OEDLayer{false}, measurements, FIXED   -> "<base>_oed_cf"   [x; G; F_1 triu; F_2 ...], p = [base p; w]
    OEDLayer{true},  measurements          -> "<base>_oed_ds"   [x; G],                    p = base p       (*)
    OEDLayer{true},  no measurements       -> "<base>_oed_du"   [x; G],                    p = base p

(*) As in the reference, the discrete variants do not append the weights to the problem's parameters
(augmentation.jl:171-203): the measurement controls (p index beyond length(p)) are filtered out of the index grid
(src/single_shooting.jl:301-304), their values stay in ps.controls, and the samples at the measurement times come from
`saveat` (oed.jl:106-113) -- Tsit5's dense output on the step that contains them (cb200_desc.saveat), not from cutting
the pieces.

Every evaluation goes through cb200_eval (`CB200_WANT_FIM` / `CB200_WANT_CRIT`); this module only builds tables.
"""
from __future__ import annotations

import numpy as np

from . import native
from .controls import ControlParameter, build_index_grid, get_timegrid
from .multiple_shooting import MultipleShootingLayer
from .problem import MODEL_REGISTRY
from .single_shooting import SingleShootingLayer, _flatten_chunks, to_vec

# base model -> (Fisher parameter indices (1-based), number of observed functions = leading states)
OED_BASE = {"lotka2": ((2, 3), 2), "lin1d": ((1,), 1)}
FIM_END, FIM_DISCRETE, FIM_DISCRETE_SAMPLED, FIM_FIXED_CONTINUOUS = 0, 1, 2, 3


def _variant(discrete: bool, sampled: bool, fixed: bool) -> str:
    if discrete:
        return "_oed_ds" if sampled else "_oed_du"
    if not sampled:
        return "_oed_cu"
    return "_oed_cf" if fixed else "_oed"


class OEDLayer:
    """OEDLayer{DISCRETE}(layer; params, measurements, observed) (oed.jl:77-180).  `layer` is a Single- or
    MultipleShootingLayer over a base model of OED_BASE; `params` / `observed` are fixed by the compiled-in
    augmentation (the reference's test problems: params = [2, 3], observed = u[1:2] for Lotka)."""

    def __init__(self, layer, discrete: bool = False, measurements=()):
        single = isinstance(layer, SingleShootingLayer)
        base = layer if single else layer.layer
        prob = base.problem
        if prob.model not in OED_BASE:
            raise KeyError(f"no compiled-in OED augmentation of model {prob.model!r}; available: {sorted(OED_BASE)}")
        params, nobs = OED_BASE[prob.model]
        measurements = list(measurements)
        self.DISCRETE = bool(discrete)
        self.SAMPLED = len(measurements) > 0
        self.FIXED = len(base.control_indices) == 0 and len(base.tunable_ic) == 0          # (:87,128)
        if self.SAMPLED and len(measurements) != nobs:
            raise ValueError(f"{nobs} observed functions need {nobs} measurement controls")
        model = prob.model + _variant(self.DISCRETE, self.SAMPLED, self.FIXED)
        if model not in MODEL_REGISTRY:
            raise KeyError(f"OED variant {model!r} is not compiled in")
        _, nx_aug, np_aug = MODEL_REGISTRY[model]
        bx, p_length = len(prob.u0), len(prob.p)
        nf = len(params)
        self.base_nx, self.n_fisher, self.n_obs = bx, nf, nobs
        samplings = list(range(1, len(measurements) + 1))
        ctrls = list(zip(base.control_indices, base.controls)) + [(p_length + k, m) for k, m in zip(samplings, measurements)]
        self.sampling_indices = [k + len(base.controls) for k in samplings]                # (:93,135)
        # "Replace the saveat with the sampling times" (:106-113): the measurement grids' points, else the ends of the horizon
        if self.SAMPLED:
            saveats = np.unique(np.concatenate([get_timegrid(m) for m in measurements]))
        else:
            saveats = np.asarray(prob.tspan, dtype=np.float64)
        newprob = prob.remake(model=model, u0=np.concatenate([prob.u0, np.zeros(nx_aug - bx)]),
                              p=np.concatenate([prob.p, np.ones(np_aug - p_length)]),
                              kwargs=dict(prob.kwargs, saveat=[float(t) for t in saveats]))
        if single:
            lb, ub = (np.concatenate([b, np.zeros(nx_aug - bx)]) for b in base.bounds_ic)  # (:117-122)
            self.layer = SingleShootingLayer(newprob, base.algorithm, controls=ctrls, tunable_ic=list(base.tunable_ic),
                                             bounds_ic=(lb, ub), state_initialization=base.state_initialization,
                                             bounds_p=base.bounds_p, parameter_initialization=base.parameter_initialization,
                                             quadrature_indices=[], device=base.device)
        else:
            pts = [t[0] for t in layer.shooting_intervals]
            self.layer = MultipleShootingLayer(newprob, base.algorithm, *pts, controls=ctrls,
                                               tunable_ic=list(base.tunable_ic),
                                               state_initialization=base.state_initialization, bounds_p=base.bounds_p,
                                               parameter_initialization=base.parameter_initialization,
                                               quadrature_indices=[], device=base.device)
        self.single = single
        self.model = model
        if self.DISCRETE:
            self.fim_mode = FIM_DISCRETE_SAMPLED if self.SAMPLED else FIM_DISCRETE
        else:
            self.fim_mode = FIM_FIXED_CONTINUOUS if (self.SAMPLED and self.FIXED) else FIM_END
        if self.fim_mode == FIM_FIXED_CONTINUOUS and not single:
            raise NotImplementedError("OEDLayer{false, true, true} reads ps.controls of ONE layer (oed.jl:352-362): single shooting only")

    def __repr__(self):
        return (f"{'Fixed ' if self.FIXED else ''}OEDLayer {'with' if self.SAMPLED else 'without'} "
                f"{'discrete' if self.DISCRETE else 'continuous'} measurement model and {len(self.sampling_indices)} observed functions.")

    def n_observed(self):
        return len(self.sampling_indices)

    def _base_layer(self):
        return self.layer if self.single else self.layer.layer

    # ---- LuxCore triple
    def initialparameters(self, rng):
        return self.layer.initialparameters(rng)

    def _weighting_grid(self, tspan, restrict: bool):
        """oed.jl:291-311 (single shooting: restrict = False) / :341-357 (per interval of a multiple-shooting layer)"""
        controls = self._base_layer().controls
        grids = [get_timegrid(c) for c in controls]
        overall = np.unique(np.concatenate(grids + [np.asarray(tspan, dtype=np.float64)]))
        if restrict:
            overall = overall[(overall >= tspan[0]) & (overall < tspan[1])]
        sidx = [s - 1 for s in self.sampling_indices]
        observed_grid = [np.nonzero(np.isin(overall, np.unique(grids[s])))[0] for s in sidx]
        mi = _flatten_chunks(build_index_grid(*controls, tspan=tspan))
        mi = np.atleast_2d(mi)
        uniq = lambda row: row[np.sort(np.unique(row, return_index=True)[1])]   # Julia's unique keeps first-seen order
        measurement_indices = [uniq(mi[s]) for s in sidx]
        rows = np.zeros((len(overall), len(sidx)), dtype=np.int64)
        for j, og in enumerate(observed_grid):
            rows[og, j] = measurement_indices[j][: len(og)]
        return rows, [uniq(r) for r in mi]

    def initialstates(self, rng):
        st = self.layer.initialstates(rng)
        nf, nobs = self.n_fisher, self.n_obs
        first = st if self.single else next(iter(st.values()))
        first["F_init"] = np.zeros((nf, nf))
        fim = dict(offset=0, n=nf, mode=self.fim_mode, base_nx=self.base_nx, n_observed=nobs, rows=None)
        if self.fim_mode in (FIM_END, FIM_FIXED_CONTINUOUS):
            fim["offset"] = self.base_nx * (1 + nf) + 1
        if self.fim_mode == FIM_DISCRETE_SAMPLED:
            if self.single:
                grid, act = self._weighting_grid(self._base_layer().problem.tspan, False)
                st["observation_grid"], st["active_controls"] = grid, act
                fim["rows"] = [grid]
            else:
                fim["rows"] = []
                for sti in st.values():
                    tsp = _flatten_chunks(sti["tspans"])
                    grid, act = self._weighting_grid((tsp[0][0], tsp[-1][1]), True)
                    sti["observation_grid"], sti["active_controls"] = grid, act
                    fim["rows"].append(grid)
        elif self.fim_mode == FIM_FIXED_CONTINUOUS:
            # active_controls = unique index-grid values of each measurement control (:304,312-318); the read-out
            # zips row j of hcat(controls[active]...) with the j-th increment of F (:357-362)
            controls = self._base_layer().controls
            mi = np.atleast_2d(_flatten_chunks(build_index_grid(*controls, tspan=self._base_layer().problem.tspan)))
            uniq = lambda row: row[np.sort(np.unique(row, return_index=True)[1])]
            act = [uniq(mi[s - 1]) for s in self.sampling_indices]
            if len({len(a) for a in act}) != 1:
                raise ValueError("reduce(hcat, ...) of measurement controls with different lengths (oed.jl:357)")
            st["active_controls"] = act
            n_pieces = len(_flatten_chunks(st["tspans"]))
            fim["rows"] = [np.stack(act, axis=1)[: n_pieces]]
        first["_fim"] = fim
        return st

    def setup(self, rng):
        return self.initialparameters(rng), self.initialstates(rng)

    def native(self, ps, st) -> native.NativeLayer:
        return self.layer._native(st, ps)

    # ---- Fisher information for a batch of candidates (oed.jl:381-461), one cb200_eval
    def fisher_information(self, X, ps_like, st, add_initial: bool = True, criteria: bool = False):
        """X: (B, nvars) flat candidates.  Returns F (B, n, n) [, criteria (B, 6)]."""
        nat = self.native(ps_like, st)
        first = st if self.single else next(iter(st.values()))
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        want = native.WANT_FIM | (native.WANT_CRIT if criteria else 0)
        r = nat.eval_host(X, want, fim_init=first["F_init"] if add_initial else None)
        native.check_status(r, "OEDLayer.fisher_information")
        F = r["fim"].reshape(X.shape[0], self.n_fisher, self.n_fisher)
        return (F, r["criteria"]) if criteria else F

    # ---- read-outs of the saved samples (oed.jl:464-534), batched: X (B, nvars)
    def _samples(self, X, ps_like, st):
        nat = self.native(ps_like, st)
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        r = nat.eval_host(X, native.WANT_TRAJ)
        return r["traj"].reshape(X.shape[0], nat.n_samples, nat.n_row)

    def sensitivities(self, X, ps_like, st):
        """sensitivities(oed, traj) (oed.jl:464-470): G(t_k) = dx/dp at every saved sample, (B, n_samples, nx_base, n_fisher)"""
        tr = self._samples(X, ps_like, st)
        bx, nf = self.base_nx, self.n_fisher
        G = tr[..., bx:bx + bx * nf].reshape(tr.shape[:-1] + (nf, bx))       # vec(G) is column-major
        return np.swapaxes(G, -1, -2)

    def observed_equations(self, X, ps_like, st):
        """observed_equations(oed, traj) (oed.jl:481-488): the observed functions = the leading states, (B, n_samples, n_obs)"""
        return self._samples(X, ps_like, st)[..., :max(self.n_obs, OED_BASE[self.model.split("_oed")[0]][1])]

    def local_information_gain(self, X, ps_like, st):
        """local_information_gain (oed.jl:500-510): (h_i G(t_k))'(h_i G(t_k)) per sample and observed function"""
        from .criteria import local_information_gain
        nobs = OED_BASE[self.model.split("_oed")[0]][1]
        return local_information_gain(self._samples(X, ps_like, st), self.base_nx, self.n_fisher, nobs)

    def global_information_gain(self, X, ps_like, st):
        """global_information_gain (oed.jl:520-534): the same with rows scaled by inv(F(t_f)), F from the layer's read-out"""
        from .criteria import global_information_gain
        nobs = OED_BASE[self.model.split("_oed")[0]][1]
        F = self.fisher_information(X, ps_like, st)
        return global_information_gain(self._samples(X, ps_like, st), F, self.base_nx, self.n_fisher, nobs)

    def update_fim(self, experiments, st):
        """update_fim (oed.jl:166-186): fold finished experiments (flat candidates, rows of `experiments`) into F_init"""
        first = st if self.single else next(iter(st.values()))
        F = self.fisher_information(experiments["X"], experiments["ps"], experiments["st"], add_initial=False)
        first["F_init"] = first["F_init"] + F.sum(axis=0)
        return st

    def fisher_gradient(self, X, ps_like, st, add_initial: bool = True):
        """F (B, n, n) and dF/d(flat variables) (B, n, n, nvars) -- what ForwardDiff through `fisher_information` yields.
        One cb200_eval per call: the samples and their sensitivities (CB200_WANT_TRAJ_SENS) come from the device, the
        product rule of the read-out (oed.jl:2-16,352-362,446-461) is host arithmetic.  The end-state read-out goes
        through criteria.criteria_gradient instead."""
        from .multiple_shooting import trajectory_jacobian
        if self.fim_mode == FIM_END:
            raise ValueError("end-state read-out: use criteria.criteria_gradient")
        nat = self.native(ps_like, st)
        first = st if self.single else next(iter(st.values()))
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        r = nat.eval_host(X, native.WANT_FIM | native.WANT_TRAJ | native.WANT_TRAJ_SENS | native.WANT_SENS,
                          fim_init=first["F_init"] if add_initial else None)
        B, n, bx, nobs = X.shape[0], self.n_fisher, self.base_nx, self.n_obs
        F = r["fim"].reshape(B, n, n)
        dF = np.zeros((B, n, n, nat.nvars))
        blocks = nat.block_structure()
        rows = []                                           # (sample, observed k, flat variable of the weight)
        if self.fim_mode != FIM_DISCRETE:
            from .multiple_shooting import sample_pieces
            sts = [st] if self.single else list(st.values())
            pss = [ps_like] if self.single else list(ps_like.values())
            pieces = sample_pieces(self._base_layer(), sts)
            s0 = 0
            for i, (sti, psi, grid) in enumerate(zip(sts, pss, first["_fim"]["rows"])):
                c0 = int(blocks[i]) + len(psi["u0"]) + len(psi["p"])
                rows += [(s0 + j, k, c0 + int(v) - 1) for j, row in enumerate(grid) for k, v in enumerate(row) if v > 0]
                s0 += len(pieces[i]) - (1 if i == len(sts) - 1 else 0)
        gi = bx + bx * np.arange(n)                         # state rows of g_k: gi + k
        ntri = n * (n + 1) // 2
        tri = np.array([[max(a, b) * (max(a, b) + 1) // 2 + min(a, b) for b in range(n)] for a in range(n)])
        f_off = bx * (1 + n)
        for b in range(B):
            u = r["traj"][b].reshape(nat.n_samples, nat.n_row).T
            J = trajectory_jacobian(self.layer, st, ps_like, r, b)             # (n_row, n_samples, nvars)
            if self.fim_mode == FIM_DISCRETE:
                s = sum(u[gi + k] for k in range(nobs))                         # (n, n_samples)
                ds = sum(J[gi + k] for k in range(nobs))                        # (n, n_samples, nvars)
                dF[b] = np.einsum("asv,bs->abv", ds, s) + np.einsum("as,bsv->abv", s, ds)
                continue
            for smp, k, var in rows:
                w = X[b, var]
                if self.fim_mode == FIM_DISCRETE_SAMPLED:
                    g, dg = u[gi + k, smp], J[gi + k, smp]
                    dF[b] += w * w * (dg[:, None, :] * g[None, :, None] + g[:, None, None] * dg[None, :, :])
                    dF[b, :, :, var] += 2.0 * w * np.outer(g, g)
                else:
                    q = f_off + k * ntri + tri
                    dF[b] += w * (J[q, smp + 1] - J[q, smp])
                    dF[b, :, :, var] += u[q, smp + 1] - u[q, smp]
        return F, dF

    def criteria_gradient(self, X, ps_like, st, add_initial: bool = True):
        """the six criteria (B, 6) and their gradients (B, 6, nvars) for any variant"""
        from .criteria import batched_criteria, criteria_chain, criteria_gradient
        first = st if self.single else next(iter(st.values()))
        if self.fim_mode == FIM_END:
            nat = self.native(ps_like, st)
            fi = first["F_init"] if add_initial else None
            return batched_criteria(nat, X, fim_init=fi)[0], criteria_gradient(nat, X, fim_init=fi)
        F, dF = self.fisher_gradient(X, ps_like, st, add_initial)
        crit = self.fisher_information(X, ps_like, st, add_initial, criteria=True)[1]
        return crit, np.stack([criteria_chain(F[b], dF[b]) for b in range(len(F))])

    def get_sampling_sums(self, ps, st):
        """get_sampling_sums (oed.jl:536-640): one number per measurement control -- the number of (weighted)
        measurements for the discrete variants (:573-580), the measured time sum_j w_j dt_j for the continuous ones
        (:582-589,610-618); a multiple-shooting layer sums over its intervals (:559-563)."""
        if not self.SAMPLED:
            return np.zeros(0)
        pairs = [(ps, st)] if self.single else [(ps[k], st[k]) for k in st]
        out = np.zeros(self.n_obs)
        for psi, sti in pairs:
            c = np.asarray(psi["controls"], dtype=np.float64)
            if self.DISCRETE:
                for k, s in enumerate(self.sampling_indices):
                    out[k] += c[np.asarray(sti["active_controls"][s - 1]) - 1].sum()
                continue
            tsp = _flatten_chunks(sti["tspans"])
            dts = np.array([b - a for a, b in tsp])
            if self.FIXED:
                for k, sub in enumerate(sti["active_controls"]):
                    out[k] += (c[np.asarray(sub) - 1] * dts).sum()
            else:
                grid = np.atleast_2d(np.asarray(_flatten_chunks(sti["index_grid"]), dtype=np.int64))
                for k, s in enumerate(self.sampling_indices):
                    out[k] += (c[grid[s - 1] - 1] * dts).sum()
        return out


class MultiExperimentLayer:
    """MultiExperimentLayer{DISCRETE}(prob, alg, nexp; ...) (lib/CorleoneOED/src/multiexperiments.jl:9-16,66-74,88-91): `n_exp`
    experiments of ONE OED layer, jointly designed; parameters and states are grouped by experiment (`experiment_i`), the
    Fisher information is the sum over the experiments plus F_init of the first (:279-303).

    On the device the experiment axis is just more of the batch axis: a batch of B joint designs is evaluated as
    B * n_exp candidates of the underlying layer in ONE cb200_eval and summed per design.  The variant that splits the
    parameter set among the experiments (:76-86,94-107) needs one augmentation per parameter subset and is not compiled
    in."""

    def __init__(self, oed: OEDLayer, n_exp: int):
        if n_exp < 1:
            raise ValueError("n_exp must be >= 1")
        self.layers, self.n_exp = oed, int(n_exp)
        self.names = [f"experiment_{i + 1}" for i in range(self.n_exp)]

    def __repr__(self):
        o = self.layers
        return (f"{'Fixed ' if o.FIXED else ''}MultiExperimentLayer with {'discrete' if o.DISCRETE else 'continuous'} "
                f"measurement model and {self.n_exp} experiments.")

    def n_observed(self):
        return self.n_exp * self.layers.n_observed()                                         # (:189)

    def initialparameters(self, rng):
        return {k: self.layers.initialparameters(rng) for k in self.names}                   # (:120-124)

    def initialstates(self, rng):
        return {k: self.layers.initialstates(rng) for k in self.names}                       # (:166-170)

    def setup(self, rng):
        return self.initialparameters(rng), self.initialstates(rng)

    def get_bounds(self):
        from .multiple_shooting import get_bounds
        lo, hi = get_bounds(self.layers.layer)
        return {k: lo for k in self.names}, {k: hi for k in self.names}                      # (:318-324)

    def get_block_structure(self):
        from .multiple_shooting import get_block_structure
        b = np.asarray(get_block_structure(self.layers.layer), dtype=np.int64)               # (:339-351)
        return np.concatenate([b if i == 0 else b[1:] + i * b[-1] for i in range(self.n_exp)])

    def get_number_of_shooting_constraints(self, ps, st):
        if self.layers.single:
            return 0
        from .multiple_shooting import get_number_of_shooting_constraints
        k = self.names[0]
        return self.n_exp * get_number_of_shooting_constraints(self.layers.layer, ps[k], st[k])   # (:190)

    def get_sampling_sums(self, ps, st):
        return np.concatenate([self.layers.get_sampling_sums(ps[k], st[k]) for k in self.names])   # (:249-255)

    def fisher_information(self, X, ps_like, st, add_initial: bool = True):
        """X: (B, n_exp * nvars) joint designs, experiments in order.  Returns F (B, n, n) = F_init + sum_e F_e (:279-289)."""
        k0 = self.names[0]
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        nv = len(to_vec(ps_like[k0]))
        assert X.shape[1] == self.n_exp * nv, (X.shape, self.n_exp, nv)
        F = self.layers.fisher_information(X.reshape(X.shape[0] * self.n_exp, nv), ps_like[k0], st[k0], add_initial=False)
        F = F.reshape(X.shape[0], self.n_exp, *F.shape[1:]).sum(axis=1)
        if add_initial:
            first = st[k0] if self.layers.single else next(iter(st[k0].values()))
            F = F + first["F_init"]
        return F

    def update_fim(self, experiments, st):
        """update_fim (:195-221): fold finished joint designs (rows of experiments["X"]) into F_init of the first experiment"""
        k0 = self.names[0]
        first = st[k0] if self.layers.single else next(iter(st[k0].values()))
        F = self.fisher_information(experiments["X"], experiments["ps"], experiments["st"], add_initial=False)
        first["F_init"] = first["F_init"] + F.sum(axis=0)
        return st


def sampling_sum_matrix(oed: OEDLayer, ps_like, st) -> np.ndarray:
    """get_sampling_sums is linear in the variables for every variant: its matrix (n_observed, nvars)."""
    from .single_shooting import from_vec
    x0 = to_vec(ps_like)
    A = np.zeros((oed.n_obs if oed.SAMPLED else 0, len(x0)))
    for v in range(len(x0)):
        e = np.zeros(len(x0))
        e[v] = 1.0
        A[:, v] = oed.get_sampling_sums(from_vec(e, ps_like), st)
    return A


def solve_slsqp(oed: OEDLayer, ps_like, st, criterion: int = 0, M=None, x0=None, maxiter=300, ftol=1e-12):
    """OptimizationProblem(oed, criterion; M) (lib/CorleoneOED/ext/CorleoneOEDOptimizationExtension.jl) with SciPy's
    SLSQP standing in for Ipopt: minimise one criterion (index into criteria.ORDER) within the variable bounds subject
    to sampling sums <= M; value and gradient from the device (criteria_gradient).  Single shooting."""
    from scipy.optimize import Bounds, LinearConstraint, minimize
    from .multiple_shooting import get_bounds
    x0 = to_vec(ps_like) if x0 is None else np.asarray(x0, dtype=np.float64)
    lo, hi = get_bounds(oed.layer)
    cache = {}

    def ev(x):
        key = x.tobytes()
        if cache.get("key") != key:
            c, g = oed.criteria_gradient(x[None, :], ps_like, st)
            cache.update(key=key, val=(float(c[0, criterion]), g[0, criterion].copy()))
        return cache["val"]

    cons = []
    if M is not None and oed.SAMPLED:
        cons.append(LinearConstraint(sampling_sum_matrix(oed, ps_like, st), -np.inf, np.asarray(M, dtype=np.float64)))
    return minimize(lambda x: ev(x)[0], x0, jac=lambda x: ev(x)[1], bounds=Bounds(to_vec(lo), to_vec(hi)),
                    constraints=cons, method="SLSQP", options=dict(maxiter=maxiter, ftol=ftol))


def flat(ps) -> np.ndarray:
    return to_vec(ps)
