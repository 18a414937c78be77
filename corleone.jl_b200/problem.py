"""ODEProblem / algorithm stand-ins for the SciMLBase objects the reference layers are built from.

The reference integrates arbitrary Julia closures with OrdinaryDiffEq (src/single_shooting.jl:443-453).
Here the right-hand side is one of the models compiled into libcorleone_b200.so (csrc/models.cuh) and the
algorithm is a fixed-tableau explicit RK run with a fixed number of steps per control piece -- the
`adaptive=false, dt=h` mode of the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# name -> (cb200_model id, nx, np_all)
MODEL_REGISTRY = {
    "lotka3": (0, 3, 3),
    "lqr": (1, 2, 3),
    "lotka2": (2, 2, 3),
    "lotka2_oed": (3, 9, 5),
    "lin1d": (4, 1, 1),
    "lin1d_oed": (5, 3, 2),
    "dense20": (6, 20, 1),
    "egerstedt": (7, 3, 3),
    "vdp3": (8, 3, 1),
    "cartpole5": (9, 5, 1),
    "rocket3": (10, 3, 2),
    "quadrotor7": (11, 7, 4),
    "threetank4": (12, 4, 3),
    "batchreactor2": (13, 2, 1),
    "lotkacomp3": (14, 3, 1),
    "f8aircraft4": (15, 4, 2),
    "doubleosc6": (16, 6, 1),
    "lotka2_oed_cu": (17, 9, 3),
    "lotka2_oed_cf": (18, 12, 5),
    "lotka2_oed_ds": (19, 6, 3),
    "lotka2_oed_du": (20, 6, 3),
    "lin1d_oed_cf": (21, 3, 2),
    "lin1d_oed_ds": (22, 2, 1),
    "brysondenham3": (23, 3, 1),
    "catalystmixing2": (24, 2, 1),
    "cushioned3": (25, 3, 2),
    "denbigh3": (26, 3, 1),
    "dielectrophoretic3": (27, 3, 2),
    "electriccar4": (28, 4, 1),
    "fuller3": (29, 3, 1),
    "jackson3": (30, 3, 1),
    "mountaincar3": (31, 3, 2),
    "particlesteering5": (32, 5, 2),
    "raomease2": (33, 2, 1),
    "robbins4": (34, 4, 1),
    "moonlanding3": (35, 3, 2),
    "lotkashared4": (36, 4, 1),
    "hangingchain3": (37, 3, 1),
    "tubularreactor2": (38, 2, 1),
    "ocean4": (39, 4, 2),
    "ductedfan7": (40, 7, 3),
    "hangglider4": (41, 4, 2),
    "dense16": (42, 16, 1),
    "dense24": (43, 24, 1),
    "dense32": (44, 32, 1),
}

# base model -> (augmented model, Fisher parameter indices (1-based), number of observed functions)
OED_AUGMENTED = {
    "lotka2": ("lotka2_oed", (2, 3), 2),
    "lin1d": ("lin1d_oed", (1,), 1),
}


@dataclass
class ODEProblem:
    """ODEProblem(f, u0, tspan, p): `model` names the compiled-in right-hand side."""
    model: str
    u0: np.ndarray
    tspan: tuple
    p: np.ndarray
    model_data: np.ndarray | None = None
    kwargs: dict = field(default_factory=dict)

    def __post_init__(self):
        if self.model not in MODEL_REGISTRY:
            raise KeyError(f"unknown model {self.model!r}; compiled-in models: {sorted(MODEL_REGISTRY)}")
        self.u0 = np.array(self.u0, dtype=np.float64).ravel()
        self.p = np.array(self.p, dtype=np.float64).ravel()
        self.tspan = (float(self.tspan[0]), float(self.tspan[1]))
        _, nx, npar = MODEL_REGISTRY[self.model]
        if len(self.u0) != nx or len(self.p) != npar:
            raise ValueError(f"model {self.model!r} has {nx} states and {npar} parameters; "
                             f"got {len(self.u0)} and {len(self.p)}")
        if self.model_data is not None:
            self.model_data = np.ascontiguousarray(self.model_data, dtype=np.float64).ravel()

    def remake(self, **kw):
        d = dict(model=self.model, u0=self.u0.copy(), tspan=self.tspan, p=self.p.copy(),
                 model_data=self.model_data, kwargs=dict(self.kwargs))
        d.update(kw)
        return ODEProblem(**d)


class Tsit5:
    """Tsitouras 5(4).

    Tsit5()                    the reference's own call: adaptive, OrdinaryDiffEq's step-size control with the problem's
                               `abstol` / `reltol` keyword arguments (ODEProblem(...; abstol, reltol); defaults 1e-6 / 1e-3)
    Tsit5(K) / Tsit5(steps_per_piece=K)
                               fixed step: K steps h = (t1-t0)/K on every control piece -- `adaptive=false, dt=h`, the mode
                               BASELINE.json benchmarks
    Tsit5(adaptive=True)       adaptive, explicitly"""

    def __init__(self, steps_per_piece: int | None = None, adaptive: bool | None = None):
        if adaptive is None:
            adaptive = steps_per_piece is None
        if adaptive and steps_per_piece not in (None, 1):
            raise ValueError("steps_per_piece has no meaning for the adaptive mode")
        self.steps_per_piece = 1 if steps_per_piece is None else int(steps_per_piece)
        self.adaptive = bool(adaptive)
        if self.steps_per_piece < 1:
            raise ValueError("steps_per_piece must be >= 1")

    @property
    def method_id(self) -> int:
        return 2 if self.adaptive else 0

    def __repr__(self):
        return "Tsit5()" if self.adaptive else f"Tsit5(steps_per_piece={self.steps_per_piece})"

    def __eq__(self, other):
        return isinstance(other, Tsit5) and (self.steps_per_piece, self.adaptive) == (other.steps_per_piece, other.adaptive)

    def __hash__(self):
        return hash(("Tsit5", self.steps_per_piece, self.adaptive))


@dataclass(frozen=True)
class RK4:
    """Classical Runge-Kutta 4, fixed step."""
    steps_per_piece: int = 1
    method_id: int = 1
