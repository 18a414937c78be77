"""Trajectory -- result container of the shooting layers (mirror of src/trajectory.jl).

`u` is an (n_samples, n_row) array for one candidate (so `traj.u[k]` is the k-th sample vector
[x(t_k); controls active on the piece ending at t_k], src/single_shooting.jl:456) or (B, n_samples, n_row)
for a batch.  The device produces the structure-of-arrays form; this class only wraps it.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_SUB = str.maketrans("0123456789", "₀₁₂₃₄₅₆₇₈₉")


def subscript(i: int) -> str:
    """single_shooting.jl:221"""
    return str(i).translate(_SUB)


@dataclass
class SymbolCache:
    """variables: name -> 0-based row of a sample; parameters: name -> 0-based index into ps.p"""
    variables: dict
    parameters: dict
    independent: tuple = ("t",)


def retrieve_symbol_cache(nx: int, np_all: int, control_indices, control_names) -> SymbolCache:
    """single_shooting.jl:223-253 for problems without a symbolic system: states x₁.., tunable
    parameters p₁.., controls by name; sample rows are [states; controls in layer order]."""
    control_indices = list(control_indices)
    state_symbols = [f"x{subscript(i)}" for i in range(1, nx + 1)]
    u_id = p_id = 0
    psyms = []
    for i in range(1, np_all + 1):
        if i in control_indices:
            psyms.append(control_names[u_id])
            u_id += 1
        else:
            p_id += 1
            psyms.append(f"p{subscript(p_id)}")
    variables = {s: k for k, s in enumerate(state_symbols + [psyms[i - 1] for i in control_indices])}
    nonidx = [i for i in range(1, np_all + 1) if i not in control_indices]
    parameters = {psyms[i - 1]: k for k, i in enumerate(nonidx)}
    return SymbolCache(variables, parameters)


@dataclass
class Trajectory:
    """src/trajectory.jl:9-22"""
    sys: SymbolCache
    u: np.ndarray
    p: np.ndarray
    t: np.ndarray
    shooting: dict = field(default_factory=dict)
    shooting_indices: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int64))

    def __getitem__(self, sym: str):
        """src/trajectory.jl:42-49"""
        if sym in self.sys.variables:
            return self.u[..., self.sys.variables[sym]]
        if sym in self.sys.parameters:
            return self.p[..., self.sys.parameters[sym]]
        if sym == "t":
            return self.t
        raise KeyError(f"Invalid index: :{sym}")


def is_shooting_solution(traj: Trajectory) -> bool:
    return bool(traj.shooting)


def shooting_violations(traj: Trajectory):
    return traj.shooting
