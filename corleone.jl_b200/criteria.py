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
