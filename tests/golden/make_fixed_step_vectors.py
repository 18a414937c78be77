#!/usr/bin/env python
"""Generate tests/golden/fixed_step_vectors.json: fixed-step (and one adaptive) input/output vectors of the hot path computed by the
CPU oracle (oracle/corleone_oracle.c) on seeded inputs.  The reference itself is Julia and cannot run in the
build container, so these are ORACLE outputs (the oracle is pinned on the reference's own goldens in
tests/test_oracle_golden.py); committing them guards against the oracle and the CUDA path drifting together.

    python tests/golden/make_fixed_step_vectors.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import problems as P  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def case(name, spec_fn, method, K, B, seed, objective_row, scale=0.05):
    rng = np.random.default_rng(seed)
    spec = spec_fn(rng)
    O = po.OracleLayer(spec)
    ps = O.default_ps()[None, :] + scale * rng.standard_normal((B, O.nvars))
    out = dict(name=name, method=method, K=K, seed=seed, objective_row=objective_row, scale=scale, B=B,
               nvars=O.nvars, ps=ps.tolist(), xend=[], defk=[], defs=[], objective=[], grad=[], sens_blocks=[])
    if spec.get("saveat") is not None:   # cases with saveat also carry the merged trajectory samples and their times
        out.update(saveat=list(spec["saveat"]), traj=[], t=None)
    if method == "tsit5_adaptive":   # the reference's default mode at the tolerances of its core tests
        out.update(abstol=1.0e-8, reltol=1.0e-6)
        po.set_tolerances(out["abstol"], out["reltol"])
    blocks = O.block_structure()
    for b in range(B):
        o = O.eval(ps[b], method=method, K=K)
        Jx, Jd, Jl = O.jacobians(ps[b], method=method, K=K)
        out["xend"].append(o["xend"].T.ravel().tolist())          # [i][k]
        out["defk"].append(o["defk"].tolist())
        out["defs"].append(o["defs"].tolist())
        out["objective"].append(float(o["last"][objective_row - 1]))
        if "traj" in out:
            out["traj"].append(o["u"].T.ravel().tolist())          # [sample][row] as cb200_out.traj
            out["t"] = o["t"].tolist()
        out["grad"].append(Jl[objective_row - 1].tolist())
        sens = []
        for i in range(O.I):   # rows (interval, direction, state) as cb200_out.sens
            blk = Jx[O.nx * i:O.nx * (i + 1), blocks[i]:blocks[i + 1]]
            sens += blk.T.ravel().tolist()
        out["sens_blocks"].append(sens)
    return out


SAVEAT = [0.013, 0.55, 1.2345, 2.95, 3.0, 3.05, 4.44, 6.0001, 7.77, 8.9999, 9.0, 11.95]

CASES = [
    ("lotka_ms4_tsit5", lambda r: P.lotka_spec(controls=r.uniform(0, 1, 120)), "tsit5", 2, 2, 101, 3),
    ("lotka_ms4_quadrature_rk4", lambda r: P.lotka_spec(controls=r.uniform(0, 1, 120), quadrature=True), "rk4", 3, 2, 102, 3),
    ("lqr_ms10_tsit5", lambda r: P.lqr_spec(10, r), "tsit5", 2, 2, 103, 2),
    ("egerstedt_ss_tsit5", lambda r: P.egerstedt_spec(), "tsit5", 2, 2, 104, 3),
    ("dense20_ms3_tsit5", lambda r: P.dense20_spec(3, r), "tsit5", 2, 2, 105, 1),
    ("lotka_ms4_adaptive", lambda r: P.lotka_spec(controls=r.uniform(0, 1, 120)), "tsit5_adaptive", 1, 2, 106, 3),
    # round 2: saveat times inside pieces (samples from Tsit5's dense output), fixed step and adaptive; the dense family at 24 states
    ("lotka_ms4_saveat_tsit5", lambda r: dict(P.lotka_spec(controls=r.uniform(0, 1, 120), quadrature=True), saveat=SAVEAT), "tsit5", 3, 2, 107, 3),
    ("lotka_ms4_saveat_adaptive", lambda r: dict(P.lotka_spec(controls=r.uniform(0, 1, 120)), saveat=SAVEAT), "tsit5_adaptive", 1, 2, 108, 3),
    ("dense24_ms3_tsit5", lambda r: P.dense20_spec(3, r, nx=24), "tsit5", 2, 2, 109, 1),
]

if __name__ == "__main__":
    data = [case(*c) for c in CASES]
    with open(os.path.join(HERE, "fixed_step_vectors.json"), "w") as f:
        json.dump(data, f)
    print("wrote", len(data), "cases,", os.path.getsize(os.path.join(HERE, "fixed_step_vectors.json")), "bytes")
