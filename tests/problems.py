"""Problem specs shared by the oracle (oracle/pyoracle.py) and the product layers (corleone_b200).

The spec builders live in corleone_b200.configs (plain data); the oracle consumes the same dicts.
"""
import numpy as np

from corleone_b200.configs import (config1_lotka_ms24, config2_lqr, config3_lotka_oed, config4_dense20,  # noqa: F401
                                   dense20_data, egerstedt_spec, layer_from_spec, lin1d_oed_spec,
                                   lotka_oed_spec, lotka_spec, native_from_spec, cartpole_spec, rocket_spec, vdp_spec,
                                   batchreactor_spec, quadrotor_spec, threetank_spec, doubleosc_spec,
                                   f8_spec, lotkacomp_spec, bench_spec, BENCH_TABLE)


def lotka_ms24_spec(rng):
    return config1_lotka_ms24(rng)


def lqr_spec(n_intervals=100, rng=None):
    return config2_lqr(n_intervals, rng)


def dense20_spec(n_intervals=8, rng=None, nx=20):
    return config4_dense20(n_intervals, rng, nx=nx)


def product_layer(spec, alg, **kw):
    return layer_from_spec(spec, alg, **kw)
