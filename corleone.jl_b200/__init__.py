"""corleone_b200 -- B200-native drop-in for Corleone.jl's batched shooting hot path.

Host side mirrors the reference layer API (SingleShootingLayer / MultipleShootingLayer, ControlParameter,
node initialisers, Trajectory); every evaluation runs in libcorleone_b200.so (hand-written sm_100a kernels
behind a C ABI, include/corleone_b200.h).  No CPU fallback.
"""
from .controls import (ControlParameter, build_index_grid, check_consistency, collect_local_control_bounds,
                       collect_local_controls, collect_tspans, control_length, default_bounds, default_u,
                       find_shooting_indices, get_controls, get_subvector_indices, get_timegrid)
from .controls import get_bounds as get_control_bounds
from .multiple_shooting import (MultipleShootingLayer, control_matchings, default_initialization, get_block_structure,
                                get_bounds, get_number_of_control_matchings, get_number_of_parameter_matchings,
                                get_number_of_shooting_constraints, get_number_of_state_matchings,
                                parameter_matchings, shooting_constraints, shooting_constraints_,
                                stage_ordered_shooting_constraints, stage_ordered_shooting_constraints_,
                                state_matchings, trajectory_jacobian)
from .node_initialization import (constant_initialization, custom_initialization, forward_initialization,
                                  forward_initialization_batch,
                                  hybrid_initialization, linear_initialization, random_initialization)
from .problem import MODEL_REGISTRY, ODEProblem, RK4, Tsit5
from .single_shooting import InitialConditionRemaker, SingleShootingLayer, from_vec, to_vec
from .trajectory import Trajectory, is_shooting_solution, shooting_violations
from . import native
from .dynprob import CorleoneDynamicOptProblem, solve_slsqp
from .criteria import (ACriterion, DCriterion, ECriterion, FisherACriterion, FisherDCriterion, FisherECriterion,
                       batched_criteria, criteria_gradient, global_information_gain, local_information_gain,
                       sampling_sum_matrix, solve_oed_slsqp)
from . import oed
from .oed import MultiExperimentLayer, OEDLayer


def setup(rng, layer):
    """LuxCore.setup(rng, layer) -> (ps, st)"""
    return layer.initialparameters(rng), layer.initialstates(rng)
