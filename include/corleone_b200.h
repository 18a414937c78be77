/*
 * corleone_b200.h -- C ABI of libcorleone_b200.so: the B200-native replacement of Corleone.jl's
 * data-parallel hot path (batched integration of every multiple-shooting interval with forward
 * sensitivities, for a whole batch of candidates).
 *
 * The reference has no FFI; its plug-in points are Julia dispatch hooks.  Each entry point below names
 * the reference code it replaces (paths relative to SciML/Corleone.jl):
 *
 *   cb200_create   <- LuxCore.initialstates(rng, layer): the static `st` tables
 *                     (src/single_shooting.jl:274-352, src/multiple_shooting.jl:107-116) are passed in
 *                     as Julia produced them (Int64, 1-based) and uploaded once.
 *   cb200_eval     <- (shooting::MultipleShootingLayer)(u0, ps, st) = _parallel_solve + Trajectory merge
 *                     (src/multiple_shooting.jl:118-182), i.e. mythreadmap over intervals
 *                     (src/Corleone.jl:22-24) of _sequential_solve (src/single_shooting.jl:421-471),
 *                     for B candidates at once; plus what the NLP closures derive from the trajectory:
 *                     objective (src/dynprob.jl:68-73), shooting_constraints! (src/multiple_shooting.jl:
 *                     289-295), the ForwardDiff derivatives of both (src/dynprob.jl:143-164), and the
 *                     continuous Fisher information (lib/CorleoneOED/src/oed.jl:381-398).
 *   cb200_jac_structure <- sparsity of cons_j in get_block_structure order (src/multiple_shooting.jl:320-328).
 *   cb200_destroy, cb200_last_error, cb200_version, cb200_set_stream, cb200_set_interval_range: lifetime / plumbing (new).
 *
 * Conventions
 *   - plain C, no exceptions cross the boundary; every call returns 0 or a negative cb200_status;
 *     cb200_last_error() gives the thread-local message.
 *   - the caller owns every buffer it passes; the library never retains a caller pointer after return
 *     and owns all device memory behind the opaque handle.
 *   - there is NO CPU fallback: without a CUDA device cb200_create fails with CB200_ERR_CUDA.
 *   - all floating point data are FP64; index tables are Int64 and 1-based exactly as Julia holds them.
 */
#ifndef CORLEONE_B200_H
#define CORLEONE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_VERSION 201

typedef enum {
    CB200_OK = 0,
    CB200_ERR_ARG = -1,      /* bad descriptor / argument */
    CB200_ERR_CUDA = -2,     /* CUDA runtime failure (message has the CUDA error string) */
    CB200_ERR_UNSUPPORTED = -3, /* model/method/feature combination not compiled in */
    CB200_ERR_NOMEM = -4,
    CB200_ERR_COMM = -5      /* multi-GPU data plane: NCCL failure, a peer window that cannot be mapped, a rank that never arrived */
} cb200_status;

/* compiled-in model registry (the reference integrates arbitrary Julia closures; a C-ABI CUDA library
 * cannot -- see DESIGN.md "user RHS") */
typedef enum {
    CB200_MODEL_LOTKA3 = 0,     /* Lotka fishing + Lagrange state; examples/lotka_fishing_optimal_control/main.jl:38-42 */
    CB200_MODEL_LQR = 1,        /* examples/linear_quadratic/main.jl:45-50 */
    CB200_MODEL_LOTKA2 = 2,     /* lib/CorleoneOED/examples/lotka_oed/main.jl:61-65 */
    CB200_MODEL_LOTKA2_OED = 3, /* [x; G; F_triu] augmentation, params=[2,3], observed=u[1:2] (augmentation.jl:131-253) */
    CB200_MODEL_LIN1D = 4,      /* lib/CorleoneOED/test/core/1d_oed.jl:11-13 */
    CB200_MODEL_LIN1D_OED = 5,
    CB200_MODEL_DENSE20 = 6,    /* synthetic BASELINE config 4: x' = W x - x^3 + b u */
    CB200_MODEL_EGERSTEDT = 7,  /* test/core/local_controls.jl:46-52 */
    /* lib/OptimalControlBenchmarks/src/problems: right-hand sides with their Lagrange term as last state */
    CB200_MODEL_VDP3 = 8,       /* van_der_pol.jl:16-33 */
    CB200_MODEL_CARTPOLE5 = 9,  /* cart_pendulum.jl:17-48 */
    CB200_MODEL_ROCKET3 = 10,   /* goddarts_rocket.jl:16-40, p = (T, u) */
    CB200_MODEL_QUADROTOR7 = 11,    /* quadrotor.jl:9-52, p = (w1, w2, w3, u) */
    CB200_MODEL_THREETANK4 = 12,    /* three_tank.jl:9-53, p = (w1, w2, w3) */
    CB200_MODEL_BATCHREACTOR2 = 13, /* batch_reactor.jl:9-17, p = (u) */
    CB200_MODEL_LOTKACOMP3 = 14,    /* lotka_competitive.jl:9-31 */
    CB200_MODEL_F8AIRCRAFT4 = 15,   /* f8.jl:9-41, p = (T, u) */
    CB200_MODEL_DOUBLEOSC6 = 16,    /* double_oscillator.jl:9-38; time rides along as a state (tau' = 1) */
    /* the other variants of the OED augmentation (lib/CorleoneOED/src/augmentation.jl), Lotka: params=[2,3],
       observed=u[1:2]; 1-D: params=[1], observed=[u[1]] */
    CB200_MODEL_LOTKA2_OED_CU = 17, /* Continuous, no measurements (:217-220): [x; G; F_triu], F' = (sum_i h_i G)'(sum_i h_i G) */
    CB200_MODEL_LOTKA2_OED_CF = 18, /* ContinuousSampled, fixed (:255-292): [x; G; F_1 triu; F_2 triu], p = (base p; w) */
    CB200_MODEL_LOTKA2_OED_DS = 19, /* DiscreteSampled (:171-203): [x; G], p = base p; the measurement weights live in ps.controls only (filtered out of the index grid) and weight the samples at the measurement times (saveat) */
    CB200_MODEL_LOTKA2_OED_DU = 20, /* Discrete, no measurements (:171-203): [x; G], p = base p */
    CB200_MODEL_LIN1D_OED_CF = 21,
    CB200_MODEL_LIN1D_OED_DS = 22,
    /* the rest of lib/OptimalControlBenchmarks/src/problems (Lagrange terms and free-time clocks as last states) */
    CB200_MODEL_BRYSONDENHAM3 = 23,   /* bryson_denham.jl:9-36; states (x, v, cost), p = (w) */
    CB200_MODEL_CATALYSTMIXING2 = 24,   /* catalyst_mixing.jl:9-22; states (x1, x2), p = (w) */
    CB200_MODEL_CUSHIONED3 = 25,   /* cushioned_oscillation.jl:9-37; states (x, v, obj), p = (t_s, u) */
    CB200_MODEL_DENBIGH3 = 26,   /* denbigh.jl:9-37; states (x1, x2, x3), p = (T) */
    CB200_MODEL_DIELECTROPHORETIC3 = 27,   /* dielectrophoretic_particle.jl:9-34; states (x0, x1, obj), p = (t_s, u) */
    CB200_MODEL_ELECTRICCAR4 = 28,   /* electric_car.jl:9-62; states (x1, x2, x3, cost), p = (u) */
    CB200_MODEL_FULLER3 = 29,   /* fuller.jl:9-32; states (x0, x1, cost), p = (u) */
    CB200_MODEL_JACKSON3 = 30,   /* jackson.jl:9-50; states (x1, x2, x3), p = (u) */
    CB200_MODEL_MOUNTAINCAR3 = 31,   /* mountain_car.jl:9-31; states (x, v, obj), p = (t_s, u) */
    CB200_MODEL_PARTICLESTEERING5 = 32,   /* particle_steering.jl:9-42; states (x1, x2, dx1, dx2, obj), p = (t_s, u) */
    CB200_MODEL_RAOMEASE2 = 33,   /* rao_mease.jl:9-26; states (x, cost), p = (u) */
    CB200_MODEL_ROBBINS4 = 34,   /* robbins.jl:9-44; states (x1, x2, x3, cost), p = (u) */
    CB200_MODEL_MOONLANDING3 = 35,   /* moon_landing.jl:9-31; states (h, v, m), p = (t_s, T) */
    CB200_MODEL_LOTKASHARED4 = 36,   /* lotka_shared.jl:9-32; states (x0, x1, x2, cost), p = (u) */
    CB200_MODEL_HANGINGCHAIN3 = 37,   /* hanging_chain.jl:9-48; states (x1, x2, x3), p = (u) */
    CB200_MODEL_TUBULARREACTOR2 = 38,   /* tubular_reactor.jl:9-22; states (x1, x2), p = (u) */
    CB200_MODEL_OCEAN4 = 39,   /* ocean.jl:9-64; states (S, R, tau, cost), p = (u1, u2) */
    CB200_MODEL_DUCTEDFAN7 = 40,   /* ducted_fan.jl:9-61; states (x1, v1, x2, v2, alpha, valpha, cost), p = (t_s, u1, u2) */
    CB200_MODEL_HANGGLIDER4 = 41,  /* hang_glider.jl:9-66; states (x, y, vx, vy), p = (T, cL) */
    /* the dense family x' = W x - x^3 + b u at the other sizes the FP64 tensor-path kernel is built for (model_data = W
       row-major nx x nx, then b[nx]) */
    CB200_MODEL_DENSE16 = 42,
    CB200_MODEL_DENSE24 = 43,
    CB200_MODEL_DENSE32 = 44
} cb200_model;

typedef enum {
    CB200_TSIT5 = 0,            /* fixed step h = (t1 - t0)/steps_per_piece on every piece (adaptive = false, dt = h) */
    CB200_RK4 = 1,              /* classical RK4, fixed step */
    CB200_TSIT5_ADAPTIVE = 2    /* the reference's default: adaptive Tsit5 with the problem's abstol / reltol, every piece a
                                   fresh `solve` (src/single_shooting.jl:443-453); OrdinaryDiffEq's controller restated --
                                   reproduces the reference's golden vectors to ~1e-15.  steps_per_piece is ignored.
                                   DENSE20 runs this mode on the generic kernel (thread per direction), not on the tensor path */
} cb200_method;

/*
 * Static description of one layer == the reference's `st` (one entry per shooting interval; a
 * SingleShootingLayer is n_intervals == 1).  Per-interval arrays are concatenated in interval order.
 */
typedef struct {
    int32_t struct_size;        /* sizeof(cb200_desc), for ABI checking */
    int32_t model;              /* cb200_model */
    int32_t method;             /* cb200_method: alg of the layer (Tsit5() / RK4()) */
    int32_t steps_per_piece;    /* K: fixed steps per control piece, h = (t1-t0)/K  (adaptive=false, dt=h) */
    int32_t nx;                 /* length(problem.u0) */
    int32_t np_all;             /* length(problem.p), controls included */
    int32_t n_controls;         /* active controls = rows of index_grid */
    int32_t n_tunable_p;        /* length(ps.p) */
    int32_t n_intervals;        /* I */
    int32_t n_quadrature;       /* length(quadrature_indices) */
    int32_t device;             /* CUDA ordinal */
    int32_t fim_offset;         /* 1-based state index of the first F_triu state, 0 = no Fisher block */
    int32_t fim_n;              /* n: F is n x n, stored as n(n+1)/2 column-major upper-triangle states */
    int32_t fim_mode;           /* cb200_fim_mode: how the Fisher information is read out (0 = from the end state) */
    const double *u0;                   /* [nx] problem.u0 (fixval of the first interval, single_shooting.jl:359) */
    const int64_t *tunable_p;           /* [n_tunable_p] 1-based indices into problem.p (single_shooting.jl:109) */
    const int64_t *control_indices;     /* [n_controls] 1-based p index per control, user order (layer.control_indices) */
    const int64_t *quadrature_indices;  /* [n_quadrature] 1-based */
    /* per interval */
    const int32_t *n_pieces;            /* [I] M_i = length(st.tspans) (chunks flattened) */
    const double *tspans;               /* [sum M_i][2] st.tspans as (t0,t1) pairs */
    const int64_t *index_grid;          /* [sum n_controls*M_i] st.index_grid, column-major n_controls x M_i, 1-based */
    const int32_t *n_u0;                /* [I] length(ps.interval_i.u0) */
    const int32_t *n_local_controls;    /* [I] length(ps.interval_i.controls) */
    const int32_t *is_shooting;         /* [I] shooting_layer flag (fixval = zeros) */
    const int64_t *ic_sorting;          /* [I*nx] InitialConditionRemaker.sorting, 1-based */
    const int32_t *ic_n_constants;      /* [I] */
    const int64_t *ic_constants;        /* [sum ic_n_constants] InitialConditionRemaker.constants, 1-based */
    const int32_t *n_shooting_indices;  /* [I] */
    const int64_t *shooting_indices;    /* [sum] st.shooting_indices, 1-based over [states; controls] */
    const double *model_data;           /* model constants (DENSEnn: W row-major nx x nx, then b[nx]) or NULL.  Live handles of one
                                           dense model on one device must carry the same data (one __constant__ block) */
    int64_t model_data_len;
    /* sample-based Fisher read-outs (fim_mode != 0); G = the nx_base x fim_n block behind the base states, the
       observed functions are the leading n_observed base states */
    int32_t fim_base_nx;                /* states of the un-augmented problem */
    int32_t n_observed;                 /* n_observed(oed) = number of observed functions (oed.jl:283) */
    const int32_t *n_observation_rows;  /* [I] length(st.observation_grid.grid) of each interval (oed.jl:307-323,
                                           350-366); NULL when fim_mode == CB200_FIM_DISCRETE */
    const int64_t *observation_grid;    /* [sum rows][n_observed] WeightedObservation.grid, row-major as Julia's
                                           Vector{Vector{Int64}}: 1-based index into the interval's ps.controls of
                                           the weight applied at the row's sample, 0 = not measured (oed.jl:2-16).
                                           Row j of interval i belongs to the interval's j-th saved sample.  For
                                           CB200_FIM_FIXED_CONTINUOUS row j holds the j-th entries of
                                           st.active_controls (oed.jl:352-362) and weights F(sample j+1) - F(sample j) */
    double abstol, reltol;              /* CB200_TSIT5_ADAPTIVE: problem.kwargs abstol / reltol (OrdinaryDiffEq defaults
                                           1e-6 / 1e-3); ignored by the fixed-step methods */
    /* ---- version 201 */
    const double *saveat;               /* [n_saveat] problem.kwargs saveat, sorted ascending, or NULL.  The discrete OED layers
                                           with measurements put their measurement times here (lib/CorleoneOED/src/oed.jl:106-113):
                                           every control piece -- a fresh `solve(...; save_start, save_end = true)`,
                                           src/single_shooting.jl:443-453 -- additionally saves the state at the saveat times
                                           STRICTLY inside it, taken from Tsit5's dense output on the step that contains them
                                           (OrdinaryDiffEq's 4th-order interpolant; steps are NOT cut at these times).  The
                                           samples of `traj` / `traj_sens` and the rows of observation_grid then run over
                                           piece starts, saveat times and piece ends in time order (cb200_sample_times).
                                           Tsit5 methods only (fixed step and adaptive) */
    int32_t n_saveat;
    int32_t reserved_desc;
} cb200_desc;

/* Fisher read-out (lib/CorleoneOED/src/oed.jl:381-461) */
typedef enum {
    CB200_FIM_END = 0,              /* continuous: F_init + Symmetric(F_triu(t_end))                       (:388-398) */
    CB200_FIM_DISCRETE = 1,         /* OEDLayer{true,false}: sum over all saved samples of (sum_i h_i G)'(sum_i h_i G) (:446-461) */
    CB200_FIM_DISCRETE_SAMPLED = 2, /* OEDLayer{true,true}: sum_s sum_i (w_i h_i G)'(w_i h_i G)            (:401-432) */
    CB200_FIM_FIXED_CONTINUOUS = 3  /* OEDLayer{false,true,true}: sum_j sum_i w_i[j] (F_i(t_j+1) - F_i(t_j)) (:352-362,464-470) */
} cb200_fim_mode;

/* what cb200_eval should produce (bit mask) */
#define CB200_WANT_XEND 0x001u      /* raw end state of every interval */
#define CB200_WANT_SENS 0x002u      /* d x_end(i) / d (u0_i, p_i, controls_i) */
#define CB200_WANT_TRAJ 0x004u      /* merged Trajectory samples, quadratures accumulated */
#define CB200_WANT_DEFECTS 0x008u   /* shooting constraints, both orderings */
#define CB200_WANT_OBJ 0x010u       /* last(traj)[objective_row] per candidate */
#define CB200_WANT_GRAD 0x020u      /* d objective / d ps */
#define CB200_WANT_FIM 0x040u       /* F_init + the layer's Fisher read-out (cb200_fim_mode) per candidate, and its sum over candidates */
#define CB200_WANT_CRIT 0x100u      /* the six OED criteria of Symmetric(F) per candidate (criteria.jl:33-63) */
#define CB200_WANT_GRAD_COMPACT 0x200u /* d objective / d ps without the structural zeros (cb200_grad_structure) */
#define CB200_WANT_TRAJ_SENS 0x400u /* sensitivities at every saved sample (point constraints, src/dynprob.jl:30-58) */
#define CB200_WANT_JAC 0x080u       /* values of d(shooting_constraints)/d(ps), cb200_jac_structure order (cons_j) */

/* memory / layout flags */
#define CB200_MEM_DEVICE 0x1u       /* ps and all outputs are device pointers (default: host pointers) */
#define CB200_PS_SOA 0x2u           /* ps is variable-major: ps[v*ldb + b]   (default candidate-major
                                       ps[b*ldb + v], i.e. a Julia nvars x B Matrix with leading dim ldb) */
#define CB200_OUT_SOA 0x4u          /* outputs are quantity-major out[q*B + b] (default candidate-major
                                       out[b*n + q], i.e. Julia n x B matrices) */
/* multi-GPU flags (the handle has a communicator, cb200_comm_init*).  An evaluation with either flag is COLLECTIVE: every
 * rank issues the same sequence of such calls (B may differ from rank to rank) */
#define CB200_COMM_REDUCE 0x8u      /* objective_sum, grad_sum and fim_sum are all-reduced over the ranks: every rank
                                       receives the sums over the candidates of ALL ranks (bit-identical everywhere) */
#define CB200_COMM_GATHER 0x10u     /* the kind-major shooting defects of all ranks are all-gathered into every rank's
                                       window: cb200_out.defects_gathered / cb200_comm_gathered (needs WANT_DEFECTS,
                                       cb200_out.comm_B_total and comm_b0) */

/*
 * Output buffers; NULL = not wanted.  Sizes in doubles, n = the leading count, per candidate:
 *   xend           n = nx*I            [k + nx*i]
 *   sens           n = cb200_sizes.n_sens   rows ordered interval, then direction d (order u0,p,controls
 *                                           of the interval = its block of the flat ps), then state k
 *   traj           n = nrow*nsamples   [r + nrow*sample], nrow = nx + n_controls (single_shooting.jl:456)
 *   defects_kind   n = n_defects       shooting_constraints order  (multiple_shooting.jl:281-295)
 *   defects_stage  n = n_defects       stage_ordered_shooting_constraints order (:229,249-252)
 *   objective      n = 1
 *   grad           n = nvars
 *   grad_compact   n = cb200_grad_structure(...)   the gradient entries of the variables the objective depends on
 *   fim            n = fim_n*fim_n     column-major symmetric matrix
 *   criteria       n = 6               A, D, E, FisherA, FisherD, FisherE of F = F_init + F(t_end)
 *   traj_sens      n = n_traj_sens     d (state k at the interval's j-th sample) / d (variable d of the SAME interval):
 *                                       rows ordered interval, sample j (0 = the interval's start; the end sample of a
 *                                       non-final interval is dropped as in `traj`), direction d, state k.  Raw
 *                                       per-interval values: the running sums of quadrature rows are not added
 *   jac_vals       n = jac_nnz_var     sensitivity-valued nonzeros of the constraint Jacobian = the leading
 *                                       triplets of cb200_jac_structure (the constant +-1 entries follow them)
 *   fim_sum        fim_n*fim_n doubles total: sum over the B candidates of `fim`
 *   objective_sum  1 double total: sum over the B candidates of `objective`   (expectation-type objectives;
 *   grad_sum       nvars doubles total: sum over the B candidates of `grad`    the per-rank partial sums that
 *                                                                              are all-reduced across GPUs)
 */
typedef struct {
    double *xend;
    double *sens;
    double *traj;
    double *defects_kind;
    double *defects_stage;
    double *objective;
    double *grad;
    double *fim;
    double *fim_sum;
    double *objective_sum;
    double *grad_sum;
    const double *fim_init;     /* [fim_n*fim_n] st.F_init or NULL = zeros (oed.jl:309-315) */
    int32_t objective_row;      /* 1-based row of a trajectory sample: state k or nx + control r */
    int32_t reserved;
    double *jac_vals;
    double *criteria;
    double *grad_compact;
    double *traj_sens;
    int32_t *status;            /* [B] failure flags per candidate (always filled when non-NULL): bit 0 = a non-finite
                                   end state on some interval (in the adaptive mode also: more than 1e5 steps in a piece or a
                                   step below resolution -- OrdinaryDiffEq's MaxIters / DtLessThanMin / Unstable).  The reference
                                   never inspects the solver's retcode. */
    /* ---- version 200 */
    double *defects_gathered;   /* CB200_COMM_GATHER: n = n_defects, the candidates of ALL ranks in global order:
                                   [q * comm_B_total + b] (CB200_OUT_SOA) or [b * n_defects + q]; NULL = leave the result in
                                   the handle's window (cb200_comm_gathered) */
    int64_t comm_B_total;       /* CB200_COMM_GATHER: candidates of all ranks together */
    int64_t comm_b0;            /* CB200_COMM_GATHER: global index of this rank's first candidate */
    uint32_t resident;          /* want-bits (CB200_WANT_*) of results that stay on the device: their buffer pointers above are
                                   ignored, the result is kept in handle-owned memory, quantity-major [q * B + b], until the
                                   next evaluation (cb200_resident) -- e.g. the constraint-Jacobian values of 65 536
                                   candidates never cross PCIe when the consumer factorises them on the device */
    uint32_t reserved2;
} cb200_out;

typedef struct {
    int64_t nvars;          /* LuxCore.parameterlength(layer) */
    int64_t n_defects;      /* get_number_of_shooting_constraints */
    int64_t n_samples;      /* samples of the merged trajectory */
    int64_t n_row;          /* nx + n_controls */
    int64_t n_sens;         /* sum_i nx * ns_i */
    int64_t n_units;        /* I: (interval) units per candidate */
    int64_t jac_nnz;        /* nonzeros of d(shooting_constraints)/d(ps) */
    int64_t jac_nnz_var;    /* the leading jac_nnz_var triplets carry sensitivity values (jac_vals); the rest are +-1 */
    int64_t n_traj_sens;    /* sum_i samples_i * ns_i * nx */
} cb200_sizes;

typedef struct cb200_handle cb200_handle;

int cb200_version(void);
const char *cb200_last_error(void);

int cb200_create(const cb200_desc *desc, cb200_handle **out);
int cb200_destroy(cb200_handle *h);
int cb200_get_sizes(const cb200_handle *h, cb200_sizes *sizes);

/* times of the n_samples samples of the merged trajectory (`traj`): piece starts / saveat times / piece ends in order, the
 * duplicated boundary point of all but the last interval dropped (src/multiple_shooting.jl:145-147,178-180) */
int cb200_sample_times(const cb200_handle *h, double *t /* n_samples */);

/* block offsets of the flat variable vector, I+1 entries, 0-based offsets as get_block_structure returns */
int cb200_block_structure(const cb200_handle *h, int64_t *blocks);

/* optional: run on a caller-provided cudaStream_t (default: a private non-blocking stream). */
int cb200_set_stream(cb200_handle *h, void *cuda_stream);

/* Interval sharding ("B < #GPUs"): restrict the following cb200_eval calls to the shooting
 * intervals first:last (1-based, inclusive, as a Julia range; first:first-1 = none, 1:I = all = the default) -- the
 * slice of `mythreadmap` over intervals (src/Corleone.jl:22-24, src/multiple_shooting.jl:118-130) one device takes.
 * Every per-candidate output keeps its full size and order; what the other intervals would have written is ZERO
 * (matching defects belong to the interval that ends at the node), so the results of devices holding disjoint
 * ranges assemble by summation: one all-reduce(SUM).  Not available with a partial range: the Fisher read-outs and
 * trajectory samples with quadrature running sums (both need every interval's end state). */
int cb200_set_interval_range(cb200_handle *h, int32_t first, int32_t last);

/* Evaluate B candidates.  ldb: leading dimension of ps (>= nvars candidate-major, >= B variable-major).
 * Returns after the results are in the caller's buffers (host pointers) or after the work has been
 * enqueued on the handle's stream (device pointers). */
int cb200_eval(cb200_handle *h, const double *ps, int64_t B, int64_t ldb, uint32_t want, uint32_t flags,
               const cb200_out *out);

/* Variables the objective row depends on: the contiguous range [*first, *first + n) (1-based) of the flat variable
 * vector -- everything for a quadrature row, the last interval's block for another state row; returns n (< 0: error).
 * `grad_compact` holds exactly these entries: a dense gradient of a 100-interval layer is 99 % structural zeros. */
int64_t cb200_grad_structure(const cb200_handle *h, int32_t objective_row, int64_t *first);

/* forward_initialization(rng, shooting; ps, fixed_indices) (src/node_initialization.jl:152-170) for B candidates at
 * once: walks the intervals in order and overwrites each interval's node values ps.interval_i.u0 in place with the
 * previous interval's end state; intervals listed in fixed_indices (1-based) keep their nodes.  flags: CB200_MEM_DEVICE
 * and CB200_PS_SOA as for cb200_eval. */
int cb200_forward_init(cb200_handle *h, double *ps, int64_t B, int64_t ldb, uint32_t flags,
                       const int32_t *fixed_indices, int32_t n_fixed);

/* Sparsity of d(shooting_constraints kind-major)/d(ps): nnz triplets (1-based row, col) and, per entry,
 * src >= 0: 0-based row into `sens`;  src == -1: constant +1;  src == -2: constant -1.  The jac_nnz_var entries
 * with src >= 0 come first, in the order of `jac_vals`. */
int cb200_jac_structure(const cb200_handle *h, int64_t *rows, int64_t *cols, int64_t *src);

/* Device-resident results of the last cb200_eval (cb200_out.resident): *ptr = device pointer of the result selected by ONE
 * CB200_WANT_* bit (for CB200_WANT_DEFECTS: defects_kind), laid out [q * (*ld) + b]; valid until the next evaluation on
 * the handle.  Work enqueued on the handle's stream (cb200_set_stream) is ordered after the evaluation. */
int cb200_resident(const cb200_handle *h, uint32_t want_bit, double **ptr, int64_t *ld);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU data plane: one rank (process or thread) per GPU of one box.  Replaces mythreadmap(::EnsembleDistributed) =
 * pmap over intervals (src/Corleone.jl:24, src/multiple_shooting.jl:118-130), whose per-interval results travel to the
 * master over TCP: here every rank evaluates a contiguous slice of the candidate axis and only the cross-candidate sums
 * (objective, gradient, Fisher information: all-reduce) and the shooting defects (all-gather) cross NVLink, written by
 * the evaluation's own kernels into peer-mapped windows (no collective library call per evaluation; NCCL is the
 * fallback transport and one of two ways to bootstrap).
 * ------------------------------------------------------------------------------------------------------------------ */
#define CB200_COMM_ID_BYTES 128
#define CB200_TRANSPORT_NONE 0
#define CB200_TRANSPORT_NCCL 1   /* ncclAllReduce / ncclAllGather on the handle's stream after the kernels */
#define CB200_TRANSPORT_P2P 2    /* fused: P2P stores + one-shot all-reduce in the tail of the evaluation's last kernel */

/* all-gather provided by the host: every rank contributes `bytes` at `send`, `recv` receives nranks * bytes in rank
 * order; returns 0.  (Julia: MPI.Allgather!, or a Distributed.jl remotecall_fetch round; tests: a threading.Barrier.) */
typedef int (*cb200_allgather_fn)(void *ctx, const void *send, int64_t bytes, void *recv);

typedef struct {
    int32_t nranks, rank;
    int32_t transport;      /* CB200_TRANSPORT_* */
    int32_t reserved;
    int64_t epoch;          /* collective evaluations so far */
    int64_t max_B_total;    /* capacity of the gather window in candidates */
    int64_t error;          /* != 0: a peer timed out (also reported by the next cb200_eval as CB200_ERR_COMM) */
} cb200_comm_info;

/* rank 0 creates the 128-byte id (ncclGetUniqueId) and the host hands it to the other ranks */
int cb200_comm_unique_id(void *id /* CB200_COMM_ID_BYTES */);
/* Collective over the nranks handles of one layer (same descriptor on every rank, each on its own device): creates the
 * communicator (ncclCommInitRank), the handle's windows and the peer mappings.  max_B_total: the largest
 * cb200_out.comm_B_total a CB200_COMM_GATHER evaluation will use (0 = no gather window). */
int cb200_comm_init(cb200_handle *h, int32_t nranks, int32_t rank, const void *id, int64_t max_B_total);
/* the same without NCCL: the host provides the bootstrap all-gather (P2P transport only; the ranks may be threads of one
 * process, then the windows are shared by plain peer access) */
int cb200_comm_init_host(cb200_handle *h, int32_t nranks, int32_t rank, cb200_allgather_fn allgather, void *ctx,
                         int64_t max_B_total);
int cb200_comm_destroy(cb200_handle *h);
int cb200_comm_query(const cb200_handle *h, cb200_comm_info *info);
/* the all-gathered defects of the last CB200_COMM_GATHER evaluation, in the handle's window: [q * (*ld) + b], *ld =
 * comm_B_total.  Valid until the second-next collective evaluation (the window is double-buffered). */
int cb200_comm_gathered(const cb200_handle *h, double **ptr, int64_t *ld);

/* layout self-check for bindings that restate the structs by hand (julia/CorleoneB200.jl): sizeof and the offsets of
 * the last fields; which: 0 = cb200_desc, 1 = cb200_out, 2 = cb200_sizes, 3 = cb200_comm_info.  Returns the number
 * of values written to `vals` (sizeof first), at most n. */
int cb200_abi_layout(int32_t which, int64_t *vals, int32_t n);

/* Device time of the last cb200_eval's integration kernel(s) in milliseconds (CUDA events on the
 * handle's stream); < 0 if none. */
float cb200_last_kernel_ms(const cb200_handle *h);
/* Device times of the integration kernel of the last n cb200_eval calls (newest first, at most 256 kept);
 * returns how many were written.  Synchronises on the recorded events only. */
int cb200_kernel_ms_history(const cb200_handle *h, float *ms, int n);
/* kernels launched by this handle since creation */
int64_t cb200_launch_count(const cb200_handle *h);

/* FP64 FMA-chain peak probe: runs `iters` dependent DFMA chains on every SM, returns TFLOP/s (<0 on error) */
double cb200_probe_fp64_tflops(int device, int iters);
/* same through the FP64 tensor path (mma.sync.m8n8k4.f64); on B200 both share one datapath, the roof is the max */
double cb200_probe_fp64_mma_tflops(int device, int iters);

#ifdef __cplusplus
}
#endif
#endif
