/*
 * corleone_oracle.h -- CPU restatement of Corleone.jl's shooting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker / CPU baseline.
 *
 * Parity status: PINNED on the reference's own golden vectors.  The real
 * reference (Julia + un-vendored OrdinaryDiffEq / ForwardDiff) cannot run in
 * the build container, but its tests hold golden values produced by its
 * default mode (adaptive Tsit5 per control piece).  With OrdinaryDiffEq's
 * step-size control restated (CO_TSIT5_ADAPTIVE, corleone_oracle.c) this
 * oracle reproduces them to ~1e-15: G1, G2, G3, G4 at abstol 1e-8 / reltol
 * 1e-6 and the OED A-criteria G7 at the default 1e-6 / 1e-3
 * (tests/test_oracle_adaptive.py) -- which pins the right-hand sides, the
 * tableau, the piece loop and save semantics, the merge with defects and
 * quadrature sums, the forward initialisation, the OED augmentation and the
 * Fisher read-out in one go.  The fixed-step modes (CO_TSIT5, CO_RK4:
 * `adaptive = false, dt = h`) share every one of those pieces and differ only
 * in taking h as given; no reference test runs them, so they are checked
 * against the goldens at the accuracy a converged fixed step reaches
 * (<= 1e-10 for G1/G4, 1e-9..1e-7 for G3/G7, tests/test_oracle_golden.py),
 * against closed forms (LQR, 1-D OED) and a 40-digit restatement.
 *
 * All reference citations are relative to /root/reference.
 */
#ifndef CORLEONE_ORACLE_H
#define CORLEONE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* forward-mode tangent lanes carried per pass (ForwardDiff-style chunk) */
#define CO_ND 16
#define CO_MAX_NX 32
#define CO_MAX_NP 16
#define CO_MAX_CTRL 8

enum co_model {
    CO_LOTKA3 = 0,      /* Lotka fishing + Lagrange state, test/core/multiple_shooting.jl:11-17 */
    CO_LQR = 1,         /* examples/linear_quadratic/main.jl:45-50 */
    CO_LOTKA2 = 2,      /* lib/CorleoneOED/test/core/lotka_oed.jl:15-20 */
    CO_LOTKA2_OED = 3,  /* [x; G; F_triu] augmentation of CO_LOTKA2, params=[2,3], observed=u[1:2] */
    CO_LIN1D = 4,       /* lib/CorleoneOED/test/core/1d_oed.jl:11-13 */
    CO_LIN1D_OED = 5,   /* augmentation of CO_LIN1D, params=[1], observed=[u[1]] */
    CO_DENSE20 = 6,     /* synthetic BASELINE config 4: x' = W x - x^3 + b u */
    CO_EGERSTEDT = 7,   /* test/core/local_controls.jl:46-52 */
    CO_VDP3 = 8,        /* lib/OptimalControlBenchmarks/src/problems/van_der_pol.jl:16-33 (+ Lagrange state) */
    CO_CARTPOLE5 = 9,   /* .../cart_pendulum.jl:17-48 (+ Lagrange state) */
    CO_ROCKET3 = 10,    /* .../goddarts_rocket.jl:16-40, p = (T, u) */
    CO_QUADROTOR7 = 11, /* .../quadrotor.jl:9-52 (+ Lagrange state), p = (w1, w2, w3, u) */
    CO_THREETANK4 = 12, /* .../three_tank.jl:9-53 (+ Lagrange state), p = (w1, w2, w3) */
    CO_BATCHREACTOR2 = 13, /* .../batch_reactor.jl:9-17, p = (u) */
    CO_LOTKACOMP3 = 14,  /* .../lotka_competitive.jl:9-31 (+ Lagrange state) */
    CO_F8AIRCRAFT4 = 15, /* .../f8.jl:9-41, p = (T, u) */
    CO_DOUBLEOSC6 = 16,  /* .../double_oscillator.jl:9-38 (+ time and Lagrange states) */
    /* the other variants of lib/CorleoneOED/src/augmentation.jl for the Lotka / 1-D problems */
    CO_LOTKA2_OED_CU = 17, /* Continuous, no measurements (:217-220) */
    CO_LOTKA2_OED_CF = 18, /* ContinuousSampled, fixed (:255-292) */
    CO_LOTKA2_OED_DS = 19, /* DiscreteSampled (:171-203) */
    CO_LOTKA2_OED_DU = 20, /* Discrete (:171-203) */
    CO_LIN1D_OED_CF = 21,
    CO_LIN1D_OED_DS = 22,
    /* the rest of lib/OptimalControlBenchmarks/src/problems */
    CO_BRYSONDENHAM3 = 23,
    CO_CATALYSTMIXING2 = 24,
    CO_CUSHIONED3 = 25,
    CO_DENBIGH3 = 26,
    CO_DIELECTROPHORETIC3 = 27,
    CO_ELECTRICCAR4 = 28,
    CO_FULLER3 = 29,
    CO_JACKSON3 = 30,
    CO_MOUNTAINCAR3 = 31,
    CO_PARTICLESTEERING5 = 32,
    CO_RAOMEASE2 = 33,
    CO_ROBBINS4 = 34,
    CO_MOONLANDING3 = 35,
    CO_LOTKASHARED4 = 36,
    CO_HANGINGCHAIN3 = 37,
    CO_TUBULARREACTOR2 = 38,
    CO_OCEAN4 = 39,
    CO_DUCTEDFAN7 = 40,
    CO_HANGGLIDER4 = 41,
    CO_DENSE16 = 42, /* the dense family at the other sizes of the tensor-path kernel */
    CO_DENSE24 = 43,
    CO_DENSE32 = 44
};

enum co_method { CO_TSIT5 = 0, CO_RK4 = 1, CO_TSIT5_ADAPTIVE = 2 /* the reference's default mode; K is ignored */ };
/* abstol / reltol of CO_TSIT5_ADAPTIVE (process-wide; defaults 1e-6 / 1e-3 as OrdinaryDiffEq's) */
void co_set_tolerances(double abstol, double reltol);

typedef struct {
    int model;
    int nx, np_all;                 /* states, length of problem.p (controls included) */
    const double *u0;               /* nx */
    const double *p0;               /* np_all */
    double tspan[2];
    int ncontrols;                  /* as the user passed them (order kept) */
    const int *control_indices;     /* 1-based index into problem.p, per control */
    const double *const *ctrl_t;    /* per control: time grid */
    const double *const *ctrl_u;    /* per control: initial values (same length) or NULL -> zeros */
    const int *ctrl_nt;
    int n_tunable_ic;
    const int *tunable_ic;          /* 1-based */
    int nquad;
    const int *quad_idx;            /* 1-based quadrature_indices */
    int n_shoot;                    /* 0 -> single shooting */
    const double *shoot_pts;        /* tpoints of MultipleShootingLayer */
    const double *model_data;       /* CO_DENSE20: W (20x20 row-major) then b (20) */
    int n_saveat;                   /* problem.kwargs saveat (lib/CorleoneOED/src/oed.jl:106-113): sorted times; every piece */
    const double *saveat;           /* also saves at those strictly inside it, from Tsit5's dense output */
} co_problem;

typedef struct co_layer co_layer;

/* ---- index logic, callable on its own (src/local_controls.jl) ---- */
/* get_timegrid (:53-57): indices k with t0 <= t[k] < t1; returns count, writes 0-based positions */
int co_restrict_grid(const double *t, int nt, double t0, double t1, int *pos);
/* build_index_grid (:149-178) + collect_tspans (:183-195), unchunked.  idx is column-major
 * ncontrols x M, 1-based exactly as Julia returns it; grid has M+1 points.  Returns M. */
int co_build_index_grid(int ncontrols, const double *const *t, const int *nt, double t0, double t1,
                        int64_t *idx, double *grid, int cap_pieces);
/* get_subvector_indices (:120-147): writes 1-based (start,stop) pairs, returns chunk count */
int co_subvector_indices(int M, int L, int *start_stop);

/* Tsit5's interpolation weights b_1..b_7 at theta (OrdinaryDiffEq's dense output for Tsit5) */
void co_tsit5_interp_weights(double theta, double *bw /* 7 */);

/* ---- layer = SingleShootingLayer or MultipleShootingLayer + its `st` ---- */
co_layer *co_layer_create(const co_problem *prob);
void co_layer_destroy(co_layer *L);
int co_n_intervals(const co_layer *L);
int co_nvars(const co_layer *L);                    /* LuxCore.parameterlength */
int co_nrow(const co_layer *L);                     /* nx + active controls: rows of a trajectory sample */
void co_block_structure(const co_layer *L, int *blocks /* I+1 */);
void co_interval_info(const co_layer *L, int i, double *tspan2, int *M, int *n_u0, int *n_p, int *n_c,
                      int *n_shooting_idx);
void co_interval_tables(const co_layer *L, int i, int64_t *idx /* nact x M */, double *grid /* M+1 */,
                        int *shooting_indices /* 1-based */, int *ic_sorting, int *ic_constants,
                        int *n_constants);
void co_default_ps(const co_layer *L, double *ps);  /* default_initialization, flat (u0,p,controls) per interval */
int co_n_defects(const co_layer *L);                /* get_number_of_shooting_constraints */
int co_n_samples(const co_layer *L);                /* samples of the merged Trajectory */

/* ---- evaluation with up to CO_ND tangent directions ----
 * ps: nvars values; dps: nvars x nd column-major tangents (may be NULL when nd == 0).
 * K fixed RK steps per control piece, h = (t1 - t0)/K.
 * Outputs (any may be NULL); tangent outputs have nd trailing columns (column-major):
 *   t_out   [nsamples], u_out [nrow x nsamples] column-major (merged Trajectory, quadratures accumulated)
 *   xend    [nx x I] raw end state of every interval (before quadrature accumulation)
 *   defk    [ndef] kind-major shooting_constraints (multiple_shooting.jl:281-295)
 *   defs    [ndef] stage-major stage_ordered_shooting_constraints (:229,249-252)
 *   last    [nrow] last sample of the merged trajectory (objective = one row of it, dynprob.jl:70-71)
 */
int co_eval(const co_layer *L, const double *ps, const double *dps, int nd, int method, int K,
            double *t_out, double *u_out, double *xend, double *dxend, double *defk, double *ddefk,
            double *defs, double *ddefs, double *last, double *dlast);

/* tangents of the merged trajectory samples: du_out[r + nrow*(sample + nsamples*d)] (column-major, nd trailing) */
int co_eval_traj(const co_layer *L, const double *ps, const double *dps, int nd, int method, int K, double *u_out,
                 double *du_out);

/* forward_initialization (src/node_initialization.jl:152-170): overwrites the u0 entries of ps in place */
int co_forward_init(const co_layer *L, double *ps, const int *fixed_indices, int nfixed, int method, int K);

/* ---- CPU baseline: every (interval, candidate) unit with its own local sensitivities ----
 * ps_batch: nvars x B column-major (candidate per column).  Integrates each unit with identity seeds on
 * its (u0,p,controls) block, in ceil(ns_i/CO_ND) passes, OpenMP over units (EnsembleThreads analogue,
 * src/Corleone.jl:23).  Returns seconds; checksum guards against dead-code elimination.
 * sens_out (optional): per unit nx*ns_i doubles packed, interval-major then candidate, row-major [dir][state]. */
double co_bench_units(const co_layer *L, const double *ps_batch, int64_t B, int64_t b_begin, int64_t b_end,
                      int method, int K, int with_sens, int nthreads, double *checksum);

int co_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
