/*
 * corleone_oracle.c -- CPU restatement of Corleone.jl's shooting hot path (see corleone_oracle.h).
 * TEST INFRASTRUCTURE ONLY: checker and CPU baseline, never the product path.
 *
 * What is restated, and from where (paths relative to /root/reference):
 *   control grid restriction / index grid / tspans   src/local_controls.jl:53-57,120-195
 *   InitialConditionRemaker, parameter scatter       src/single_shooting.jl:255-270,306-323,357-362
 *   piece loop and save semantics                    src/single_shooting.jl:421-471
 *   interval split, merge, defects, quadratures      src/multiple_shooting.jl:86-95,139-182,229-310
 *   default / forward node initialisation            src/multiple_shooting.jl:45-55, src/node_initialization.jl:152-170
 *   OED augmentation [x; G; F_triu]                  lib/CorleoneOED/src/augmentation.jl:131-159,206-253
 * Third-party arithmetic the reference calls but does not contain:
 *   OrdinaryDiffEqTsit5 (Project.toml:44, compat "1, 2", unpinned -- no Manifest): explicit 7-stage FSAL
 *   Tsitouras 5(4) scheme, published tableau (Tsitouras 2011).  CO_TSIT5 / CO_RK4 run it with adaptive=false (only the
 *   propagation weights); CO_TSIT5_ADAPTIVE restates OrdinaryDiffEq's step-size control as well (see
 *   integrate_piece_adaptive) and reproduces the reference's golden vectors to ~1e-15.
 *   ForwardDiff (dual numbers through `solve`): restated as `dual` below.
 */
#include "corleone_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <time.h>

/* ------------------------------------------------------------------ duals */
typedef struct {
    double v;
    double d[CO_ND];
} dual;

static _Thread_local int g_nd = 0;

static inline dual dconst(double c)
{
    dual r;
    r.v = c;
    for (int k = 0; k < g_nd; k++) r.d[k] = 0.0;
    return r;
}
static inline dual dadd(dual a, dual b)
{
    dual r;
    r.v = a.v + b.v;
    for (int k = 0; k < g_nd; k++) r.d[k] = a.d[k] + b.d[k];
    return r;
}
static inline dual dsub(dual a, dual b)
{
    dual r;
    r.v = a.v - b.v;
    for (int k = 0; k < g_nd; k++) r.d[k] = a.d[k] - b.d[k];
    return r;
}
static inline dual dmul(dual a, dual b)
{
    dual r;
    r.v = a.v * b.v;
    for (int k = 0; k < g_nd; k++) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
}
static inline dual dscale(dual a, double c)
{
    dual r;
    r.v = a.v * c;
    for (int k = 0; k < g_nd; k++) r.d[k] = a.d[k] * c;
    return r;
}
static inline dual daddc(dual a, double c)
{
    a.v += c;
    return a;
}
static inline dual dneg(dual a) { return dscale(a, -1.0); }
static inline dual ddiv(dual a, dual b)
{
    dual r;
    r.v = a.v / b.v;
    for (int k = 0; k < g_nd; k++) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v;
    return r;
}
static inline dual dsin(dual a)
{
    dual r;
    const double c = cos(a.v);
    r.v = sin(a.v);
    for (int k = 0; k < g_nd; k++) r.d[k] = c * a.d[k];
    return r;
}
static inline dual dcos(dual a)
{
    dual r;
    const double s = sin(a.v);
    r.v = cos(a.v);
    for (int k = 0; k < g_nd; k++) r.d[k] = -s * a.d[k];
    return r;
}
static inline dual dsqrt(dual a)
{
    dual r;
    r.v = sqrt(a.v);
    for (int k = 0; k < g_nd; k++) r.d[k] = 0.5 * a.d[k] / r.v;
    return r;
}
static inline dual dexp(dual a)
{
    dual r;
    r.v = exp(a.v);
    for (int k = 0; k < g_nd; k++) r.d[k] = r.v * a.d[k];
    return r;
}
/* r += c * a */
static inline void daxpy(dual *r, double c, const dual *a)
{
    r->v += c * a->v;
    for (int k = 0; k < g_nd; k++) r->d[k] += c * a->d[k];
}

/* ------------------------------------------------------------------ models */
typedef struct {
    int model;
    const double *data; /* dense20: W row-major 20x20, then b[20] */
} model_ctx;

/* Lotka fishing with Lagrange term; test/core/multiple_shooting.jl:11-17 */
static void rhs_lotka3(const dual *u, const dual *p, dual *du)
{
    dual u12 = dmul(u[0], u[1]);
    du[0] = dsub(dsub(u[0], dmul(p[1], u12)), dscale(dmul(p[0], u[0]), 0.4));
    du[1] = dsub(dadd(dneg(u[1]), dmul(p[2], u12)), dscale(dmul(p[0], u[1]), 0.2));
    dual a = daddc(u[0], -1.0), b = daddc(u[1], -1.0);
    du[2] = dadd(dmul(a, a), dmul(b, b));
}
/* lib/CorleoneOED/test/core/lotka_oed.jl:15-20 */
static void rhs_lotka2(const dual *u, const dual *p, dual *du)
{
    dual u12 = dmul(u[0], u[1]);
    du[0] = dsub(dsub(u[0], dmul(p[1], u12)), dscale(dmul(p[0], u[0]), 0.4));
    du[1] = dsub(dadd(dneg(u[1]), dmul(p[2], u12)), dscale(dmul(p[0], u[1]), 0.2));
}
/* Jacobians of lotka2: Jx row-major 2x2; Jp columns for params [2,3] (row-major 2x2) */
static void jac_lotka2(const dual *u, const dual *p, dual *Jx, dual *Jp)
{
    /* f1 = u1 - p2 u1 u2 - 0.4 p1 u1 ; f2 = -u2 + p3 u1 u2 - 0.2 p1 u2 */
    Jx[0] = dsub(dsub(dconst(1.0), dmul(p[1], u[1])), dscale(p[0], 0.4));
    Jx[1] = dneg(dmul(p[1], u[0]));
    Jx[2] = dmul(p[2], u[1]);
    Jx[3] = dsub(daddc(dmul(p[2], u[0]), -1.0), dscale(p[0], 0.2));
    dual u12 = dmul(u[0], u[1]);
    Jp[0] = dneg(u12);
    Jp[1] = dconst(0.0);
    Jp[2] = dconst(0.0);
    Jp[3] = u12;
}
/* examples/linear_quadratic/main.jl:45-50 ; p = (a, b, u) */
static void rhs_lqr(const dual *x, const dual *p, dual *du)
{
    du[0] = dadd(dmul(p[0], x[0]), dmul(p[1], p[2]));
    dual e = daddc(x[0], -3.0);
    du[1] = dadd(dscale(dmul(e, e), 10.0), dscale(dmul(p[2], p[2]), 0.1));
}
/* lib/CorleoneOED/test/core/1d_oed.jl:11-13 */
static void rhs_lin1d(const dual *u, const dual *p, dual *du) { du[0] = dmul(p[0], u[0]); }
static void jac_lin1d(const dual *u, const dual *p, dual *Jx, dual *Jp)
{
    Jx[0] = p[0];
    Jp[0] = u[0];
}
/* test/core/local_controls.jl:46-52 */
static void rhs_egerstedt(const dual *u, const dual *p, dual *du)
{
    dual x = u[0], y = u[1];
    dual xpy = dadd(x, y), xmy = dsub(x, y);
    du[0] = dadd(dadd(dneg(dmul(x, p[0])), dmul(xpy, p[1])), dmul(xmy, p[2]));
    dual x2y = dadd(x, dscale(y, 2.0)), xm2y = dsub(x, dscale(y, 2.0));
    du[1] = dadd(dadd(dmul(x2y, p[0]), dmul(xm2y, p[1])), dmul(xpy, p[2]));
    du[2] = dadd(dmul(x, x), dmul(y, y));
}
/* synthetic BASELINE config 4: x' = W x - x.^3 + b u ; p = (u); n = 20, and the same family at 16, 24, 32 */
static void rhs_dense(int n, const dual *x, const dual *p, dual *du, const double *data)
{
    const double *W = data, *b = data + n * n;
    for (int i = 0; i < n; i++) {
        dual acc = dscale(p[0], b[i]);
        for (int j = 0; j < n; j++) daxpy(&acc, W[i * n + j], &x[j]);
        dual x3 = dmul(dmul(x[i], x[i]), x[i]);
        du[i] = dsub(acc, x3);
    }
}

/*
 * OED augmentation, ContinuousSampled non-fixed variant (augmentation.jl:206-253 with measurements):
 *   state  = [x (nx); vec(G) column-major nx x nF; F[triu] column-major]           (:239, selector :211)
 *   params = [base p; w_1..w_nobs]                                                 (:228-230)
 *   G'     = f_x G + f_p[:, params]                                                (:147-154)
 *   F'     = sum_i w_i (h_x[i,:] G)' (h_x[i,:] G), upper triangle                  (:222-233)
 * observed = leading nobs states (u[1:2] for Lotka, [u[1]] for the 1-D problem), so h_x rows are unit rows.
 */
/* lib/OptimalControlBenchmarks/src/problems/van_der_pol.jl:16-33, Lagrange term as third state; p = (u) */
static void rhs_vdp3(const dual *x, const dual *p, dual *du)
{
    dual one_m = dsub(dconst(1.0), dmul(x[1], x[1]));
    du[0] = dadd(dsub(dmul(one_m, x[0]), x[1]), p[0]);
    du[1] = x[0];
    du[2] = dadd(dadd(dmul(x[0], x[0]), dmul(x[1], x[1])), dmul(p[0], p[0]));
}
/* cart_pendulum.jl:17-48: states (x, theta, dx, dtheta, cost), p = (w); the two denominators differ in the reference */
static void rhs_cartpole5(const dual *x, const dual *p, dual *du)
{
    const double al = 10.0, be = 50.0, ga = 0.5, M = 1.0, m = 0.1, g = 9.81, pi = 3.141592653589793;
    dual s = dsin(x[1]), c = dcos(x[1]);
    dual num = dadd(dadd(p[0], dscale(dmul(s, c), m * g)), dscale(dmul(dmul(x[3], x[3]), s), m));
    dual omc = dsub(dconst(1.0), c);
    du[0] = x[2];
    du[1] = x[3];
    du[2] = ddiv(num, daddc(dscale(dmul(omc, omc), m), M));
    dual den2 = daddc(dscale(dsub(dconst(1.0), dmul(c, c)), m), M);
    du[3] = dsub(dscale(s, -g), dmul(ddiv(num, den2), c));
    dual dth = daddc(x[1], -pi);
    du[4] = dscale(dadd(dadd(dscale(dmul(x[0], x[0]), al), dscale(dmul(dth, dth), be)), dscale(dmul(p[0], p[0]), ga)), 1.0e-3);
}
/* goddarts_rocket.jl:16-40: states (r, v, m), p = (T, u) */
static void rhs_rocket3(const dual *x, const dual *p, dual *du)
{
    const double b = 7.0, Tmax = 3.5, A = 310.0, k = 500.0, r0 = 1.0;
    dual drag = dmul(dscale(dmul(x[1], x[1]), A), dexp(dscale(daddc(x[0], -r0), -k)));
    du[0] = dmul(p[0], x[1]);
    dual grav = ddiv(dconst(-1.0), dmul(x[0], x[0]));
    dual thrust = dmul(ddiv(dconst(1.0), x[2]), dsub(dscale(p[1], Tmax), drag));
    du[1] = dmul(p[0], dadd(grav, thrust));
    du[2] = dmul(p[0], dscale(p[1], -b));
}

/* quadrotor.jl:9-52: states (x1..x6, cost), p = (w1, w2, w3, u) */
static void rhs_quadrotor7(const dual *x, const dual *p, dual *du)
{
    const double g = 9.81, M = 1.3, L = 0.305, I = 0.0605;
    dual s5 = dsin(x[4]), c5 = dcos(x[4]);
    dual wu = dmul(p[0], p[3]);
    du[0] = x[1];
    du[1] = dadd(dscale(s5, g), dscale(dmul(wu, s5), 1.0 / M));
    du[2] = x[3];
    du[3] = dadd(daddc(dscale(c5, g), -g), dscale(dmul(wu, c5), 1.0 / M));
    du[4] = x[5];
    du[5] = dmul(dscale(dsub(p[2], p[1]), L / I), p[3]);
    du[6] = dscale(dmul(p[3], p[3]), 5.0);
}
/* three_tank.jl:9-53: states (x1, x2, x3, cost), p = (w1, w2, w3) */
static void rhs_threetank4(const dual *x, const dual *p, dual *du)
{
    const double c1 = 1.0, c2 = 2.0, c3 = 0.8, k1 = 2.0, k2 = 3.0, k3 = 1.0, k4 = 3.0;
    dual r1 = dsqrt(x[0]), r2 = dsqrt(x[1]), r3 = dsqrt(x[2]);
    dual q = dmul(p[2], dsqrt(dscale(x[0], c3)));
    du[0] = dsub(dsub(dadd(dscale(p[0], c1), dscale(p[1], c2)), r2), q);
    du[1] = dsub(r1, r2);
    du[2] = dadd(dsub(r2, r3), q);
    dual e2 = daddc(x[1], -k2), e3 = daddc(x[2], -k4);
    du[3] = dadd(dscale(dmul(e2, e2), k1), dscale(dmul(e3, e3), k3));
}
/* batch_reactor.jl:9-17: states (x1, x2), p = (u) */
static void rhs_batchreactor2(const dual *x, const dual *p, dual *du)
{
    dual r1 = dscale(dmul(dexp(ddiv(dconst(-2500.0), p[0])), dmul(x[0], x[0])), 4.0e3);
    dual r2 = dscale(dmul(dexp(ddiv(dconst(-5000.0), p[0])), dmul(x[1], x[1])), 62.0e4);
    du[0] = dneg(r1);
    du[1] = dsub(r1, r2);
}

/* lotka_competitive.jl:9-31: states (x0, x1, cost), p = (u) */
static void rhs_lotkacomp3(const dual *x, const dual *p, dual *du)
{
    const double c1 = 0.1, c2 = 0.4, al = 1.2, K = 1.8;
    dual a = dsub(dconst(1.0), dscale(dadd(x[0], dscale(x[1], al)), 1.0 / K));
    dual b = dsub(dconst(1.0), dscale(dadd(x[0], x[1]), 1.0 / K));
    du[0] = dsub(dmul(x[0], a), dscale(dmul(x[0], p[0]), c1));
    du[1] = dsub(dmul(x[1], b), dscale(dmul(x[1], p[0]), c2));
    dual e0 = daddc(x[0], -1.0), e1 = daddc(x[1], -1.0);
    du[2] = dadd(dmul(e0, e0), dmul(e1, e1));
}
/* f8.jl:9-41: states (x0, x1, x2, obj), p = (T, u) */
static void rhs_f8aircraft4(const dual *x, const dual *p, dual *du)
{
    const double xi = 0.05236;
    dual x0sq = dmul(x[0], x[0]), x0cu = dmul(x0sq, x[0]);
    dual d1 = dadd(dscale(x[0], -0.877), x[2]);
    d1 = dsub(d1, dscale(dmul(x[0], x[2]), 0.088));
    d1 = dadd(d1, dscale(x0sq, 0.47));
    d1 = dsub(d1, dscale(dmul(x[1], x[1]), 0.019));
    d1 = dsub(d1, dmul(x0sq, x[2]));
    d1 = dadd(d1, dscale(x0cu, 3.846));
    dual q1 = dsub(dconst(0.215 * xi - 0.63 * xi * xi * xi), dscale(x0sq, 0.28 * xi));
    d1 = dadd(dadd(d1, q1), dscale(x[0], 0.47 * xi * xi));
    d1 = dsub(d1, dmul(q1, dscale(p[1], 2.0)));
    dual d3 = dsub(dsub(dsub(dscale(x[0], -4.208), dscale(x[2], 0.396)), dscale(x0sq, 0.47)), dscale(x0cu, 3.564));
    dual q3 = dsub(dconst(20.967 * xi - 61.4 * xi * xi * xi), dscale(x0sq, 6.265 * xi));
    d3 = dadd(dadd(d3, q3), dscale(x[0], 46.0 * xi * xi));
    d3 = dsub(d3, dmul(q3, dscale(p[1], 2.0)));
    du[0] = dmul(p[0], d1);
    du[1] = dmul(p[0], x[2]);
    du[2] = dmul(p[0], d3);
    du[3] = p[0];
}
/* double_oscillator.jl:9-38 with time as a state: (x0, x1, x2, x3, tau, cost), p = (u) */
static void rhs_doubleosc6(const dual *x, const dual *p, dual *du)
{
    const double m1 = 200.0, m2 = 2.0, k1 = 100.0, k2 = 3.0, c = 0.5;
    du[0] = x[2];
    du[1] = x[3];
    du[2] = dadd(dadd(dscale(x[0], -(k1 + k2) / m1), dscale(x[1], k2 / m1)), dscale(dsin(x[4]), 1.0 / m1));
    du[3] = dsub(dsub(dscale(x[0], k2 / m2), dscale(x[1], k2 / m2)),
                 dscale(dmul(dsub(dconst(1.0), p[0]), x[3]), c / m2));
    du[4] = dconst(1.0);
    du[5] = dscale(dadd(dadd(dmul(x[0], x[0]), dmul(x[1], x[1])), dmul(p[0], p[0])), 0.5);
}

/* lib/OptimalControlBenchmarks/src/problems/bryson_denham.jl:9-36; states (x, v, cost), p = (w) */
static void rhs_brysondenham3(const dual *x, const dual *p, dual *du)
{
    du[0] = x[1];
    du[1] = p[0];
    du[2] = dmul(p[0], p[0]);
}
/* lib/OptimalControlBenchmarks/src/problems/catalyst_mixing.jl:9-22; states (x1, x2), p = (w) */
static void rhs_catalystmixing2(const dual *x, const dual *p, dual *du)
{
    dual a = dsub(dscale(x[1], 10.0), x[0]);
    du[0] = dmul(p[0], a);
    du[1] = dsub(dmul(p[0], dsub(x[0], dscale(x[1], 10.0))), dmul(dsub(dconst(1.0), p[0]), x[1]));
}
/* lib/OptimalControlBenchmarks/src/problems/cushioned_oscillation.jl:9-37; states (x, v, obj), p = (t_s, u) */
static void rhs_cushioned3(const dual *x, const dual *p, dual *du)
{
    const double m = 5.0, c = 10.0;
    du[0] = dmul(p[0], x[1]);
    du[1] = dmul(p[0], dscale(dsub(p[1], dscale(x[0], c)), 1.0 / m));
    du[2] = p[0];
}
/* lib/OptimalControlBenchmarks/src/problems/denbigh.jl:9-37; states (x1, x2, x3), p = (T); the second equation is `k1 x1 - k3 + k4 x2` as the reference writes it */
static void rhs_denbigh3(const dual *x, const dual *p, dual *du)
{
    dual e1 = dexp(ddiv(dconst(-3.0e3), p[0])), e2 = dexp(ddiv(dconst(-6.0e3), p[0]));
    dual k1 = dscale(e1, 1.0e3), k2 = dscale(e2, 1.0e7), k3 = dscale(e1, 1.0e1);
    du[0] = dsub(dneg(dmul(k1, x[0])), dmul(k2, x[0]));
    du[1] = dadd(dsub(dmul(k1, x[0]), k3), dscale(x[1], 1.0e-3));
    du[2] = dmul(k3, x[1]);
}
/* lib/OptimalControlBenchmarks/src/problems/dielectrophoretic_particle.jl:9-34; states (x0, x1, obj), p = (t_s, u) */
static void rhs_dielectrophoretic3(const dual *x, const dual *p, dual *du)
{
    du[0] = dmul(p[0], dadd(dmul(x[1], p[1]), dscale(dmul(p[1], p[1]), -0.75)));
    du[1] = dmul(p[0], dsub(p[1], x[1]));
    du[2] = p[0];
}
/* lib/OptimalControlBenchmarks/src/problems/electric_car.jl:9-62; states (x1, x2, x3, cost), p = (u) */
static void rhs_electriccar4(const dual *x, const dual *p, dual *du)
{
    const double Kr = 10.0, rho = 1.293, Cx = 0.4, S = 2.0, r = 0.33, Kf = 0.03, Km = 0.27, Rm = 0.03, Lm = 0.05,
                 M = 250.0, g = 9.81, V = 150.0, Rb = 0.05;
    du[0] = dscale(dsub(dsub(dscale(p[0], V), dscale(x[0], Rm)), dscale(x[1], Km)), 1.0 / Lm);
    dual drag = daddc(dscale(dmul(x[1], x[1]), 0.5 * rho * S * Cx * r * r / (Kr * Kr)), M * g * Kf);
    du[1] = dscale(dsub(dscale(x[0], Km), dscale(drag, r / Kr)), Kr * Kr / (M * r * r));
    du[2] = dscale(x[1], r / Kr);
    du[3] = dadd(dmul(dscale(p[0], V), x[0]), dscale(dmul(x[0], x[0]), Rb));
}
/* lib/OptimalControlBenchmarks/src/problems/fuller.jl:9-32; states (x0, x1, cost), p = (u) */
static void rhs_fuller3(const dual *x, const dual *p, dual *du)
{
    du[0] = x[1];
    du[1] = dsub(dconst(1.0), dscale(p[0], 2.0));
    du[2] = dmul(x[0], x[0]);
}
/* lib/OptimalControlBenchmarks/src/problems/jackson.jl:9-50; states (x1, x2, x3), p = (u) */
static void rhs_jackson3(const dual *x, const dual *p, dual *du)
{
    dual a = dsub(x[0], dscale(x[1], 10.0));
    dual b = dmul(dsub(dconst(1.0), p[0]), x[1]);
    du[0] = dneg(dmul(p[0], a));
    du[1] = dsub(dmul(p[0], a), b);
    du[2] = b;
}
/* lib/OptimalControlBenchmarks/src/problems/mountain_car.jl:9-31; states (x, v, obj), p = (t_s, u) */
static void rhs_mountaincar3(const dual *x, const dual *p, dual *du)
{
    du[0] = dmul(p[0], x[1]);
    du[1] = dmul(p[0], dsub(dscale(p[1], 1.0e-3), dscale(dcos(dscale(x[0], 3.0)), 2.5e-3)));
    du[2] = p[0];
}
/* lib/OptimalControlBenchmarks/src/problems/particle_steering.jl:9-42; states (x1, x2, dx1, dx2, obj), p = (t_s, u) */
static void rhs_particlesteering5(const dual *x, const dual *p, dual *du)
{
    du[0] = dmul(p[0], x[2]);
    du[1] = dmul(p[0], x[3]);
    du[2] = dmul(p[0], dscale(dcos(p[1]), 100.0));
    du[3] = dmul(p[0], dscale(dsin(p[1]), 100.0));
    du[4] = p[0];
}
/* lib/OptimalControlBenchmarks/src/problems/rao_mease.jl:9-26; states (x, cost), p = (u) */
static void rhs_raomease2(const dual *x, const dual *p, dual *du)
{
    du[0] = dadd(dneg(dmul(dmul(x[0], x[0]), x[0])), p[0]);
    du[1] = dadd(dmul(x[0], x[0]), dmul(p[0], p[0]));
}
/* lib/OptimalControlBenchmarks/src/problems/robbins.jl:9-44; states (x1, x2, x3, cost), p = (u); alpha = 3, beta = 0, gamma = 0.5 */
static void rhs_robbins4(const dual *x, const dual *p, dual *du)
{
    du[0] = x[1];
    du[1] = x[2];
    du[2] = p[0];
    du[3] = dadd(dadd(dscale(x[0], 3.0), dscale(dmul(x[0], x[0]), 0.0)), dscale(dmul(p[0], p[0]), 0.5));
}
/* lib/OptimalControlBenchmarks/src/problems/moon_landing.jl:9-31; states (h, v, m), p = (t_s, T) */
static void rhs_moonlanding3(const dual *x, const dual *p, dual *du)
{
    du[0] = dmul(p[0], x[1]);
    du[1] = dmul(p[0], daddc(ddiv(p[1], x[2]), -1.0));
    du[2] = dmul(p[0], dneg(dscale(p[1], 1.0 / 2.349)));
}
/* lib/OptimalControlBenchmarks/src/problems/lotka_shared.jl:9-32; states (x0, x1, x2, cost), p = (u) */
static void rhs_lotkashared4(const dual *x, const dual *p, dual *du)
{
    du[0] = dsub(dsub(x[0], dmul(x[0], x[1])), dmul(x[0], x[2]));
    du[1] = dsub(dadd(dneg(x[1]), dmul(x[0], x[1])), dmul(dscale(x[1], 0.1), p[0]));
    du[2] = dsub(dadd(dneg(x[2]), dmul(dscale(x[0], 1.2), x[2])), dmul(dscale(x[2], 0.4), p[0]));
    dual a = daddc(x[0], -1.5), b = daddc(x[1], -1.0), c = daddc(x[2], -1.0);
    du[3] = dadd(dadd(dmul(a, a), dmul(b, b)), dmul(c, c));
}
/* lib/OptimalControlBenchmarks/src/problems/hanging_chain.jl:9-48; states (x1, x2, x3), p = (u) */
static void rhs_hangingchain3(const dual *x, const dual *p, dual *du)
{
    dual s = dsqrt(daddc(dmul(p[0], p[0]), 1.0));
    du[0] = p[0];
    du[1] = dmul(x[0], s);
    du[2] = s;
}
/* lib/OptimalControlBenchmarks/src/problems/tubular_reactor.jl:9-22; states (x1, x2), p = (u) */
static void rhs_tubularreactor2(const dual *x, const dual *p, dual *du)
{
    du[0] = dneg(dmul(dadd(p[0], dscale(dmul(p[0], p[0]), 0.5)), x[0]));
    du[1] = dmul(p[0], x[0]);
}
/* lib/OptimalControlBenchmarks/src/problems/ocean.jl:9-64; states (S, R, tau, cost), p = (u1, u2); time rides along as a state (tau' = 1) for the discount factor */
static void rhs_ocean4(const dual *x, const dual *p, dual *du)
{
    dual Dl = dsub(dsub(dconst(3.5e4), x[1]), x[0]);
    du[0] = dneg(p[0]);
    du[1] = dsub(dsub(p[0], p[1]), dscale(dsub(x[0], dscale(Dl, 0.1)), 1.0e-3));
    du[2] = dconst(1.0);
    dual U = dsub(dscale(p[0], 50.0), dscale(dmul(p[0], p[0]), 0.5));
    dual A = dadd(dscale(p[1], 2.0), dscale(dmul(p[1], p[1]), 2.0));
    dual C = dsub(dconst(50.0), dscale(x[1], 4.0e-3));
    dual q = daddc(dscale(x[0], 0.3), -6.0e2);
    du[3] = dneg(dmul(dexp(dscale(x[2], -3.0e-2)), dsub(dsub(dsub(U, A), dmul(p[0], C)), dmul(q, q))));
}
/* lib/OptimalControlBenchmarks/src/problems/ducted_fan.jl:9-61; states (x1, v1, x2, v2, alpha, valpha, cost), p = (t_s, u1, u2) */
static void rhs_ductedfan7(const dual *x, const dual *p, dual *du)
{
    const double m = 2.2, J = 0.05, r = 0.2, mg = 4.0, mu = 1.0;
    dual s = dsin(x[4]), c = dcos(x[4]);
    du[0] = dmul(p[0], x[1]);
    du[1] = dmul(p[0], dscale(dsub(dmul(p[1], c), dmul(p[2], s)), 1.0 / m));
    du[2] = dmul(p[0], x[3]);
    du[3] = dmul(p[0], dscale(dadd(daddc(dmul(p[1], s), -mg), dmul(p[2], c)), 1.0 / m));
    du[4] = dmul(p[0], x[5]);
    du[5] = dmul(p[0], dscale(p[1], r / J));
    du[6] = dadd(dadd(dscale(dmul(p[1], p[1]), 2.0), dmul(p[2], p[2])), dscale(p[0], mu));
}
/* lib/OptimalControlBenchmarks/src/problems/hang_glider.jl:9-66; states (x, y, vx, vy), p = (T, cL) */
static void rhs_hangglider4(const dual *x, const dual *p, dual *du)
{
    const double uc = 2.5, rc = 100.0, c0 = 0.034, c1 = 0.069662, S = 14.0, rho = 1.13, m = 100.0, g = 9.81;
    dual rr = daddc(dscale(x[0], 1.0 / rc), -2.5);
    dual r = dmul(rr, rr);
    dual Uup = dmul(dscale(dsub(dconst(1.0), r), uc), dexp(dneg(r)));
    dual w = dsub(x[3], Uup);
    dual v2 = dadd(dmul(x[2], x[2]), dmul(w, w));
    dual v = dsqrt(v2);
    dual Dr = dmul(dscale(daddc(dscale(dmul(p[1], p[1]), c1), c0), 0.5 * rho * S), v2);
    dual L = dmul(dscale(p[1], 0.5 * rho * S), v2);
    dual mv = dscale(v, m);
    du[0] = dmul(p[0], x[2]);
    du[1] = dmul(p[0], x[3]);
    du[2] = ddiv(dmul(dneg(p[0]), dadd(dmul(L, w), dmul(Dr, x[2]))), mv);
    du[3] = dmul(p[0], daddc(ddiv(dsub(dmul(L, x[2]), dmul(Dr, w)), mv), -g));
}

typedef void (*rhs_fn)(const dual *, const dual *, dual *);
typedef void (*jac_fn)(const dual *, const dual *, dual *, dual *);
/* variants of the augmentation (augmentation.jl): which F states exist and what they integrate */
enum { AUG_CS = 0, /* ContinuousSampled, non-fixed :206-253: F' = sum_i w_i g_i'g_i, p = [base; w]         */
       AUG_CU = 1, /* Continuous, no measurements :217-220: F' = (sum_i g_i)'(sum_i g_i), p = base        */
       AUG_CF = 2, /* ContinuousSampled, fixed :255-292: F_k' = g_k'g_k per observed k, p = [base; w]      */
       AUG_D = 3   /* Discrete / DiscreteSampled :171-203: no F states                                     */ };
static void rhs_oed(rhs_fn f, jac_fn jac, int nx, int npb, int nF, int nobs, int variant, const dual *u,
                    const dual *p, dual *du)
{
    dual Jx[CO_MAX_NX], Jp[CO_MAX_NX];
    f(u, p, du);
    jac(u, p, Jx, Jp);
    const dual *G = u + nx; /* G[r + nx*c] */
    dual *dG = du + nx;
    for (int c = 0; c < nF; c++)
        for (int r = 0; r < nx; r++) {
            dual acc = Jp[r * nF + c];
            for (int k = 0; k < nx; k++) acc = dadd(acc, dmul(Jx[r * nx + k], G[k + nx * c]));
            dG[r + nx * c] = acc;
        }
    dual *dF = du + nx + nx * nF;
    const dual *w = p + npb;
    int q = 0;
    if (variant == AUG_CS) {
        for (int b = 0; b < nF; b++)
            for (int a = 0; a <= b; a++) {
                dual acc = dconst(0.0);
                for (int i = 0; i < nobs; i++) acc = dadd(acc, dmul(w[i], dmul(G[i + nx * a], G[i + nx * b])));
                dF[q++] = acc;
            }
    } else if (variant == AUG_CU) { /* rows of h_x G are summed first, then the outer product (:217-220) */
        dual sg[CO_MAX_NX];
        for (int a = 0; a < nF; a++) {
            sg[a] = dconst(0.0);
            for (int i = 0; i < nobs; i++) sg[a] = dadd(sg[a], G[i + nx * a]);
        }
        for (int b = 0; b < nF; b++)
            for (int a = 0; a <= b; a++) dF[q++] = dmul(sg[a], sg[b]);
    } else if (variant == AUG_CF) { /* _F = vcat over k of F[k,:,:][selector] (:271) */
        for (int i = 0; i < nobs; i++)
            for (int b = 0; b < nF; b++)
                for (int a = 0; a <= b; a++) dF[q++] = dmul(G[i + nx * a], G[i + nx * b]);
    }
}

static void model_rhs(const model_ctx *m, const dual *u, const dual *p, dual *du)
{
    switch (m->model) {
    case CO_LOTKA3: rhs_lotka3(u, p, du); break;
    case CO_LQR: rhs_lqr(u, p, du); break;
    case CO_LOTKA2: rhs_lotka2(u, p, du); break;
    case CO_LOTKA2_OED: rhs_oed(rhs_lotka2, jac_lotka2, 2, 3, 2, 2, AUG_CS, u, p, du); break;
    case CO_LIN1D: rhs_lin1d(u, p, du); break;
    case CO_LIN1D_OED: rhs_oed(rhs_lin1d, jac_lin1d, 1, 1, 1, 1, AUG_CS, u, p, du); break;
    case CO_DENSE20: rhs_dense(20, u, p, du, m->data); break;
    case CO_DENSE16: rhs_dense(16, u, p, du, m->data); break;
    case CO_DENSE24: rhs_dense(24, u, p, du, m->data); break;
    case CO_DENSE32: rhs_dense(32, u, p, du, m->data); break;
    case CO_EGERSTEDT: rhs_egerstedt(u, p, du); break;
    case CO_VDP3: rhs_vdp3(u, p, du); break;
    case CO_CARTPOLE5: rhs_cartpole5(u, p, du); break;
    case CO_ROCKET3: rhs_rocket3(u, p, du); break;
    case CO_QUADROTOR7: rhs_quadrotor7(u, p, du); break;
    case CO_THREETANK4: rhs_threetank4(u, p, du); break;
    case CO_BATCHREACTOR2: rhs_batchreactor2(u, p, du); break;
    case CO_LOTKACOMP3: rhs_lotkacomp3(u, p, du); break;
    case CO_F8AIRCRAFT4: rhs_f8aircraft4(u, p, du); break;
    case CO_DOUBLEOSC6: rhs_doubleosc6(u, p, du); break;
    case CO_LOTKA2_OED_CU: rhs_oed(rhs_lotka2, jac_lotka2, 2, 3, 2, 2, AUG_CU, u, p, du); break;
    case CO_LOTKA2_OED_CF: rhs_oed(rhs_lotka2, jac_lotka2, 2, 3, 2, 2, AUG_CF, u, p, du); break;
    case CO_LOTKA2_OED_DS:
    case CO_LOTKA2_OED_DU: rhs_oed(rhs_lotka2, jac_lotka2, 2, 3, 2, 2, AUG_D, u, p, du); break;
    case CO_LIN1D_OED_CF: rhs_oed(rhs_lin1d, jac_lin1d, 1, 1, 1, 1, AUG_CF, u, p, du); break;
    case CO_LIN1D_OED_DS: rhs_oed(rhs_lin1d, jac_lin1d, 1, 1, 1, 1, AUG_D, u, p, du); break;
    case CO_BRYSONDENHAM3: rhs_brysondenham3(u, p, du); break;
    case CO_CATALYSTMIXING2: rhs_catalystmixing2(u, p, du); break;
    case CO_CUSHIONED3: rhs_cushioned3(u, p, du); break;
    case CO_DENBIGH3: rhs_denbigh3(u, p, du); break;
    case CO_DIELECTROPHORETIC3: rhs_dielectrophoretic3(u, p, du); break;
    case CO_ELECTRICCAR4: rhs_electriccar4(u, p, du); break;
    case CO_FULLER3: rhs_fuller3(u, p, du); break;
    case CO_JACKSON3: rhs_jackson3(u, p, du); break;
    case CO_MOUNTAINCAR3: rhs_mountaincar3(u, p, du); break;
    case CO_PARTICLESTEERING5: rhs_particlesteering5(u, p, du); break;
    case CO_RAOMEASE2: rhs_raomease2(u, p, du); break;
    case CO_ROBBINS4: rhs_robbins4(u, p, du); break;
    case CO_MOONLANDING3: rhs_moonlanding3(u, p, du); break;
    case CO_LOTKASHARED4: rhs_lotkashared4(u, p, du); break;
    case CO_HANGINGCHAIN3: rhs_hangingchain3(u, p, du); break;
    case CO_TUBULARREACTOR2: rhs_tubularreactor2(u, p, du); break;
    case CO_OCEAN4: rhs_ocean4(u, p, du); break;
    case CO_DUCTEDFAN7: rhs_ductedfan7(u, p, du); break;
    case CO_HANGGLIDER4: rhs_hangglider4(u, p, du); break;
    default: break;
    }
}

/* ------------------------------------------------------------------ fixed-step RK */
/* Tsit5 propagation tableau (OrdinaryDiffEqTsit5 constant cache) */
static const double TS_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
     2.324710524099774}};

/* one control piece: K steps of h = (t1-t0)/K; k1 is re-evaluated at the piece start because every piece is
 * a fresh `solve` in the reference (src/single_shooting.jl:443-453); FSAL inside the piece. */
/* ------------------------------------------------------------------ adaptive Tsit5
 * The reference's DEFAULT mode: `solve(problem, Tsit5(); ...)` per control piece (src/single_shooting.jl:443-453) with
 * the problem's abstol / reltol (test/core/multiple_shooting.jl:23: 1e-8 / 1e-6; OrdinaryDiffEq defaults 1e-6 / 1e-3).
 * The stepping itself is OrdinaryDiffEq's (third party, not under /root/reference; compat "1, 2", no Manifest) and is
 * restated here from its published algorithm:
 *   - initial step: Hairer-Norsett-Wanner heuristic as OrdinaryDiffEq's `ode_determine_initdt` applies it
 *     (d0 = |u0/sk|, d1 = |f0/sk|, dt0 = 0.01 d0/d1, one explicit Euler probe, dt1 = 10^(-(2+log10 max(d1,d2))/5),
 *     dt = min(100 dt0, dt1, dtmax)), norms = RMS, sk = abstol + |u0| reltol, dtmax = the span of the piece;
 *   - embedded error estimate utilde = dt sum_j btilde_j k_j, EEst = RMS(utilde / (abstol + max(|uprev|,|u|) reltol));
 *   - PI controller: q11 = EEst^beta1, q = q11 / qold^beta2 (beta1 = 7/50, beta2 = 2/25), q = clamp(q / gamma, 1/qmax,
 *     1/qmin) with gamma = 0.9, qmin = 0.2, qmax = 10; accept iff EEst <= 1, then qold = max(EEst, 1e-4), dt <- dt/q;
 *     reject: dt <- dt / min(1/qmin, q11/gamma); the powers are taken with DiffEqBase's `fastpow` (Float32: Goldberg's
 *     rational approximation of log2, then exp2);
 *   - every step is clipped to the end of the piece and t snaps to it within 100 eps.
 * Every piece is a fresh `solve`: initial step and controller memory start over.
 * PIN: with this restatement the reference's golden vectors G1 / G4 (test/core/multiple_shooting.jl:54-72,220), which
 * a converged fixed-step integration only reaches to 5e-11, are reproduced to ~1e-15 (tests/test_oracle_golden.py) --
 * the accepted step sequence is the reference's.
 * Forward sensitivities ride along on the SAME step sequence (error control on the values only): the derivative of
 * the computed map with the steps frozen.  (ForwardDiff through `solve` lets the partials take part in the norm, so
 * the reference's own derivatives depend on the AD chunking; the values do not.) */
static double g_abstol = 1.0e-6, g_reltol = 1.0e-3;
void co_set_tolerances(double abstol, double reltol)
{
    g_abstol = abstol;
    g_reltol = reltol;
}
static const double TS_BT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                                -0.1447110071732629,     0.5823571654525552,     -0.45808210592918697,
                                0.015151515151515152};
/* DiffEqBase fastpow(x, y) = exp2(Float32(y) * fastlog2(Float32(x))) */
static float co_fastlog2(float x)
{
    const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
    unsigned ux, ux2;
    memcpy(&ux, &x, 4);
    const unsigned e = (ux & 0x7F800000u) >> 23;
    volatile float signif, fexp, t1, t2, t3; /* volatile: no FMA contraction, as in Julia without @muladd */
    float sf;
    if (ux & 0x00400000u) {
        ux2 = (ux & 0x007FFFFFu) | 0x3f000000u;
        fexp = (float)e - 126.0f;
    } else {
        ux2 = (ux & 0x007FFFFFu) | 0x3f800000u;
        fexp = (float)e - 127.0f;
    }
    memcpy(&sf, &ux2, 4);
    signif = sf - 1.0f;
    t1 = a * signif;
    t1 = t1 + b;
    t2 = t1 * signif;
    t3 = signif + c;
    t2 = t2 / t3;
    return fexp + t2;
}
static double co_fastpow(double x, double y)
{
    if (x == 0.0) return 0.0;
    volatile float e = (float)y * co_fastlog2((float)x);
    return (double)(float)exp2((double)e); /* Float32 result, correctly rounded */
}
static double rms_state(const double *v, int n)
{
    double s = 0.0;
    for (int c = 0; c < n; c++) s += v[c] * v[c];
    return sqrt(s / (double)n);
}
/* Tsit5's continuous extension (Tsitouras 2011; OrdinaryDiffEq's interpolant for Tsit5, used by `saveat`):
 *   y(t + theta dt) = y + dt sum_i b_i(theta) k_i,  b_1 = theta (r11 + theta (r12 + theta (r13 + theta r14))),
 *   b_i = theta^2 (ri2 + theta (ri3 + theta ri4)), i = 2..7.  b_i(1) = a_7i and the order conditions up to 4 hold for every
 *   theta (tests/test_saveat_host.py). */
static const double TS_R[7][4] = {{1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
                                  {0.0, 0.13169999999999998, -0.2234, 0.1017},
                                  {0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253},
                                  {0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902},
                                  {0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928},
                                  {0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661},
                                  {0.0, 1.5, -4.0, 2.5}};
void co_tsit5_interp_weights(double th, double *bw)
{
    bw[0] = th * (TS_R[0][0] + th * (TS_R[0][1] + th * (TS_R[0][2] + th * TS_R[0][3])));
    for (int i = 1; i < 7; i++) bw[i] = th * th * (TS_R[i][1] + th * (TS_R[i][2] + th * TS_R[i][3]));
}
/* pending saveat times of one piece: sv[0..n), results into out[m * nx + c]; `done` counts the emitted ones */
typedef struct {
    int n, done;
    const double *t;
    dual *out;
} sv_state;
/* samples of the pending times <= tnew from the step (tprev, dt) that starts at y0 with stage derivatives k[7] */
static void sv_emit(sv_state *sv, int nx, double tprev, double dt, double tnew, const dual *y0, dual (*k)[CO_MAX_NX])
{
    while (sv && sv->done < sv->n && sv->t[sv->done] <= tnew) {
        double bw[7];
        co_tsit5_interp_weights((sv->t[sv->done] - tprev) / dt, bw);
        for (int c = 0; c < nx; c++) {
            dual acc = dscale(k[0][c], bw[0]);
            for (int j = 1; j < 7; j++) daxpy(&acc, bw[j], &k[j][c]);
            dual z = y0[c];
            daxpy(&z, dt, &acc);
            sv->out[(size_t)sv->done * nx + c] = z;
        }
        sv->done++;
    }
}

static void integrate_piece_adaptive(const model_ctx *m, int nx, dual *x, const dual *p, double t0, double t1, sv_state *sv)
{
    dual k[7][CO_MAX_NX], tmp[CO_MAX_NX], xn[CO_MAX_NX];
    double w[CO_MAX_NX], sk[CO_MAX_NX];
    const double abstol = g_abstol, reltol = g_reltol, dtmax = t1 - t0;
    const double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 0.9, qmin = 0.2, qmax = 10.0, qoldinit = 1.0e-4;
    double t = t0, dt, qold = qoldinit;
    model_rhs(m, x, p, k[0]);
    { /* initial step */
        const double smalldt = 1.0e-6;
        for (int c = 0; c < nx; c++) sk[c] = abstol + fabs(x[c].v) * reltol;
        for (int c = 0; c < nx; c++) w[c] = x[c].v / sk[c];
        const double d0 = rms_state(w, nx);
        for (int c = 0; c < nx; c++) w[c] = k[0][c].v / sk[c];
        const double d1 = rms_state(w, nx);
        double dt0 = (d0 < 1.0e-5 || d1 < 1.0e-5) ? smalldt : 0.01 * (d0 / d1);
        if (dt0 > dtmax) dt0 = dtmax;
        if (dt0 < 10.0 * 2.220446049250313e-16) {
            dt = smalldt;
        } else {
            const int nd = g_nd;
            g_nd = 0; /* the probe needs values only */
            for (int c = 0; c < nx; c++) tmp[c] = dconst(x[c].v + dt0 * k[0][c].v);
            model_rhs(m, tmp, p, xn);
            g_nd = nd;
            for (int c = 0; c < nx; c++) w[c] = (xn[c].v - k[0][c].v) / sk[c];
            const double d2 = rms_state(w, nx) / dt0;
            const double md = d1 > d2 ? d1 : d2;
            const double dt1 = md <= 1.0e-15 ? (smalldt > dt0 * 1.0e-3 ? smalldt : dt0 * 1.0e-3)
                                             : pow(10.0, -(2.0 + log10(md)) / 5.0);
            dt = 100.0 * dt0;
            if (dt1 < dt) dt = dt1;
            if (dtmax < dt) dt = dtmax;
        }
    }
    for (long iter = 0; t < t1; iter++) {
        if (dt > dtmax) dt = dtmax;
        if (dt > t1 - t) dt = t1 - t;
        if (iter >= 100000L || !(dt > 0.0)) { /* OrdinaryDiffEq: retcode MaxIters (maxiters = 1e5) / DtLessThanMin */
            for (int c = 0; c < nx; c++) x[c] = dconst(NAN);
            return;
        }
        for (int st = 1; st < 6; st++) {
            for (int c = 0; c < nx; c++) {
                dual acc = dscale(k[0][c], TS_A[st][0]);
                for (int j = 1; j < st; j++) daxpy(&acc, TS_A[st][j], &k[j][c]);
                tmp[c] = x[c];
                daxpy(&tmp[c], dt, &acc);
            }
            model_rhs(m, tmp, p, k[st]);
        }
        for (int c = 0; c < nx; c++) {
            dual acc = dscale(k[0][c], TS_A[6][0]);
            for (int j = 1; j < 6; j++) daxpy(&acc, TS_A[6][j], &k[j][c]);
            xn[c] = x[c];
            daxpy(&xn[c], dt, &acc);
        }
        model_rhs(m, xn, p, k[6]);
        for (int c = 0; c < nx; c++) {
            double ut = 0.0;
            for (int j = 0; j < 7; j++) ut += TS_BT[j] * k[j][c].v;
            const double a0 = fabs(x[c].v), a1 = fabs(xn[c].v);
            w[c] = dt * ut / (abstol + (a0 > a1 ? a0 : a1) * reltol);
        }
        const double EEst = rms_state(w, nx);
        if (EEst != EEst) { /* non-finite: the reference would stop with retcode Unstable; poison the state */
            for (int c = 0; c < nx; c++) x[c] = dconst(NAN);
            return;
        }
        double q, q11 = 0.0;
        if (EEst == 0.0) {
            q = 1.0 / qmax;
        } else {
            q11 = co_fastpow(EEst, beta1);
            q = q11 / co_fastpow(qold, beta2);
            q = q / gamma;
            if (q > 1.0 / qmin) q = 1.0 / qmin;
            if (q < 1.0 / qmax) q = 1.0 / qmax;
        }
        if (EEst <= 1.0) {
            qold = EEst > qoldinit ? EEst : qoldinit;
            const double dtnew = dt / q;
            double tt = t + dt;
            const double big = fabs(t1);
            if (fabs(tt - t1) < 100.0 * (nextafter(big, INFINITY) - big)) tt = t1;
            sv_emit(sv, nx, t, dt, tt, x, k); /* saveat times inside the accepted step: dense output */
            t = tt;
            for (int c = 0; c < nx; c++) {
                x[c] = xn[c];
                k[0][c] = k[6][c];
            }
            dt = dtnew < dtmax ? dtnew : dtmax;
        } else {
            const double r = q11 / gamma;
            dt = dt / (r < 1.0 / qmin ? r : 1.0 / qmin);
        }
    }
}

static void integrate_piece(const model_ctx *m, int nx, dual *x, const dual *p, double t0, double t1, int K,
                            int method, sv_state *sv)
{
    if (method == CO_TSIT5_ADAPTIVE) {
        integrate_piece_adaptive(m, nx, x, p, t0, t1, sv);
        return;
    }
    dual k[7][CO_MAX_NX], tmp[CO_MAX_NX], x0[CO_MAX_NX];
    double h = (t1 - t0) / (double)K;
    if (method == CO_TSIT5) {
        model_rhs(m, x, p, k[0]);
        for (int s = 0; s < K; s++) {
            for (int c = 0; c < nx; c++) x0[c] = x[c];
            for (int st = 1; st < 7; st++) {
                for (int c = 0; c < nx; c++) {
                    dual acc = x[c];
                    for (int j = 0; j < st; j++) daxpy(&acc, h * TS_A[st][j], &k[j][c]);
                    tmp[c] = acc;
                }
                if (st == 6)
                    for (int c = 0; c < nx; c++) x[c] = tmp[c];
                model_rhs(m, tmp, p, k[st]);
            }
            /* saveat times inside this step, from its dense output (the steps are not cut at them) */
            sv_emit(sv, nx, t0 + s * h, h, (s + 1 == K) ? t1 : t0 + (s + 1) * h, x0, k);
            for (int c = 0; c < nx; c++) k[0][c] = k[6][c];
        }
    } else { /* classical RK4 */
        for (int s = 0; s < K; s++) {
            model_rhs(m, x, p, k[0]);
            for (int c = 0; c < nx; c++) {
                tmp[c] = x[c];
                daxpy(&tmp[c], 0.5 * h, &k[0][c]);
            }
            model_rhs(m, tmp, p, k[1]);
            for (int c = 0; c < nx; c++) {
                tmp[c] = x[c];
                daxpy(&tmp[c], 0.5 * h, &k[1][c]);
            }
            model_rhs(m, tmp, p, k[2]);
            for (int c = 0; c < nx; c++) {
                tmp[c] = x[c];
                daxpy(&tmp[c], h, &k[2][c]);
            }
            model_rhs(m, tmp, p, k[3]);
            for (int c = 0; c < nx; c++) {
                daxpy(&x[c], h / 6.0, &k[0][c]);
                daxpy(&x[c], h / 3.0, &k[1][c]);
                daxpy(&x[c], h / 3.0, &k[2][c]);
                daxpy(&x[c], h / 6.0, &k[3][c]);
            }
        }
    }
}

/* ------------------------------------------------------------------ index logic */
int co_restrict_grid(const double *t, int nt, double t0, double t1, int *pos)
{
    int n = 0;
    for (int k = 0; k < nt; k++)
        if (t0 <= t[k] && t[k] < t1) pos[n++] = k; /* local_controls.jl:55 */
    return n;
}

static int cmp_double(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* Julia searchsortedlast: last index k (1-based) with a[k] <= x, 0 if none */
static int searchsortedlast(const double *a, int n, double x)
{
    int lo = 0, hi = n + 1;
    while (lo < hi - 1) {
        int mid = lo + ((hi - lo) >> 1);
        if (x < a[mid - 1])
            hi = mid;
        else
            lo = mid;
    }
    return lo;
}
/* Julia clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)) */
static int64_t jl_clamp(int64_t x, int64_t lo, int64_t hi) { return x > hi ? hi : (x < lo ? lo : x); }

int co_build_index_grid(int nc, const double *const *t, const int *nt, double t0, double t1, int64_t *idx,
                        double *grid, int cap_pieces)
{
    int total = 2;
    for (int i = 0; i < nc; i++) total += nt[i];
    double *all = (double *)malloc(sizeof(double) * (size_t)total);
    double **ts = (double **)malloc(sizeof(double *) * (size_t)(nc > 0 ? nc : 1));
    int *nts = (int *)malloc(sizeof(int) * (size_t)(nc > 0 ? nc : 1));
    int n = 0;
    for (int i = 0; i < nc; i++) {
        ts[i] = (double *)malloc(sizeof(double) * (size_t)(nt[i] > 0 ? nt[i] : 1));
        nts[i] = 0;
        for (int k = 0; k < nt[i]; k++)
            if (t0 <= t[i][k] && t[i][k] < t1) {
                ts[i][nts[i]++] = t[i][k];
                all[n++] = t[i][k];
            }
    }
    all[n++] = t0;
    all[n++] = t1;
    qsort(all, (size_t)n, sizeof(double), cmp_double);
    int ng = 0;
    for (int k = 0; k < n; k++) { /* unique! then filter!(isfinite) (:153) */
        if (k > 0 && all[k] == all[k - 1]) continue;
        if (!isfinite(all[k])) continue;
        if (ng < cap_pieces + 1) grid[ng] = all[k];
        ng++;
    }
    int M = ng - 1;
    if (M > cap_pieces) {
        M = -M; /* caller's buffers too small */
        goto done;
    }
    for (int i = 0; i < nc; i++)
        for (int j = 0; j < M; j++)
            idx[i + (size_t)nc * j] = jl_clamp(searchsortedlast(ts[i], nts[i], grid[j]), 1, nts[i]); /* :156-159 */
    for (int i = 1; i < nc; i++) { /* :162-168, uses the already offset previous row */
        int64_t mx = idx[(i - 1)];
        for (int j = 1; j < M; j++)
            if (idx[(i - 1) + (size_t)nc * j] > mx) mx = idx[(i - 1) + (size_t)nc * j];
        if (M == 0) mx = 0;
        for (int j = 0; j < M; j++) idx[i + (size_t)nc * j] += mx;
    }
    if (nc > 0 && M > 0) { /* :172 */
        int64_t mn = idx[0];
        for (int q = 0; q < nc * M; q++)
            if (idx[q] < mn) mn = idx[q];
        for (int q = 0; q < nc * M; q++) idx[q] -= mn - 1;
    }
done:
    for (int i = 0; i < nc; i++) free(ts[i]);
    free(ts);
    free(nts);
    free(all);
    return M;
}

int co_subvector_indices(int M, int L, int *ss)
{
    int N = M / L, n = 0;
    for (int i = 1; i <= N; i++) {
        ss[2 * n] = (i - 1) * L + 1;
        ss[2 * n + 1] = i * L;
        n++;
    }
    if (M - N * L > 0) {
        ss[2 * n] = N * L + 1;
        ss[2 * n + 1] = M;
        n++;
    }
    return n;
}

/* ------------------------------------------------------------------ layer */
typedef struct {
    double t0, t1;
    int shooting; /* i > 1 */
    int M;
    double *grid;  /* M+1 */
    int64_t *idx;  /* nact x M col-major, 1-based */
    int n_u0, n_c; /* lengths of ps.u0 and ps.controls of this interval */
    int n_const;
    int ic_constants[CO_MAX_NX]; /* 1-based */
    int ic_sorting[CO_MAX_NX];   /* 1-based */
    int n_sidx;
    int sidx[CO_MAX_NX + CO_MAX_CTRL]; /* shooting_indices, 1-based over [states; controls] */
    int var_off;                       /* offset of the block in the flat ps */
    double *c_init;                    /* default local controls */
    int NS;                            /* samples after the start sample: M + saveat times strictly inside pieces */
    int *sv_n;                         /* [M] saveat times strictly inside piece j */
    int *sv_off;                       /* [M] offset of the piece's times in sv_t */
    double *sv_t;                      /* those times */
} interval;

struct co_layer {
    model_ctx m;
    int nx, np_all;
    double u0[CO_MAX_NX], p0[CO_MAX_NP];
    double tspan[2];
    int nact;                     /* active controls (control_indices <= np_all), user order kept */
    int cidx[CO_MAX_CTRL];        /* 1-based p index of active control r */
    int ntp;                      /* tunable parameters */
    int tunable_p[CO_MAX_NP];     /* 1-based, ascending: setdiff(eachindex(p), control_indices) */
    int pmap_kind[CO_MAX_NP];     /* p_full[q]: 0 -> ps.p[pmap_idx], 1 -> controls_j[pmap_idx] (scatter A,B :306-323) */
    int pmap_idx[CO_MAX_NP];
    int n_tic, tic[CO_MAX_NX];
    int nquad, quad[CO_MAX_NX];
    int I;
    interval *iv;
    int nvars, ndef, nsamples;
};

static int in_list(const int *l, int n, int v)
{
    for (int k = 0; k < n; k++)
        if (l[k] == v) return 1;
    return 0;
}

/* sortperm(vcat(constants, tunables)) -- stable, all entries distinct */
static void sortperm_cat(const int *a, int na, const int *b, int nb, int *perm)
{
    int n = na + nb, v[2 * CO_MAX_NX];
    for (int k = 0; k < na; k++) v[k] = a[k];
    for (int k = 0; k < nb; k++) v[na + k] = b[k];
    for (int k = 0; k < n; k++) perm[k] = k + 1;
    for (int i = 1; i < n; i++) { /* insertion sort by key */
        int pi = perm[i], j = i - 1;
        while (j >= 0 && v[perm[j] - 1] > v[pi - 1]) {
            perm[j + 1] = perm[j];
            j--;
        }
        perm[j + 1] = pi;
    }
}

co_layer *co_layer_create(const co_problem *P)
{
    co_layer *L = (co_layer *)calloc(1, sizeof(co_layer));
    L->m.model = P->model;
    L->m.data = P->model_data;
    L->nx = P->nx;
    L->np_all = P->np_all;
    memcpy(L->u0, P->u0, sizeof(double) * (size_t)P->nx);
    memcpy(L->p0, P->p0, sizeof(double) * (size_t)P->np_all);
    L->tspan[0] = P->tspan[0];
    L->tspan[1] = P->tspan[1];
    L->n_tic = P->n_tunable_ic;
    for (int k = 0; k < L->n_tic; k++) L->tic[k] = P->tunable_ic[k];
    L->nquad = P->nquad;
    for (int k = 0; k < L->nquad; k++) L->quad[k] = P->quad_idx[k];

    /* active controls: control_indices .<= lastindex(p_vec) (single_shooting.jl:302-304) */
    const double *act_t[CO_MAX_CTRL], *act_u[CO_MAX_CTRL];
    int act_nt[CO_MAX_CTRL];
    int allc[CO_MAX_CTRL];
    for (int r = 0; r < P->ncontrols; r++) {
        allc[r] = P->control_indices[r];
        if (P->control_indices[r] <= P->np_all) {
            L->cidx[L->nact] = P->control_indices[r];
            act_t[L->nact] = P->ctrl_t[r];
            act_u[L->nact] = P->ctrl_u ? P->ctrl_u[r] : NULL;
            act_nt[L->nact] = P->ctrl_nt[r];
            L->nact++;
        }
    }
    /* tunable_p = setdiff(eachindex(p_vec), control_indices) (:109); scatter matrices A, B (:306-318):
     * column ids are handed out in ascending p index, for parameters and for controls alike */
    int pid = 0, cid = 0;
    for (int q = 1; q <= P->np_all; q++) {
        if (in_list(L->cidx, L->nact, q)) {
            L->pmap_kind[q - 1] = 1;
            L->pmap_idx[q - 1] = cid++;
        } else {
            L->pmap_kind[q - 1] = 0;
            L->pmap_idx[q - 1] = pid++;
            L->tunable_p[L->ntp++] = q;
        }
    }
    (void)allc;
    (void)act_u;

    /* shooting intervals: sort!(unique!(vcat(tpoints, tspan))) -> consecutive pairs (multiple_shooting.jl:86-90) */
    double *pts = (double *)malloc(sizeof(double) * (size_t)(P->n_shoot + 2));
    int np = 0;
    for (int k = 0; k < P->n_shoot; k++) pts[np++] = P->shoot_pts[k];
    pts[np++] = P->tspan[0];
    pts[np++] = P->tspan[1];
    qsort(pts, (size_t)np, sizeof(double), cmp_double);
    int nu = 0;
    for (int k = 0; k < np; k++)
        if (k == 0 || pts[k] != pts[k - 1]) pts[nu++] = pts[k];
    L->I = (P->n_shoot > 0) ? nu - 1 : 1;
    L->iv = (interval *)calloc((size_t)L->I, sizeof(interval));

    int off = 0;
    for (int i = 0; i < L->I; i++) {
        interval *it = &L->iv[i];
        it->t0 = (P->n_shoot > 0) ? pts[i] : P->tspan[0];
        it->t1 = (P->n_shoot > 0) ? pts[i + 1] : P->tspan[1];
        it->shooting = (i > 0);
        /* pieces */
        if (L->nact > 0) {
            int cap = 2;
            for (int r = 0; r < L->nact; r++) cap += act_nt[r];
            it->grid = (double *)malloc(sizeof(double) * (size_t)(cap + 1));
            it->idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(L->nact * cap));
            it->M = co_build_index_grid(L->nact, act_t, act_nt, it->t0, it->t1, it->idx, it->grid, cap);
            if (it->M > 0) { /* keep what the interval uses, not room for every grid point of every control */
                it->grid = (double *)realloc(it->grid, sizeof(double) * (size_t)(it->M + 1));
                it->idx = (int64_t *)realloc(it->idx, sizeof(int64_t) * (size_t)(L->nact * it->M));
            }
        } else {
            /* no controls: tspans = (problem.tspan,) -- the *problem's* span, also on sub-intervals
             * (single_shooting.jl:329-332; reference quirk kept) */
            it->M = 1;
            it->grid = (double *)malloc(sizeof(double) * 2);
            it->grid[0] = P->tspan[0];
            it->grid[1] = P->tspan[1];
            it->idx = NULL;
        }
        /* InitialConditionRemaker (:284-295) */
        int tun[CO_MAX_NX], ntun = 0;
        it->n_const = 0;
        if (!it->shooting) {
            for (int k = 1; k <= L->nx; k++)
                if (!in_list(L->tic, L->n_tic, k)) it->ic_constants[it->n_const++] = k;
            for (int k = 0; k < L->n_tic; k++) tun[ntun++] = L->tic[k];
        } else {
            for (int k = 0; k < L->nquad; k++) it->ic_constants[it->n_const++] = L->quad[k];
            for (int k = 1; k <= L->nx; k++)
                if (!in_list(L->quad, L->nquad, k)) tun[ntun++] = k;
        }
        sortperm_cat(it->ic_constants, it->n_const, tun, ntun, it->ic_sorting);
        it->n_u0 = ntun;
        /* local controls (collect_local_controls, local_controls.jl:201-207) of EVERY control of the layer: controls that
         * do not act on the dynamics are filtered out of the index grid only (single_shooting.jl:301-304), ps.controls
         * keeps their values -- the measurement weights of the discrete OED layers.  The index grid of the active ones
         * addresses this vector by the offsets of the FILTERED list, as the reference does. */
        it->n_c = 0;
        for (int r = 0; r < P->ncontrols; r++)
            for (int k = 0; k < P->ctrl_nt[r]; k++)
                if (it->t0 <= P->ctrl_t[r][k] && P->ctrl_t[r][k] < it->t1) it->n_c++;
        it->c_init = (double *)malloc(sizeof(double) * (size_t)(it->n_c > 0 ? it->n_c : 1));
        int q = 0;
        for (int r = 0; r < P->ncontrols; r++)
            for (int k = 0; k < P->ctrl_nt[r]; k++)
                if (it->t0 <= P->ctrl_t[r][k] && P->ctrl_t[r][k] < it->t1)
                    it->c_init[q++] = (P->ctrl_u && P->ctrl_u[r]) ? P->ctrl_u[r][k] : 0.0;
        /* saveat (problem.kwargs; lib/CorleoneOED/src/oed.jl:106-113): every piece is a fresh solve that also saves at
         * the saveat times strictly inside its span (OrdinaryDiffEq keeps t0 < t < tf; the ends follow save_start /
         * save_end) */
        it->sv_n = (int *)calloc((size_t)it->M, sizeof(int));
        it->sv_off = (int *)calloc((size_t)it->M, sizeof(int));
        it->sv_t = (double *)malloc(sizeof(double) * (size_t)(P->n_saveat > 0 ? P->n_saveat : 1));
        it->NS = 0;
        int nsv = 0;
        for (int j = 0; j < it->M; j++) {
            it->sv_off[j] = nsv;
            for (int k = 0; k < P->n_saveat; k++)
                if (P->saveat[k] > it->grid[j] && P->saveat[k] < it->grid[j + 1]) {
                    it->sv_t[nsv++] = P->saveat[k];
                    it->sv_n[j]++;
                }
            it->NS += it->sv_n[j] + 1;
        }
        /* shooting_indices (:333-340): non-quadrature states; controls with no grid point at the interval start */
        it->n_sidx = 0;
        if (it->shooting) {
            for (int k = 1; k <= L->nx; k++)
                if (!in_list(L->quad, L->nquad, k)) it->sidx[it->n_sidx++] = k;
            for (int r = 0; r < L->nact; r++) {
                int found = 0; /* find_shooting_indices: any(first(tspan) .== control.t), local_controls.jl:180 */
                for (int k = 0; k < act_nt[r]; k++)
                    if (act_t[r][k] == it->grid[0]) found = 1;
                if (!found) it->sidx[it->n_sidx++] = L->nx + r + 1;
            }
        }
        it->var_off = off;
        off += it->n_u0 + L->ntp + it->n_c;
    }
    L->nvars = off;
    L->ndef = 0;
    L->nsamples = 1;
    for (int i = 0; i < L->I; i++) {
        L->nsamples += L->iv[i].NS;
        if (i > 0) L->ndef += L->iv[i].n_sidx + L->ntp;
    }
    free(pts);
    return L;
}

void co_layer_destroy(co_layer *L)
{
    if (!L) return;
    for (int i = 0; i < L->I; i++) {
        free(L->iv[i].grid);
        free(L->iv[i].idx);
        free(L->iv[i].c_init);
        free(L->iv[i].sv_n);
        free(L->iv[i].sv_off);
        free(L->iv[i].sv_t);
    }
    free(L->iv);
    free(L);
}

int co_n_intervals(const co_layer *L) { return L->I; }
int co_nvars(const co_layer *L) { return L->nvars; }
int co_nrow(const co_layer *L) { return L->nx + L->nact; }
int co_n_defects(const co_layer *L) { return L->ndef; }
int co_n_samples(const co_layer *L) { return L->nsamples; }

void co_block_structure(const co_layer *L, int *blocks)
{
    blocks[0] = 0;
    for (int i = 0; i < L->I; i++) blocks[i + 1] = L->iv[i].var_off + L->iv[i].n_u0 + L->ntp + L->iv[i].n_c;
}

void co_interval_info(const co_layer *L, int i, double *ts, int *M, int *n_u0, int *n_p, int *n_c, int *n_sidx)
{
    const interval *it = &L->iv[i];
    ts[0] = it->t0;
    ts[1] = it->t1;
    *M = it->M;
    *n_u0 = it->n_u0;
    *n_p = L->ntp;
    *n_c = it->n_c;
    *n_sidx = it->n_sidx;
}

void co_interval_tables(const co_layer *L, int i, int64_t *idx, double *grid, int *sidx, int *sorting,
                        int *constants, int *n_constants)
{
    const interval *it = &L->iv[i];
    if (idx && it->idx) memcpy(idx, it->idx, sizeof(int64_t) * (size_t)(L->nact * it->M));
    if (grid) memcpy(grid, it->grid, sizeof(double) * (size_t)(it->M + 1));
    if (sidx) memcpy(sidx, it->sidx, sizeof(int) * (size_t)it->n_sidx);
    if (sorting) memcpy(sorting, it->ic_sorting, sizeof(int) * (size_t)L->nx);
    if (constants) memcpy(constants, it->ic_constants, sizeof(int) * (size_t)it->n_const);
    if (n_constants) *n_constants = it->n_const;
}

/* default_initialization (multiple_shooting.jl:45-55) -> __initialparameters (single_shooting.jl:170-201);
 * bounds are infinite here, so the clamps of default_u0/default_p0 are the identity */
void co_default_ps(const co_layer *L, double *ps)
{
    for (int i = 0; i < L->I; i++) {
        const interval *it = &L->iv[i];
        double *v = ps + it->var_off;
        if (!it->shooting)
            for (int k = 0; k < L->n_tic; k++) *v++ = L->u0[L->tic[k] - 1];
        else
            for (int k = 1; k <= L->nx; k++)
                if (!in_list(L->quad, L->nquad, k)) *v++ = L->u0[k - 1];
        for (int k = 0; k < L->ntp; k++) *v++ = L->p0[L->tunable_p[k] - 1];
        for (int k = 0; k < it->n_c; k++) *v++ = it->c_init[k];
    }
}

/* one interval: (layer::SingleShootingLayer)(::Any, ps, st, shooting_layer) (single_shooting.jl:357-374)
 * followed by _sequential_solve (:421-471).  samples: (NS+1) x nrow, sample-major (NS = M without saveat). */
static void solve_interval(const co_layer *L, int i, const dual *vars, int method, int K, dual *samples,
                           double *tsamp, int shooting_flag)
{
    const interval *it = &L->iv[i];
    const int nx = L->nx, nrow = nx + L->nact;
    const dual *v_u0 = vars, *v_p = vars + it->n_u0, *v_c = v_p + L->ntp;
    dual x[CO_MAX_NX], cat[2 * CO_MAX_NX], pfull[CO_MAX_NP], ctrl[CO_MAX_CTRL];
    /* u0 = initial_condition(ps.u0, fixval); fixval = zeros for shooting layers (:359) */
    if (it->n_u0 == 0) {
        for (int k = 0; k < nx; k++) x[k] = dconst(shooting_flag ? 0.0 : L->u0[k]);
    } else {
        for (int k = 0; k < it->n_const; k++)
            cat[k] = dconst(shooting_flag ? 0.0 : L->u0[it->ic_constants[k] - 1]);
        for (int k = 0; k < it->n_u0; k++) cat[it->n_const + k] = v_u0[k];
        for (int k = 0; k < nx; k++) x[k] = cat[it->ic_sorting[k] - 1];
    }
    int ns = 0; /* samples written after the start sample */
    for (int j = 0; j < it->M; j++) {
        for (int r = 0; r < L->nact; r++) ctrl[r] = v_c[it->idx[r + (size_t)L->nact * j] - 1]; /* :439 */
        for (int q = 0; q < L->np_all; q++)
            pfull[q] = L->pmap_kind[q] ? ctrl[L->pmap_idx[q]] : v_p[L->pmap_idx[q]]; /* A*p + B*controls */
        if (j == 0) { /* save_start = (j == 1) */
            for (int k = 0; k < nx; k++) samples[k] = x[k];
            for (int r = 0; r < L->nact; r++) samples[nx + r] = ctrl[r];
            if (tsamp) tsamp[0] = it->grid[0];
        }
        sv_state sv;
        dual *svbuf = it->sv_n[j] > 0 ? (dual *)malloc(sizeof(dual) * (size_t)it->sv_n[j] * (size_t)nx) : NULL;
        sv.n = it->sv_n[j];
        sv.done = 0;
        sv.t = it->sv_t + it->sv_off[j];
        sv.out = svbuf;
        integrate_piece(&L->m, nx, x, pfull, it->grid[j], it->grid[j + 1], K, method, sv.n > 0 ? &sv : NULL);
        for (int q = 0; q < it->sv_n[j]; q++) { /* saveat samples of the piece: u = vcat(x(t), p_j) */
            dual *s = samples + (size_t)(ns + 1 + q) * nrow;
            for (int k = 0; k < nx; k++) s[k] = svbuf[(size_t)q * nx + k];
            for (int r = 0; r < L->nact; r++) s[nx + r] = ctrl[r];
            if (tsamp) tsamp[ns + 1 + q] = sv.t[q];
        }
        free(svbuf);
        ns += it->sv_n[j] + 1;
        dual *s = samples + (size_t)ns * nrow; /* save_end; u = vcat(x, p_j) (:456) */
        for (int k = 0; k < nx; k++) s[k] = x[k];
        for (int r = 0; r < L->nact; r++) s[nx + r] = ctrl[r];
        if (tsamp) tsamp[ns] = it->grid[j + 1];
    }
}

int co_eval(const co_layer *L, const double *ps, const double *dps, int nd, int method, int K, double *t_out,
            double *u_out, double *xend, double *dxend, double *defk, double *ddefk, double *defs,
            double *ddefs, double *last, double *dlast)
{
    if (nd > CO_ND) return -1;
    g_nd = nd;
    const int nx = L->nx, nrow = nx + L->nact, I = L->I;
    int maxM = 0, maxv = 0;
    for (int i = 0; i < I; i++) {
        if (L->iv[i].NS > maxM) maxM = L->iv[i].NS;
        int nv = L->iv[i].n_u0 + L->ntp + L->iv[i].n_c;
        if (nv > maxv) maxv = nv;
    }
    dual **S = (dual **)malloc(sizeof(dual *) * (size_t)I);
    double *ts = (double *)malloc(sizeof(double) * (size_t)(maxM + 1));
    dual *vars = (dual *)malloc(sizeof(dual) * (size_t)(maxv > 0 ? maxv : 1));
    int so = 0;
    for (int i = 0; i < I; i++) {
        const interval *it = &L->iv[i];
        int nv = it->n_u0 + L->ntp + it->n_c;
        for (int k = 0; k < nv; k++) {
            vars[k].v = ps[it->var_off + k];
            for (int d = 0; d < nd; d++) vars[k].d[d] = dps[(size_t)d * L->nvars + it->var_off + k];
        }
        S[i] = (dual *)malloc(sizeof(dual) * (size_t)((it->NS + 1) * nrow));
        solve_interval(L, i, vars, method, K, S[i], ts, it->shooting);
        if (t_out) /* drop the duplicated boundary point of all but the last interval (:145-147) */
            for (int j = 0; j < it->NS + (i == I - 1); j++) t_out[so + j] = ts[j];
        if (xend)
            for (int k = 0; k < nx; k++) {
                xend[k + nx * i] = S[i][(size_t)it->NS * nrow + k].v;
                if (dxend)
                    for (int d = 0; d < nd; d++)
                        dxend[k + nx * i + (size_t)d * nx * I] = S[i][(size_t)it->NS * nrow + k].d[d];
            }
        so += it->NS;
    }
    /* shooting defects (multiple_shooting.jl:150-163) */
    if (defk || defs) {
        int n_u = 0, n_p = 0;
        for (int i = 1; i < I; i++) {
            for (int q = 0; q < L->iv[i].n_sidx; q++)
                if (L->iv[i].sidx[q] <= nx) n_u++;
            n_p += L->ntp;
        }
        int ou = 0, op = n_u, oc = n_u + n_p, os = 0;
        for (int i = 1; i < I; i++) {
            const interval *prev = &L->iv[i - 1], *it = &L->iv[i];
            const dual *lastp = S[i - 1] + (size_t)prev->NS * nrow, *first = S[i];
            /* u0 matchings, then p, then controls within the stage */
            for (int pass = 0; pass < 3; pass++) {
                if (pass == 1) {
                    for (int k = 0; k < L->ntp; k++) {
                        double v = ps[prev->var_off + prev->n_u0 + k] - ps[it->var_off + it->n_u0 + k];
                        if (defk) defk[op] = v;
                        if (defs) defs[os] = v;
                        for (int d = 0; d < nd; d++) {
                            double dv = dps[(size_t)d * L->nvars + prev->var_off + prev->n_u0 + k] -
                                        dps[(size_t)d * L->nvars + it->var_off + it->n_u0 + k];
                            if (ddefk) ddefk[op + (size_t)d * L->ndef] = dv;
                            if (ddefs) ddefs[os + (size_t)d * L->ndef] = dv;
                        }
                        op++;
                        os++;
                    }
                    continue;
                }
                for (int q = 0; q < it->n_sidx; q++) {
                    int r = it->sidx[q] - 1;
                    if ((pass == 0) != (r < nx)) continue;
                    dual df = dsub(lastp[r], first[r]);
                    int *o = (pass == 0) ? &ou : &oc;
                    if (defk) defk[*o] = df.v;
                    if (defs) defs[os] = df.v;
                    for (int d = 0; d < nd; d++) {
                        if (ddefk) ddefk[*o + (size_t)d * L->ndef] = df.d[d];
                        if (ddefs) ddefs[os + (size_t)d * L->ndef] = df.d[d];
                    }
                    (*o)++;
                    os++;
                }
            }
        }
    }
    /* quadrature accumulation (:171-177), then concatenation (:178-180) */
    for (int i = 1; i < I; i++) {
        const interval *prev = &L->iv[i - 1], *it = &L->iv[i];
        for (int q = 0; q < L->nquad; q++) {
            int r = L->quad[q] - 1;
            dual qp = S[i - 1][(size_t)prev->NS * nrow + r];
            for (int j = 0; j <= it->NS; j++) S[i][(size_t)j * nrow + r] = dadd(S[i][(size_t)j * nrow + r], qp);
        }
    }
    so = 0;
    for (int i = 0; i < I; i++) {
        const interval *it = &L->iv[i];
        if (u_out)
            for (int j = 0; j < it->NS + (i == I - 1); j++)
                for (int r = 0; r < nrow; r++) u_out[r + (size_t)nrow * (so + j)] = S[i][(size_t)j * nrow + r].v;
        so += it->NS;
    }
    if (last) {
        const dual *l = S[I - 1] + (size_t)L->iv[I - 1].NS * nrow;
        for (int r = 0; r < nrow; r++) {
            last[r] = l[r].v;
            if (dlast)
                for (int d = 0; d < nd; d++) dlast[r + (size_t)d * nrow] = l[r].d[d];
        }
    }
    for (int i = 0; i < I; i++) free(S[i]);
    free(S);
    free(ts);
    free(vars);
    g_nd = 0;
    return 0;
}

/* Tangents of the merged trajectory samples (what ForwardDiff through `traj` yields; the point constraints of
 * src/dynprob.jl:30-58,89-103 differentiate these): du_out[r + nrow*(sample + nsamples*d)], quadrature rows carry
 * the accumulated sums of multiple_shooting.jl:171-177 and therefore depend on earlier intervals' variables too. */
int co_eval_traj(const co_layer *L, const double *ps, const double *dps, int nd, int method, int K, double *u_out,
                 double *du_out)
{
    if (nd > CO_ND) return -1;
    g_nd = nd;
    const int nx = L->nx, nrow = nx + L->nact, I = L->I;
    int maxM = 0, maxv = 0, nsamples = 1;
    for (int i = 0; i < I; i++) {
        if (L->iv[i].NS > maxM) maxM = L->iv[i].NS;
        int nv = L->iv[i].n_u0 + L->ntp + L->iv[i].n_c;
        if (nv > maxv) maxv = nv;
        nsamples += L->iv[i].NS;
    }
    dual **S = (dual **)malloc(sizeof(dual *) * (size_t)I);
    double *ts = (double *)malloc(sizeof(double) * (size_t)(maxM + 1));
    dual *vars = (dual *)malloc(sizeof(dual) * (size_t)(maxv > 0 ? maxv : 1));
    for (int i = 0; i < I; i++) {
        const interval *it = &L->iv[i];
        int nv = it->n_u0 + L->ntp + it->n_c;
        for (int k = 0; k < nv; k++) {
            vars[k].v = ps[it->var_off + k];
            for (int d = 0; d < nd; d++) vars[k].d[d] = dps[(size_t)d * L->nvars + it->var_off + k];
        }
        S[i] = (dual *)malloc(sizeof(dual) * (size_t)((it->NS + 1) * nrow));
        solve_interval(L, i, vars, method, K, S[i], ts, it->shooting);
    }
    for (int i = 1; i < I; i++) {
        const interval *prev = &L->iv[i - 1], *it = &L->iv[i];
        for (int q = 0; q < L->nquad; q++) {
            int r = L->quad[q] - 1;
            dual qp = S[i - 1][(size_t)prev->NS * nrow + r];
            for (int j = 0; j <= it->NS; j++) S[i][(size_t)j * nrow + r] = dadd(S[i][(size_t)j * nrow + r], qp);
        }
    }
    int so = 0;
    for (int i = 0; i < I; i++) {
        const interval *it = &L->iv[i];
        for (int j = 0; j < it->NS + (i == I - 1); j++)
            for (int r = 0; r < nrow; r++) {
                if (u_out) u_out[r + (size_t)nrow * (so + j)] = S[i][(size_t)j * nrow + r].v;
                if (du_out)
                    for (int d = 0; d < nd; d++)
                        du_out[r + (size_t)nrow * ((so + j) + (size_t)nsamples * d)] = S[i][(size_t)j * nrow + r].d[d];
            }
        so += it->NS;
    }
    for (int i = 0; i < I; i++) free(S[i]);
    free(S);
    free(ts);
    free(vars);
    g_nd = 0;
    return 0;
}

int co_forward_init(const co_layer *L, double *ps, const int *fixed, int nfixed, int method, int K)
{
    g_nd = 0;
    const int nx = L->nx, nrow = nx + L->nact;
    double carry[CO_MAX_NX + CO_MAX_CTRL];
    int ncarry = L->iv[0].n_u0; /* u0s = [first(ps).u0] */
    for (int k = 0; k < ncarry; k++) carry[k] = ps[L->iv[0].var_off + k];
    for (int i = 0; i < L->I; i++) {
        const interval *it = &L->iv[i];
        int nv = it->n_u0 + L->ntp + it->n_c;
        int is_fixed = in_list(fixed, nfixed, i + 1);
        if (!is_fixed) {
            if (ncarry < it->n_u0) return -2; /* BoundsError in the reference */
            for (int k = 0; k < it->n_u0; k++) ps[it->var_off + k] = carry[k]; /* u0[eachindex(sps.u0)] */
        }
        dual *vars = (dual *)malloc(sizeof(dual) * (size_t)(nv > 0 ? nv : 1));
        for (int k = 0; k < nv; k++) vars[k] = dconst(ps[it->var_off + k]);
        dual *S = (dual *)malloc(sizeof(dual) * (size_t)((it->NS + 1) * nrow));
        /* layer(nothing, sps, sst): the shooting flag is NOT passed here, so fixval = problem.u0 (:165) */
        solve_interval(L, i, vars, method, K, S, NULL, 0);
        for (int r = 0; r < nrow; r++) carry[r] = S[(size_t)it->NS * nrow + r].v; /* pred.u[end] */
        ncarry = nrow;
        free(S);
        free(vars);
    }
    return 0;
}

int co_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double co_bench_units(const co_layer *L, const double *ps_batch, int64_t B, int64_t b0, int64_t b1, int method,
                      int K, int with_sens, int nthreads, double *checksum)
{
    (void)B;
    const int I = L->I, nx = L->nx, nrow = nx + L->nact;
    int maxM = 0, maxv = 0;
    for (int i = 0; i < I; i++) {
        if (L->iv[i].NS > maxM) maxM = L->iv[i].NS;
        int nv = L->iv[i].n_u0 + L->ntp + L->iv[i].n_c;
        if (nv > maxv) maxv = nv;
    }
    double sum = 0.0;
    double t0 = now_s();
    int64_t nunits = (b1 - b0) * I;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel reduction(+ : sum)
    {
        dual *vars = (dual *)malloc(sizeof(dual) * (size_t)(maxv > 0 ? maxv : 1));
        dual *S = (dual *)malloc(sizeof(dual) * (size_t)((maxM + 1) * nrow));
#pragma omp for schedule(static)
        for (int64_t u = 0; u < nunits; u++) {
            int64_t b = b0 + u / I;
            int i = (int)(u % I);
            const interval *it = &L->iv[i];
            int nv = it->n_u0 + L->ntp + it->n_c;
            const double *psb = ps_batch + (size_t)b * L->nvars + it->var_off;
            int npass = with_sens ? (nv + CO_ND - 1) / CO_ND : 1;
            if (npass < 1) npass = 1;
            for (int pass = 0; pass < npass; pass++) {
                int d0 = pass * CO_ND, nd = with_sens ? (nv - d0 < CO_ND ? nv - d0 : CO_ND) : 0;
                if (nd < 0) nd = 0;
                g_nd = nd;
                for (int k = 0; k < nv; k++) {
                    vars[k].v = psb[k];
                    for (int d = 0; d < nd; d++) vars[k].d[d] = (k == d0 + d) ? 1.0 : 0.0;
                }
                solve_interval(L, i, vars, method, K, S, NULL, it->shooting);
                const dual *e = S + (size_t)it->NS * nrow;
                for (int k = 0; k < nx; k++) {
                    sum += e[k].v;
                    for (int d = 0; d < nd; d++) sum += e[k].d[d];
                }
            }
        }
        free(vars);
        free(S);
        g_nd = 0;
    }
    double t1 = now_s();
    if (checksum) *checksum = sum;
    return t1 - t0;
}
