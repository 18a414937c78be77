// Device model registry: right-hand sides and their directional derivatives.
//
// The reference integrates arbitrary Julia closures (src/single_shooting.jl:443-453) and differentiates
// them with ForwardDiff; a C-ABI CUDA library needs the RHS compiled in.  Every model gives
//   f(x, p, dx)                      dx = f(x, p)
//   jvp(x, p, v, pcol, dv)           dv = f_x(x, p) v + (pcol >= 0 ? f_p(x, p)[:, pcol] : 0)
// where p is the FULL problem.p vector (controls scattered in, src/single_shooting.jl:306-323) and pcol is
// a 0-based column of it.  All loops are unrolled and all arrays live in registers.
#pragma once
#include <cuda_runtime.h>

namespace cb200 {

#define CB_DEV static __device__ __forceinline__

// model constants for DENSE20 (W row-major 20x20, then b[20]); one copy per translation unit that uses it
#ifdef CB200_NEED_DENSE20_DATA
__constant__ double c_dense20[420];
#endif

// ---- Lotka-Volterra fishing with Lagrange state (examples/lotka_fishing_optimal_control/main.jl:38-42)
struct Lotka3 {
    static constexpr int NX = 3, NP = 3;
    static constexpr bool HAS_JVP = true;
    CB_DEV void f(const double *x, const double *p, double *dx)
    {
        const double x12 = x[0] * x[1];
        dx[0] = x[0] - p[1] * x12 - 0.4 * p[0] * x[0];
        dx[1] = -x[1] + p[2] * x12 - 0.2 * p[0] * x[1];
        const double a = x[0] - 1.0, b = x[1] - 1.0;
        dx[2] = a * a + b * b;
    }
    CB_DEV void jvp(const double *x, const double *p, const double *v, int pcol, double *dv)
    {
        const double j00 = 1.0 - p[1] * x[1] - 0.4 * p[0], j01 = -p[1] * x[0];
        const double j10 = p[2] * x[1], j11 = -1.0 + p[2] * x[0] - 0.2 * p[0];
        double d0 = j00 * v[0] + j01 * v[1];
        double d1 = j10 * v[0] + j11 * v[1];
        const double d2 = 2.0 * (x[0] - 1.0) * v[0] + 2.0 * (x[1] - 1.0) * v[1];
        if (pcol == 0) {
            d0 -= 0.4 * x[0];
            d1 -= 0.2 * x[1];
        } else if (pcol == 1) {
            d0 -= x[0] * x[1];
        } else if (pcol == 2) {
            d1 += x[0] * x[1];
        }
        dv[0] = d0;
        dv[1] = d1;
        dv[2] = d2;
    }
};

// ---- Lotka-Volterra without the Lagrange state (lib/CorleoneOED/examples/lotka_oed/main.jl:61-65)
struct Lotka2 {
    static constexpr int NX = 2, NP = 3;
    static constexpr bool HAS_JVP = true;
    CB_DEV void f(const double *x, const double *p, double *dx)
    {
        const double x12 = x[0] * x[1];
        dx[0] = x[0] - p[1] * x12 - 0.4 * p[0] * x[0];
        dx[1] = -x[1] + p[2] * x12 - 0.2 * p[0] * x[1];
    }
    CB_DEV void jvp(const double *x, const double *p, const double *v, int pcol, double *dv)
    {
        const double j00 = 1.0 - p[1] * x[1] - 0.4 * p[0], j01 = -p[1] * x[0];
        const double j10 = p[2] * x[1], j11 = -1.0 + p[2] * x[0] - 0.2 * p[0];
        double d0 = j00 * v[0] + j01 * v[1];
        double d1 = j10 * v[0] + j11 * v[1];
        if (pcol == 0) {
            d0 -= 0.4 * x[0];
            d1 -= 0.2 * x[1];
        } else if (pcol == 1) {
            d0 -= x[0] * x[1];
        } else if (pcol == 2) {
            d1 += x[0] * x[1];
        }
        dv[0] = d0;
        dv[1] = d1;
    }
};

// ---- linear-quadratic regulator (examples/linear_quadratic/main.jl:45-50), p = (a, b, u)
struct Lqr {
    static constexpr int NX = 2, NP = 3;
    static constexpr bool HAS_JVP = true;
    CB_DEV void f(const double *x, const double *p, double *dx)
    {
        dx[0] = p[0] * x[0] + p[1] * p[2];
        const double e = x[0] - 3.0;
        dx[1] = 10.0 * e * e + 0.1 * p[2] * p[2];
    }
    CB_DEV void jvp(const double *x, const double *p, const double *v, int pcol, double *dv)
    {
        double d0 = p[0] * v[0];
        double d1 = 20.0 * (x[0] - 3.0) * v[0];
        if (pcol == 0) {
            d0 += x[0];
        } else if (pcol == 1) {
            d0 += p[2];
        } else if (pcol == 2) {
            d0 += p[1];
            d1 += 0.2 * p[2];
        }
        dv[0] = d0;
        dv[1] = d1;
    }
};

// ---- 1-D linear problem (lib/CorleoneOED/test/core/1d_oed.jl:11-13)
struct Lin1d {
    static constexpr int NX = 1, NP = 1;
    static constexpr bool HAS_JVP = true;
    CB_DEV void f(const double *x, const double *p, double *dx) { dx[0] = p[0] * x[0]; }
    CB_DEV void jvp(const double *x, const double *p, const double *v, int pcol, double *dv)
    {
        dv[0] = p[0] * v[0] + (pcol == 0 ? x[0] : 0.0);
    }
};

// ---- Egerstedt switching problem (test/core/local_controls.jl:46-52)
struct Egerstedt {
    static constexpr int NX = 3, NP = 3;
    static constexpr bool HAS_JVP = true;
    CB_DEV void f(const double *u, const double *p, double *du)
    {
        const double x = u[0], y = u[1];
        du[0] = -x * p[0] + (x + y) * p[1] + (x - y) * p[2];
        du[1] = (x + 2.0 * y) * p[0] + (x - 2.0 * y) * p[1] + (x + y) * p[2];
        du[2] = x * x + y * y;
    }
    CB_DEV void jvp(const double *u, const double *p, const double *v, int pcol, double *dv)
    {
        const double x = u[0], y = u[1];
        double d0 = (-p[0] + p[1] + p[2]) * v[0] + (p[1] - p[2]) * v[1];
        double d1 = (p[0] + p[1] + p[2]) * v[0] + (2.0 * p[0] - 2.0 * p[1] + p[2]) * v[1];
        const double d2 = 2.0 * x * v[0] + 2.0 * y * v[1];
        if (pcol == 0) {
            d0 -= x;
            d1 += x + 2.0 * y;
        } else if (pcol == 1) {
            d0 += x + y;
            d1 += x - 2.0 * y;
        } else if (pcol == 2) {
            d0 += x - y;
            d1 += x + y;
        }
        dv[0] = d0;
        dv[1] = d1;
        dv[2] = d2;
    }
};

#ifdef CB200_NEED_DENSE20_DATA
// ---- synthetic 20-state problem of BASELINE config 4: x' = W x - x.^3 + b u, p = (u)
struct Dense20 {
    static constexpr int NX = 20, NP = 1;
    static constexpr bool HAS_JVP = true;
    CB_DEV void f(const double *x, const double *p, double *dx)
    {
#pragma unroll
        for (int i = 0; i < 20; i++) {
            double acc = c_dense20[400 + i] * p[0];
#pragma unroll
            for (int j = 0; j < 20; j++) acc = fma(c_dense20[i * 20 + j], x[j], acc);
            dx[i] = acc - x[i] * x[i] * x[i];
        }
    }
    CB_DEV void jvp(const double *x, const double *p, const double *v, int pcol, double *dv)
    {
#pragma unroll
        for (int i = 0; i < 20; i++) {
            double acc = (pcol == 0) ? c_dense20[400 + i] : 0.0;
#pragma unroll
            for (int j = 0; j < 20; j++) acc = fma(c_dense20[i * 20 + j], v[j], acc);
            dv[i] = fma(-3.0 * x[i] * x[i], v[i], acc);
        }
    }
};
#endif

// ---- OED augmentation (lib/CorleoneOED/src/augmentation.jl), ContinuousSampled non-fixed variant:
//   state  [x; vec(G) column-major nx x NF; F[triu] column-major]     (:239, selector :211)
//   params [base p; w_1..w_NOBS]                                       (:228-230)
//   G' = f_x G + f_p[:, params]  (:147-154),  F' = sum_i w_i (h_x[i,:] G)'(h_x[i,:] G)  (:222-233)
// observed = the leading NOBS states, so h_x rows are unit rows.  Traits::PCOL(c) gives the 0-based p
// column of Fisher parameter c.
template <class Base, class Traits>
struct OedAug {
    static constexpr int BX = Base::NX, NF = Traits::NF, NOBS = Traits::NOBS;
    static constexpr int NX = BX + BX * NF + NF * (NF + 1) / 2, NP = Base::NP + NOBS;
    static constexpr bool HAS_JVP = false;
    CB_DEV void f(const double *y, const double *p, double *dy)
    {
        Base::f(y, p, dy);
#pragma unroll
        for (int c = 0; c < NF; c++) Base::jvp(y, p, y + BX + BX * c, Traits::pcol(c), dy + BX + BX * c);
        const double *G = y + BX;
        const double *w = p + Base::NP;
        double *dF = dy + BX + BX * NF;
        int q = 0;
#pragma unroll
        for (int b = 0; b < NF; b++)
#pragma unroll
            for (int a = 0; a <= b; a++) {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < NOBS; i++) acc += w[i] * (G[i + BX * a] * G[i + BX * b]);
                dF[q++] = acc;
            }
    }
    CB_DEV void jvp(const double *, const double *, const double *, int, double *) {}
};
struct LotkaOedTraits {
    static constexpr int NF = 2, NOBS = 2;
    static __device__ __forceinline__ constexpr int pcol(int c) { return c + 1; }  // params = [2, 3]
};
struct Lin1dOedTraits {
    static constexpr int NF = 1, NOBS = 1;
    static __device__ __forceinline__ constexpr int pcol(int) { return 0; }  // params = [1]
};
using Lotka2Oed = OedAug<Lotka2, LotkaOedTraits>;
using Lin1dOed = OedAug<Lin1d, Lin1dOedTraits>;

}  // namespace cb200
