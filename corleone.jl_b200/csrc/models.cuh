// Device model registry: right-hand sides and their directional derivatives.
//
// The reference integrates arbitrary Julia closures (src/single_shooting.jl:443-453) and differentiates
// them with ForwardDiff; a C-ABI CUDA library needs the RHS compiled in.  State vectors are ordered
// [NXD dynamic states; NQ quadrature states]: quadrature states (Lagrange terms, Fisher-information
// entries) never enter a right-hand side, so the integrator needs no stage storage for them.
//
// Every model gives
//   f(x, p, dx, ctx)                 dx[0..NX) = f(x, p); x holds the NXD dynamic states only;
//                                    ctx keeps what the Jacobian-vector products of the same stage reuse
//   make_force(pcol, p)              per control piece: the p-direction of a sensitivity column
//                                    (pcol = 0-based column of the FULL problem.p, -1 = none)
//   jvp(x, p, ctx, v, F, dv)         dv[0..NX) = f_x(x, p) v + f_p(x, p)[:, pcol]; v holds NXD entries
// p is the full parameter vector with the controls scattered in (src/single_shooting.jl:306-323).
// The p-direction is applied through precomputed weights (Force), not branches: the step body is one
// straight-line block of FP64 instructions.
#pragma once
#include <cuda_runtime.h>

namespace cb200 {

#define CB_DEV static __device__ __forceinline__

// FimTraits<M>::N > 0: model M carries an n x n Fisher block F_triu at state offset OFF whose end value the shooting
// kernel reads out itself (OedAug below); 0 for every other model
template <class M, class = void>
struct FimTraits {
    static constexpr int N = 0, OFF = 0;
};
template <class M>
struct FimTraits<M, decltype((void)M::FIM_N, void())> {
    static constexpr int N = M::FIM_N, OFF = M::FIM_OFF;
};

// SaveAtTraits<M>::ON: the shooting kernel is also instantiated with dense-output sampling at saveat times inside control
// pieces (cb200_desc.saveat) for model M -- models that declare `static constexpr bool SAVEAT = true`
template <class M, class = void>
struct SaveAtTraits {
    static constexpr bool ON = false;
};
template <class M>
struct SaveAtTraits<M, decltype((void)M::SAVEAT, void())> {
    static constexpr bool ON = M::SAVEAT;
};

// SmallTailTraits<M>::ON: the shooting kernel is also built with the small-evaluation tail (one-kernel evaluations of a few
// candidates) -- models that declare `static constexpr bool SMALL_TAIL = true`
template <class M, class = void>
struct SmallTailTraits {
    static constexpr bool ON = false;
};
template <class M>
struct SmallTailTraits<M, decltype((void)M::SMALL_TAIL, void())> {
    static constexpr bool ON = M::SMALL_TAIL;
};

#ifdef CB200_NEED_DENSE20_DATA
// model constants of the dense family (W row-major NX x NX, then b[NX]; NX <= 32) for the thread-per-unit kernels of the
// including translation unit: constant-bank operands of the FMAs (one copy per translation unit = per NX)
static __constant__ double c_dense[32 * 32 + 32];
#endif

// ---- Lotka-Volterra fishing with Lagrange state (examples/lotka_fishing_optimal_control/main.jl:38-42)
struct Lotka3 {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 3;
    static constexpr bool SMALL_TAIL = true;   // single-candidate evaluations are one kernel (shoot.cuh small_tail)
    static constexpr bool SAVEAT = true;   // also built with dense-output sampling (cb200_desc.saveat)
    static constexpr bool HAS_JVP = true;
    struct Ctx {
        double j00, j01, j10, j11, x12, q0, q1;
    };
    struct Force {
        double c00, c01, c10, c11;  // d0 += c00 x0 + c01 x0 x1 ; d1 += c10 x1 + c11 x0 x1
    };
    CB_DEV void f(const double *x, const double *p, double *dx, Ctx &c)
    {
        c.x12 = x[0] * x[1];
        c.j00 = 1.0 - p[1] * x[1] - 0.4 * p[0];
        c.j01 = -p[1] * x[0];
        c.j10 = p[2] * x[1];
        c.j11 = -1.0 + p[2] * x[0] - 0.2 * p[0];
        dx[0] = x[0] - p[1] * c.x12 - 0.4 * p[0] * x[0];
        dx[1] = -x[1] + p[2] * c.x12 - 0.2 * p[0] * x[1];
        const double a = x[0] - 1.0, b = x[1] - 1.0;
        c.q0 = 2.0 * a;
        c.q1 = 2.0 * b;
        dx[2] = a * a + b * b;
    }
    CB_DEV Force make_force(int pcol, const double *)
    {
        // selects, not branches: the piece prologue stays straight-line
        return Force{pcol == 0 ? -0.4 : 0.0, pcol == 1 ? -1.0 : 0.0, pcol == 0 ? -0.2 : 0.0, pcol == 2 ? 1.0 : 0.0};
    }
    CB_DEV void jvp(const double *x, const double *, const Ctx &c, const double *v, const Force &F, double *dv)
    {
        dv[0] = fma(c.j00, v[0], fma(c.j01, v[1], fma(F.c00, x[0], F.c01 * c.x12)));
        dv[1] = fma(c.j10, v[0], fma(c.j11, v[1], fma(F.c10, x[1], F.c11 * c.x12)));
        dv[2] = fma(c.q0, v[0], c.q1 * v[1]);
    }
};

// ---- Lotka-Volterra without the Lagrange state (lib/CorleoneOED/examples/lotka_oed/main.jl:61-65)
struct Lotka2 {
    static constexpr int NX = 2, NXD = 2, NQ = 0, NP = 3;
    static constexpr bool SMALL_TAIL = true;   // single-candidate evaluations are one kernel (shoot.cuh small_tail)
    static constexpr bool SAVEAT = true;   // also built with dense-output sampling (cb200_desc.saveat)
    static constexpr bool HAS_JVP = true;
    struct Ctx {
        double j00, j01, j10, j11, x12;
    };
    using Force = Lotka3::Force;
    CB_DEV void f(const double *x, const double *p, double *dx, Ctx &c)
    {
        c.x12 = x[0] * x[1];
        c.j00 = 1.0 - p[1] * x[1] - 0.4 * p[0];
        c.j01 = -p[1] * x[0];
        c.j10 = p[2] * x[1];
        c.j11 = -1.0 + p[2] * x[0] - 0.2 * p[0];
        dx[0] = x[0] - p[1] * c.x12 - 0.4 * p[0] * x[0];
        dx[1] = -x[1] + p[2] * c.x12 - 0.2 * p[0] * x[1];
    }
    CB_DEV Force make_force(int pcol, const double *p) { return Lotka3::make_force(pcol, p); }
    CB_DEV void jvp(const double *x, const double *, const Ctx &c, const double *v, const Force &F, double *dv)
    {
        dv[0] = fma(c.j00, v[0], fma(c.j01, v[1], fma(F.c00, x[0], F.c01 * c.x12)));
        dv[1] = fma(c.j10, v[0], fma(c.j11, v[1], fma(F.c10, x[1], F.c11 * c.x12)));
    }
    // second-order term of the sensitivity equations: out = d/d eps [ f_x g + f_p[:, fpcol] ] along the tangent
    // (dx, e_pcol) -- i.e. (f_xx[dx] + f_xp[e_pcol]) g + (f_px[dx] + f_pp[e_pcol])[:, fpcol]; pcol / fpcol = -1: none
    CB_DEV void d2(const double *x, const double *p, const double *dx, int pcol, const double *g, int fpcol, double *out)
    {
        const double dp0 = pcol == 0 ? 1.0 : 0.0, dp1 = pcol == 1 ? 1.0 : 0.0, dp2 = pcol == 2 ? 1.0 : 0.0;
        const double dj00 = -p[1] * dx[1] - dp1 * x[1] - 0.4 * dp0, dj01 = -p[1] * dx[0] - dp1 * x[0];
        const double dj10 = p[2] * dx[1] + dp2 * x[1], dj11 = p[2] * dx[0] + dp2 * x[0] - 0.2 * dp0;
        const double dx12 = fma(dx[0], x[1], x[0] * dx[1]);
        out[0] = fma(dj00, g[0], dj01 * g[1]) + (fpcol == 0 ? -0.4 * dx[0] : (fpcol == 1 ? -dx12 : 0.0));
        out[1] = fma(dj10, g[0], dj11 * g[1]) + (fpcol == 0 ? -0.2 * dx[1] : (fpcol == 2 ? dx12 : 0.0));
    }
};

// ---- linear-quadratic regulator (examples/linear_quadratic/main.jl:45-50), p = (a, b, u)
struct Lqr {
    static constexpr int NX = 2, NXD = 1, NQ = 1, NP = 3;
    static constexpr bool SMALL_TAIL = true;
    static constexpr bool HAS_JVP = true;
    struct Ctx {
        double j10;
    };
    struct Force {
        double a0, b0, b1;  // d0 += a0 x0 + b0 ; d1 += b1
    };
    CB_DEV void f(const double *x, const double *p, double *dx, Ctx &c)
    {
        dx[0] = fma(p[0], x[0], p[1] * p[2]);
        const double e = x[0] - 3.0;
        c.j10 = 20.0 * e;
        dx[1] = fma(10.0 * e, e, 0.1 * (p[2] * p[2]));
    }
    CB_DEV Force make_force(int pcol, const double *p)
    {
        // d f/d a = (x, 0); d f/d b = (u, 0); d f/d u = (b, 0.2 u) -- selects, not branches
        Force F;
        F.a0 = pcol == 0 ? 1.0 : 0.0;
        F.b0 = pcol == 1 ? p[2] : 0.0;
        F.b0 = pcol == 2 ? p[1] : F.b0;
        F.b1 = pcol == 2 ? 0.2 * p[2] : 0.0;
        return F;
    }
    CB_DEV void jvp(const double *x, const double *p, const Ctx &c, const double *v, const Force &F, double *dv)
    {
        dv[0] = fma(p[0], v[0], fma(F.a0, x[0], F.b0));
        dv[1] = fma(c.j10, v[0], F.b1);
    }
};

// ---- 1-D linear problem (lib/CorleoneOED/test/core/1d_oed.jl:11-13)
struct Lin1d {
    static constexpr int NX = 1, NXD = 1, NQ = 0, NP = 1;
    static constexpr bool SAVEAT = true;   // also built with dense-output sampling (cb200_desc.saveat)
    static constexpr bool HAS_JVP = true;
    struct Ctx {};
    struct Force {
        double w;
    };
    CB_DEV void f(const double *x, const double *p, double *dx, Ctx &) { dx[0] = p[0] * x[0]; }
    CB_DEV Force make_force(int pcol, const double *) { return Force{pcol == 0 ? 1.0 : 0.0}; }
    CB_DEV void jvp(const double *x, const double *p, const Ctx &, const double *v, const Force &F, double *dv)
    {
        dv[0] = fma(p[0], v[0], F.w * x[0]);
    }
    // f = p x: f_x = p, f_p = x  ->  d(f_x g + f_p[:, fpcol]) = dp g + (fpcol == 0 ? dx : 0)
    CB_DEV void d2(const double *, const double *, const double *dx, int pcol, const double *g, int fpcol, double *out)
    {
        out[0] = (pcol == 0 ? g[0] : 0.0) + (fpcol == 0 ? dx[0] : 0.0);
    }
};

// ---- Egerstedt switching problem (test/core/local_controls.jl:46-52)
struct Egerstedt {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 3;
    static constexpr bool HAS_JVP = true;
    struct Ctx {};
    struct Force {
        double w0, w1, w2;
    };
    CB_DEV void f(const double *u, const double *p, double *du, Ctx &)
    {
        const double x = u[0], y = u[1];
        du[0] = -x * p[0] + (x + y) * p[1] + (x - y) * p[2];
        du[1] = (x + 2.0 * y) * p[0] + (x - 2.0 * y) * p[1] + (x + y) * p[2];
        du[2] = x * x + y * y;
    }
    CB_DEV Force make_force(int pcol, const double *)
    {
        return Force{pcol == 0 ? 1.0 : 0.0, pcol == 1 ? 1.0 : 0.0, pcol == 2 ? 1.0 : 0.0};
    }
    CB_DEV void jvp(const double *u, const double *p, const Ctx &, const double *v, const Force &F, double *dv)
    {
        const double x = u[0], y = u[1];
        dv[0] = (-p[0] + p[1] + p[2]) * v[0] + (p[1] - p[2]) * v[1] + F.w0 * (-x) + F.w1 * (x + y) + F.w2 * (x - y);
        dv[1] = (p[0] + p[1] + p[2]) * v[0] + (2.0 * p[0] - 2.0 * p[1] + p[2]) * v[1] + F.w0 * (x + 2.0 * y) +
                F.w1 * (x - 2.0 * y) + F.w2 * (x + y);
        dv[2] = 2.0 * x * v[0] + 2.0 * y * v[1];
    }
};

#ifdef CB200_NEED_DENSE20_DATA
// ---- synthetic dense problems: x' = W x - x.^3 + b u, p = (u).  NX = 20 is BASELINE config 4; 16, 24 and 32 are the same
//      family at the other sizes the tensor-path kernel is built for
template <int N>
struct DenseN {
    static constexpr int NX = N, NXD = N, NQ = 0, NP = 1;
    static constexpr bool HAS_JVP = true;
    struct Ctx {};
    struct Force {
        double w;
    };
    CB_DEV void f(const double *x, const double *p, double *dx, Ctx &)
    {
#pragma unroll
        for (int i = 0; i < N; i++) {
            double acc = c_dense[N * N + i] * p[0];
#pragma unroll
            for (int j = 0; j < N; j++) acc = fma(c_dense[i * N + j], x[j], acc);
            dx[i] = acc - x[i] * x[i] * x[i];
        }
    }
    CB_DEV Force make_force(int pcol, const double *) { return Force{pcol == 0 ? 1.0 : 0.0}; }
    CB_DEV void jvp(const double *x, const double *, const Ctx &, const double *v, const Force &F, double *dv)
    {
#pragma unroll
        for (int i = 0; i < N; i++) {
            double acc = F.w * c_dense[N * N + i];
#pragma unroll
            for (int j = 0; j < N; j++) acc = fma(c_dense[i * N + j], v[j], acc);
            dv[i] = fma(-3.0 * x[i] * x[i], v[i], acc);
        }
    }
};
using Dense20 = DenseN<20>;
#endif

// ---- OED augmentation (lib/CorleoneOED/src/augmentation.jl).  State [x; vec(G) column-major nx x NF; F ...],
//      G' = f_x G + f_p[:, params] (:147-154); observed = the leading NOBS states, so h_x rows are unit rows and
//      g_i = (h_x[i,:] G) = row i of G.  Traits::pcol(c) is the 0-based p column of Fisher parameter c.
//      x and G are dynamic states; the F entries are quadratures.  Traits::MODE picks the variant:
//   OED_CS  ContinuousSampled, non-fixed (:206-253 with measurements): params [base p; w_1..w_NOBS] (:228-230),
//           F[triu] column-major (:239, selector :211), F' = sum_i w_i g_i' g_i (:222-227)
//   OED_CU  Continuous without measurements (:217-220): no weights, F' = s's with s = sum_i g_i (rows summed FIRST)
//   OED_CF  ContinuousSampled, fixed (:255-292): one F_k[triu] block per observed function, F_k' = g_k' g_k; the
//           weights are appended as controls (:275-278) but do not enter the right-hand side -- they weight the
//           increments of F_k in the read-out (oed.jl:352-362)
//   OED_DS / OED_DU  Discrete with / without measurements (:171-203): no F states at all; the information is summed
//           over the saved samples (oed.jl:401-428).  DS appends the weights as controls, DU does not.
enum { OED_CS = 0, OED_CU = 1, OED_CF = 2, OED_DS = 3, OED_DU = 4 };
template <class Base, class Traits>
struct OedAug {
    static_assert(Base::NQ == 0, "OED base model must not carry quadrature states");
    static constexpr int BX = Base::NX, NF = Traits::NF, NOBS = Traits::NOBS, MODE = Traits::MODE;
    static constexpr int NTRI = NF * (NF + 1) / 2;
    // weights appended to p: the continuous variants with measurements.  The discrete ones keep the base parameters
    // (augmentation.jl:171-203): their measurement controls do not act on the dynamics, are filtered out of the index grid
    // (src/single_shooting.jl:301-304) and weight the samples taken at the measurement times (saveat) in the read-out
    static constexpr int NW = (MODE == OED_CU || MODE == OED_DU || MODE == OED_DS) ? 0 : NOBS;
    static constexpr bool SAVEAT = (MODE == OED_DS || MODE == OED_DU);
    static constexpr int NXD = BX + BX * NF;
    static constexpr int NQ = (MODE == OED_DS || MODE == OED_DU) ? 0 : (MODE == OED_CF ? NOBS * NTRI : NTRI);
    static constexpr int NX = NXD + NQ, NP = Base::NP + NW;
    static constexpr bool HAS_JVP = true;
    // end-state Fisher read-out the shooting kernel can fuse into its epilogue (one F_triu block: CS and CU)
    static constexpr int FIM_N = (MODE == OED_CS || MODE == OED_CU) ? NF : 0, FIM_OFF = NXD;
    struct Ctx {
        typename Base::Ctx base;
    };
    struct Force {
        int pcol;   // tangent p column of the augmented problem: < Base::NP a model parameter, >= a weight w_i; -1 none
    };
    // F-part of the right-hand side (dG = nullptr) or of its tangent (dG = tangent of G, wcol = weight index or -1)
    template <bool dG_ON>
    static __device__ __forceinline__ void fisher_rhs(const double *G, const double *dG, const double *w, int wcol, double *dF)
    {
        if constexpr (NQ == 0) return;
        if constexpr (MODE == OED_CU) {
            double s[NF], ds[NF];
#pragma unroll
            for (int a = 0; a < NF; a++) {
                s[a] = 0.0, ds[a] = 0.0;
#pragma unroll
                for (int i = 0; i < NOBS; i++) {
                    s[a] += G[i + BX * a];
                    if constexpr (dG_ON) ds[a] += dG[i + BX * a];
                }
            }
            int q = 0;
#pragma unroll
            for (int b = 0; b < NF; b++)
#pragma unroll
                for (int a = 0; a <= b; a++) dF[q++] = dG_ON ? fma(ds[a], s[b], s[a] * ds[b]) : s[a] * s[b];
        } else if constexpr (MODE == OED_CF) {
            int q = 0;
#pragma unroll
            for (int i = 0; i < NOBS; i++)
#pragma unroll
                for (int b = 0; b < NF; b++)
#pragma unroll
                    for (int a = 0; a <= b; a++)
                        dF[q++] = dG_ON ? fma(dG[i + BX * a], G[i + BX * b], G[i + BX * a] * dG[i + BX * b])
                                     : G[i + BX * a] * G[i + BX * b];
        } else {
            int q = 0;
#pragma unroll
            for (int b = 0; b < NF; b++)
#pragma unroll
                for (int a = 0; a <= b; a++) {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < NOBS; i++) {
                        const double gg = G[i + BX * a] * G[i + BX * b];
                        if constexpr (dG_ON) {
                            const double dgg = fma(dG[i + BX * a], G[i + BX * b], G[i + BX * a] * dG[i + BX * b]);
                            acc += (wcol == i ? gg : 0.0) + w[i] * dgg;
                        } else
                            acc += w[i] * gg;
                    }
                    dF[q++] = acc;
                }
        }
    }
    CB_DEV void f(const double *y, const double *p, double *dy, Ctx &c)
    {
        Base::f(y, p, dy, c.base);
#pragma unroll
        for (int k = 0; k < NF; k++)
            Base::jvp(y, p, c.base, y + BX + BX * k, Base::make_force(Traits::pcol(k), p), dy + BX + BX * k);
        fisher_rhs<false>(y + BX, y + BX, p + Base::NP, -1, dy + NXD);
    }
    CB_DEV Force make_force(int pcol, const double *) { return Force{pcol}; }
    // tangent of the augmented right-hand side (what ForwardDiff through the OED layer propagates):
    //   dx'  = f_x dx + f_p e_pcol
    //   dG'  = f_x dG + (f_xx[dx] + f_xp[e_pcol]) G + d(f_p columns)
    //   dF'  = sum_i dw_i (h_i G)'(h_i G) + w_i [ (h_i dG)'(h_i G) + (h_i G)'(h_i dG) ]     (OED_CS)
    CB_DEV void jvp(const double *y, const double *p, const Ctx &c, const double *v, const Force &F, double *dv)
    {
        const int bp = F.pcol < Base::NP ? F.pcol : -1;   // model-parameter part of the direction
        Base::jvp(y, p, c.base, v, Base::make_force(bp, p), dv);
#pragma unroll
        for (int k = 0; k < NF; k++) {
            double lin[BX], sec[BX];
            Base::jvp(y, p, c.base, v + BX + BX * k, Base::make_force(-1, p), lin);
            Base::d2(y, p, v, bp, y + BX + BX * k, Traits::pcol(k), sec);
#pragma unroll
            for (int r = 0; r < BX; r++) dv[BX + BX * k + r] = lin[r] + sec[r];
        }
        fisher_rhs<true>(y + BX, v + BX, p + Base::NP, F.pcol - Base::NP, dv + NXD);
    }
};
template <int M>
struct LotkaOedTraitsT {
    static constexpr int NF = 2, NOBS = 2, MODE = M;
    static __device__ __forceinline__ constexpr int pcol(int c) { return c + 1; }  // params = [2, 3]
};
template <int M>
struct Lin1dOedTraitsT {
    static constexpr int NF = 1, NOBS = 1, MODE = M;
    static __device__ __forceinline__ constexpr int pcol(int) { return 0; }  // params = [1]
};
using Lotka2Oed = OedAug<Lotka2, LotkaOedTraitsT<OED_CS>>;
using Lin1dOed = OedAug<Lin1d, Lin1dOedTraitsT<OED_CS>>;
using Lotka2OedCu = OedAug<Lotka2, LotkaOedTraitsT<OED_CU>>;
using Lotka2OedCf = OedAug<Lotka2, LotkaOedTraitsT<OED_CF>>;
using Lotka2OedDs = OedAug<Lotka2, LotkaOedTraitsT<OED_DS>>;
using Lotka2OedDu = OedAug<Lotka2, LotkaOedTraitsT<OED_DU>>;
using Lin1dOedCf = OedAug<Lin1d, Lin1dOedTraitsT<OED_CF>>;
using Lin1dOedDs = OedAug<Lin1d, Lin1dOedTraitsT<OED_DS>>;


// ---- Models written once on a generic scalar: the directional derivative is taken by forward-mode AD on the
//      device (a (value, tangent) pair through the same expression) -- what ForwardDiff does to the user's Julia
//      closure in the reference (src/dynprob.jl:150,161).  Adding a model of lib/OptimalControlBenchmarks is one
//      struct with a templated rhs.
struct Dual {
    double v, d;
};
CB_DEV Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
CB_DEV Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
CB_DEV Dual operator-(Dual a) { return {-a.v, -a.d}; }
CB_DEV Dual operator*(Dual a, Dual b) { return {a.v * b.v, fma(a.d, b.v, a.v * b.d)}; }
CB_DEV Dual operator/(Dual a, Dual b)
{
    const double q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
CB_DEV Dual operator+(Dual a, double c) { return {a.v + c, a.d}; }
CB_DEV Dual operator+(double c, Dual a) { return {a.v + c, a.d}; }
CB_DEV Dual operator-(Dual a, double c) { return {a.v - c, a.d}; }
CB_DEV Dual operator-(double c, Dual a) { return {c - a.v, -a.d}; }
CB_DEV Dual operator*(Dual a, double c) { return {a.v * c, a.d * c}; }
CB_DEV Dual operator*(double c, Dual a) { return {a.v * c, a.d * c}; }
CB_DEV Dual operator/(Dual a, double c) { return {a.v / c, a.d / c}; }
CB_DEV Dual operator/(double c, Dual b)
{
    const double q = c / b.v;
    return {q, -q * b.d / b.v};
}
CB_DEV double sin(double a) { return ::sin(a); }   // the Dual overloads below would hide the global ones
CB_DEV double cos(double a) { return ::cos(a); }
CB_DEV double exp(double a) { return ::exp(a); }
CB_DEV Dual sin(Dual a) { return {::sin(a.v), ::cos(a.v) * a.d}; }
CB_DEV Dual cos(Dual a) { return {::cos(a.v), -::sin(a.v) * a.d}; }
CB_DEV double sqrt(double a) { return ::sqrt(a); }
CB_DEV Dual sqrt(Dual a)
{
    const double r = ::sqrt(a.v);
    return {r, 0.5 * a.d / r};
}
CB_DEV Dual exp(Dual a)
{
    const double e = ::exp(a.v);
    return {e, e * a.d};
}

template <class M>
struct AutoJvp {
    static constexpr int NX = M::NX, NXD = M::NXD, NQ = M::NQ, NP = M::NP;
    static constexpr bool HAS_JVP = true;
    struct Ctx {};
    struct Force {
        int pcol;
    };
    CB_DEV void f(const double *x, const double *p, double *dx, Ctx &) { M::template rhs<double>(x, p, dx); }
    CB_DEV Force make_force(int pcol, const double *) { return Force{pcol}; }
    CB_DEV void jvp(const double *x, const double *p, const Ctx &, const double *v, const Force &F, double *dv)
    {
        Dual xd[NXD], pd[NP], out[NX];
#pragma unroll
        for (int k = 0; k < NXD; k++) xd[k] = Dual{x[k], v[k]};
#pragma unroll
        for (int q = 0; q < NP; q++) pd[q] = Dual{p[q], q == F.pcol ? 1.0 : 0.0};
        M::template rhs<Dual>(xd, pd, out);
#pragma unroll
        for (int k = 0; k < NX; k++) dv[k] = out[k].d;
    }
};

// Van der Pol oscillator with its Lagrange term (lib/OptimalControlBenchmarks/src/problems/van_der_pol.jl:16-33)
struct VdpRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = (1.0 - x[1] * x[1]) * x[0] - x[1] + p[0];
        dx[1] = x[0];
        dx[2] = x[0] * x[0] + x[1] * x[1] + p[0] * p[0];
    }
};
using Vdp3 = AutoJvp<VdpRhs>;

// Cart pendulum with its Lagrange term (lib/OptimalControlBenchmarks/src/problems/cart_pendulum.jl:17-48);
// states (x, theta, dx, dtheta, cost), p = (w).  The two denominators differ in the reference and are kept so.
struct CartPoleRhs {
    static constexpr int NX = 5, NXD = 4, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double al = 10.0, be = 50.0, ga = 0.5, M = 1.0, m = 0.1, g = 9.81, pi = 3.141592653589793;
        const T s = sin(x[1]), c = cos(x[1]);
        const T num = p[0] + (m * g) * s * c + m * (x[3] * x[3]) * s;
        const T omc = 1.0 - c;
        dx[0] = x[2];
        dx[1] = x[3];
        dx[2] = num / (M + m * (omc * omc));
        dx[3] = -g * s - (num / (M + m * (1.0 - c * c))) * c;
        const T dth = x[1] - pi;
        dx[4] = 1.0e-3 * (al * (x[0] * x[0]) + be * (dth * dth) + ga * (p[0] * p[0]));
    }
};
using CartPole5 = AutoJvp<CartPoleRhs>;

// Goddard rocket with free final time (lib/OptimalControlBenchmarks/src/problems/goddarts_rocket.jl:16-40);
// states (r, v, m), p = (T, u)
struct RocketRhs {
    static constexpr int NX = 3, NXD = 3, NQ = 0, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double b = 7.0, Tmax = 3.5, A = 310.0, k = 500.0, r0 = 1.0;
        const T drag = A * (x[1] * x[1]) * exp(-k * (x[0] - r0));
        dx[0] = p[0] * x[1];
        dx[1] = p[0] * (-1.0 / (x[0] * x[0]) + (1.0 / x[2]) * (Tmax * p[1] - drag));
        dx[2] = p[0] * (-b * p[1]);
    }
};
using Rocket3 = AutoJvp<RocketRhs>;


// Quadrotor with its Lagrange term (lib/OptimalControlBenchmarks/src/problems/quadrotor.jl:9-52): states (x1..x6, cost),
// p = (w1, w2, w3, u) -- four controls
struct QuadrotorRhs {
    static constexpr int NX = 7, NXD = 6, NQ = 1, NP = 4;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double g = 9.81, M = 1.3, L = 0.305, I = 0.0605;
        const T s5 = sin(x[4]), c5 = cos(x[4]);
        dx[0] = x[1];
        dx[1] = g * s5 + p[0] * p[3] * s5 / M;
        dx[2] = x[3];
        dx[3] = g * c5 - g + p[0] * p[3] * c5 / M;
        dx[4] = x[5];
        dx[5] = (p[2] - p[1]) * (L / I) * p[3];
        dx[6] = 5.0 * (p[3] * p[3]);
    }
};
using Quadrotor7 = AutoJvp<QuadrotorRhs>;

// Three-tank system with its Lagrange term (.../three_tank.jl:9-53): states (x1, x2, x3, cost), p = (w1, w2, w3)
struct ThreeTankRhs {
    static constexpr int NX = 4, NXD = 3, NQ = 1, NP = 3;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double c1 = 1.0, c2 = 2.0, c3 = 0.8, k1 = 2.0, k2 = 3.0, k3 = 1.0, k4 = 3.0;
        const T r1 = sqrt(x[0]), r2 = sqrt(x[1]), r3 = sqrt(x[2]), q = p[2] * sqrt(c3 * x[0]);
        dx[0] = c1 * p[0] + c2 * p[1] - r2 - q;
        dx[1] = r1 - r2;
        dx[2] = r2 - r3 + q;
        const T e2 = x[1] - k2, e3 = x[2] - k4;
        dx[3] = k1 * (e2 * e2) + k3 * (e3 * e3);
    }
};
using ThreeTank4 = AutoJvp<ThreeTankRhs>;

// Batch reactor (.../batch_reactor.jl:9-17): states (x1, x2), p = (u) -- Arrhenius terms exp(-c / u)
struct BatchReactorRhs {
    static constexpr int NX = 2, NXD = 2, NQ = 0, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const T r1 = 4.0e3 * exp(-2500.0 / p[0]) * (x[0] * x[0]);
        const T r2 = 62.0e4 * exp(-5000.0 / p[0]) * (x[1] * x[1]);
        dx[0] = -r1;
        dx[1] = r1 - r2;
    }
};
using BatchReactor2 = AutoJvp<BatchReactorRhs>;


// Competitive Lotka-Volterra with its Lagrange term (.../lotka_competitive.jl:9-31): states (x0, x1, cost), p = (u)
struct LotkaCompRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double c1 = 0.1, c2 = 0.4, al = 1.2, K = 1.8;
        dx[0] = x[0] * (1.0 - (x[0] + al * x[1]) / K) - c1 * x[0] * p[0];
        dx[1] = x[1] * (1.0 - (x[0] + x[1]) / K) - c2 * x[1] * p[0];
        const T e0 = x[0] - 1.0, e1 = x[1] - 1.0;
        dx[2] = e0 * e0 + e1 * e1;
    }
};
using LotkaComp3 = AutoJvp<LotkaCompRhs>;

// F-8 aircraft, time-optimal (.../f8.jl:9-41): states (x0, x1, x2, obj), p = (T, u); obj' = T is quadrature-like
struct F8Rhs {
    static constexpr int NX = 4, NXD = 3, NQ = 1, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double xi = 0.05236;
        const T x0sq = x[0] * x[0], x0cu = x0sq * x[0];
        T d1 = -0.877 * x[0] + x[2] - 0.088 * x[0] * x[2] + 0.47 * x0sq - 0.019 * (x[1] * x[1]);
        d1 = d1 - x0sq * x[2] + 3.846 * x0cu;
        d1 = d1 + (0.215 * xi - 0.63 * xi * xi * xi) - (0.28 * xi) * x0sq + (0.47 * xi * xi) * x[0];
        d1 = d1 - ((0.215 * xi - 0.63 * xi * xi * xi) - (0.28 * xi) * x0sq) * (2.0 * p[1]);
        T d3 = -4.208 * x[0] - 0.396 * x[2] - 0.47 * x0sq - 3.564 * x0cu;
        d3 = d3 + (20.967 * xi - 61.4 * xi * xi * xi) - (6.265 * xi) * x0sq + (46.0 * xi * xi) * x[0];
        d3 = d3 - ((20.967 * xi - 61.4 * xi * xi * xi) - (6.265 * xi) * x0sq) * (2.0 * p[1]);
        dx[0] = p[0] * d1;
        dx[1] = p[0] * x[2];
        dx[2] = p[0] * d3;
        dx[3] = p[0];
    }
};
using F8Aircraft4 = AutoJvp<F8Rhs>;

// Double oscillator with its Lagrange term (.../double_oscillator.jl:9-38).  The right-hand side depends on time
// (sin(2 pi t / T)); the kernels integrate autonomous systems, so time rides along as a state with tau' = 1:
// states (x0, x1, x2, x3, tau, cost), p = (u)
struct DoubleOscRhs {
    static constexpr int NX = 6, NXD = 5, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double m1 = 200.0, m2 = 2.0, k1 = 100.0, k2 = 3.0, c = 0.5;   // T = 2 pi: sin(2 pi / T * t) = sin(t)
        dx[0] = x[2];
        dx[1] = x[3];
        dx[2] = (-(k1 + k2) / m1) * x[0] + (k2 / m1) * x[1] + (1.0 / m1) * sin(x[4]);
        dx[3] = (k2 / m2) * x[0] - (k2 / m2) * x[1] - (c / m2) * (1.0 - p[0]) * x[3];
        dx[4] = 1.0 + 0.0 * x[4];
        dx[5] = 0.5 * (x[0] * x[0] + x[1] * x[1] + p[0] * p[0]);
    }
};
using DoubleOsc6 = AutoJvp<DoubleOscRhs>;

// lib/OptimalControlBenchmarks/src/problems/bryson_denham.jl:9-36; states (x, v, cost), p = (w)
struct BrysonDenhamRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = x[1];
        dx[1] = p[0];
        dx[2] = p[0] * p[0];
    }
};
using BrysonDenham3 = AutoJvp<BrysonDenhamRhs>;

// lib/OptimalControlBenchmarks/src/problems/catalyst_mixing.jl:9-22; states (x1, x2), p = (w)
struct CatalystMixingRhs {
    static constexpr int NX = 2, NXD = 2, NQ = 0, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const T a = 10.0 * x[1] - x[0];
        dx[0] = p[0] * a;
        dx[1] = p[0] * (x[0] - 10.0 * x[1]) - (1.0 - p[0]) * x[1];
    }
};
using CatalystMixing2 = AutoJvp<CatalystMixingRhs>;

// lib/OptimalControlBenchmarks/src/problems/cushioned_oscillation.jl:9-37; states (x, v, obj), p = (t_s, u)
struct CushionedRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double m = 5.0, c = 10.0;
        dx[0] = p[0] * x[1];
        dx[1] = p[0] * ((1.0 / m) * (p[1] - c * x[0]));
        dx[2] = p[0];
    }
};
using Cushioned3 = AutoJvp<CushionedRhs>;

// lib/OptimalControlBenchmarks/src/problems/denbigh.jl:9-37; states (x1, x2, x3), p = (T); the second equation is `k1 x1 - k3 + k4 x2` as the reference writes it
struct DenbighRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const T e1 = exp(-3.0e3 / p[0]), e2 = exp(-6.0e3 / p[0]);
        const T k1 = 1.0e3 * e1, k2 = 1.0e7 * e2, k3 = 1.0e1 * e1;
        dx[0] = -(k1 * x[0]) - k2 * x[0];
        dx[1] = k1 * x[0] - k3 + 1.0e-3 * x[1];
        dx[2] = k3 * x[1];
    }
};
using Denbigh3 = AutoJvp<DenbighRhs>;

// lib/OptimalControlBenchmarks/src/problems/dielectrophoretic_particle.jl:9-34; states (x0, x1, obj), p = (t_s, u)
struct DielectrophoreticRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = p[0] * (x[1] * p[1] + (-0.75) * (p[1] * p[1]));
        dx[1] = p[0] * (p[1] - x[1]);
        dx[2] = p[0];
    }
};
using Dielectrophoretic3 = AutoJvp<DielectrophoreticRhs>;

// lib/OptimalControlBenchmarks/src/problems/electric_car.jl:9-62; states (x1, x2, x3, cost), p = (u)
struct ElectricCarRhs {
    static constexpr int NX = 4, NXD = 2, NQ = 2, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double Kr = 10.0, rho = 1.293, Cx = 0.4, S = 2.0, r = 0.33, Kf = 0.03, Km = 0.27, Rm = 0.03, Lm = 0.05,
                     M = 250.0, g = 9.81, V = 150.0, Rb = 0.05;
        dx[0] = (V * p[0] - Rm * x[0] - Km * x[1]) / Lm;
        dx[1] = (Km * x[0] - (r / Kr) * (M * g * Kf + (0.5 * rho * S * Cx * r * r / (Kr * Kr)) * (x[1] * x[1]))) * (Kr * Kr / (M * r * r));
        dx[2] = (r / Kr) * x[1];
        dx[3] = V * p[0] * x[0] + Rb * (x[0] * x[0]);
    }
};
using ElectricCar4 = AutoJvp<ElectricCarRhs>;

// lib/OptimalControlBenchmarks/src/problems/fuller.jl:9-32; states (x0, x1, cost), p = (u)
struct FullerRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = x[1];
        dx[1] = 1.0 - 2.0 * p[0];
        dx[2] = x[0] * x[0];
    }
};
using Fuller3 = AutoJvp<FullerRhs>;

// lib/OptimalControlBenchmarks/src/problems/jackson.jl:9-50; states (x1, x2, x3), p = (u)
struct JacksonRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const T a = 1.0 * x[0] - 10.0 * x[1];
        const T b = (1.0 - p[0]) * (1.0 * x[1]);
        dx[0] = -(p[0] * a);
        dx[1] = p[0] * a - b;
        dx[2] = b;
    }
};
using Jackson3 = AutoJvp<JacksonRhs>;

// lib/OptimalControlBenchmarks/src/problems/mountain_car.jl:9-31; states (x, v, obj), p = (t_s, u)
struct MountainCarRhs {
    static constexpr int NX = 3, NXD = 2, NQ = 1, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = p[0] * x[1];
        dx[1] = p[0] * (1.0e-3 * p[1] - 2.5e-3 * cos(3.0 * x[0]));
        dx[2] = p[0];
    }
};
using MountainCar3 = AutoJvp<MountainCarRhs>;

// lib/OptimalControlBenchmarks/src/problems/particle_steering.jl:9-42; states (x1, x2, dx1, dx2, obj), p = (t_s, u)
struct ParticleSteeringRhs {
    static constexpr int NX = 5, NXD = 4, NQ = 1, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = p[0] * x[2];
        dx[1] = p[0] * x[3];
        dx[2] = p[0] * (100.0 * cos(p[1]));
        dx[3] = p[0] * (100.0 * sin(p[1]));
        dx[4] = p[0];
    }
};
using ParticleSteering5 = AutoJvp<ParticleSteeringRhs>;

// lib/OptimalControlBenchmarks/src/problems/rao_mease.jl:9-26; states (x, cost), p = (u)
struct RaoMeaseRhs {
    static constexpr int NX = 2, NXD = 1, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = -(x[0] * x[0] * x[0]) + p[0];
        dx[1] = x[0] * x[0] + p[0] * p[0];
    }
};
using RaoMease2 = AutoJvp<RaoMeaseRhs>;

// lib/OptimalControlBenchmarks/src/problems/robbins.jl:9-44; states (x1, x2, x3, cost), p = (u); alpha = 3, beta = 0, gamma = 0.5
struct RobbinsRhs {
    static constexpr int NX = 4, NXD = 3, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = x[1];
        dx[1] = x[2];
        dx[2] = p[0];
        dx[3] = 3.0 * x[0] + 0.0 * (x[0] * x[0]) + 0.5 * (p[0] * p[0]);
    }
};
using Robbins4 = AutoJvp<RobbinsRhs>;

// lib/OptimalControlBenchmarks/src/problems/moon_landing.jl:9-31; states (h, v, m), p = (t_s, T)
struct MoonLandingRhs {
    static constexpr int NX = 3, NXD = 3, NQ = 0, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = p[0] * x[1];
        dx[1] = p[0] * (-1.0 + p[1] / x[2]);
        dx[2] = p[0] * (-(p[1] / 2.349));
    }
};
using MoonLanding3 = AutoJvp<MoonLandingRhs>;

// lib/OptimalControlBenchmarks/src/problems/lotka_shared.jl:9-32; states (x0, x1, x2, cost), p = (u)
struct LotkaSharedRhs {
    static constexpr int NX = 4, NXD = 3, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = x[0] - x[0] * x[1] - x[0] * x[2];
        dx[1] = -x[1] + x[0] * x[1] - 0.1 * x[1] * p[0];
        dx[2] = -x[2] + 1.2 * x[0] * x[2] - 0.4 * x[2] * p[0];
        dx[3] = (x[0] - 1.5) * (x[0] - 1.5) + (x[1] - 1.0) * (x[1] - 1.0) + (x[2] - 1.0) * (x[2] - 1.0);
    }
};
using LotkaShared4 = AutoJvp<LotkaSharedRhs>;

// lib/OptimalControlBenchmarks/src/problems/hanging_chain.jl:9-48; states (x1, x2, x3), p = (u)
struct HangingChainRhs {
    static constexpr int NX = 3, NXD = 1, NQ = 2, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const T s = sqrt(1.0 + p[0] * p[0]);
        dx[0] = p[0];
        dx[1] = x[0] * s;
        dx[2] = s;
    }
};
using HangingChain3 = AutoJvp<HangingChainRhs>;

// lib/OptimalControlBenchmarks/src/problems/tubular_reactor.jl:9-22; states (x1, x2), p = (u)
struct TubularReactorRhs {
    static constexpr int NX = 2, NXD = 1, NQ = 1, NP = 1;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        dx[0] = -((p[0] + 0.5 * (p[0] * p[0])) * x[0]);
        dx[1] = p[0] * x[0];
    }
};
using TubularReactor2 = AutoJvp<TubularReactorRhs>;

// lib/OptimalControlBenchmarks/src/problems/ocean.jl:9-64; states (S, R, tau, cost), p = (u1, u2); time rides along as a state (tau' = 1) for the discount factor
struct OceanRhs {
    static constexpr int NX = 4, NXD = 3, NQ = 1, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const T Dl = 3.5e4 - x[1] - x[0];
        dx[0] = -p[0];
        dx[1] = p[0] - p[1] - 1.0e-3 * (x[0] - 0.1 * Dl);
        dx[2] = 1.0 + 0.0 * x[2];
        const T U = 50.0 * p[0] - 0.5 * (p[0] * p[0]), A = 2.0 * p[1] + 2.0 * (p[1] * p[1]), C = 50.0 - 4.0e-3 * x[1];
        const T q = 0.3 * x[0] - 6.0e2;
        dx[3] = -(exp(-3.0e-2 * x[2]) * (U - A - p[0] * C - q * q));
    }
};
using Ocean4 = AutoJvp<OceanRhs>;

// lib/OptimalControlBenchmarks/src/problems/ducted_fan.jl:9-61; states (x1, v1, x2, v2, alpha, valpha, cost), p = (t_s, u1, u2)
struct DuctedFanRhs {
    static constexpr int NX = 7, NXD = 6, NQ = 1, NP = 3;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double m = 2.2, J = 0.05, r = 0.2, mg = 4.0, mu = 1.0;
        const T s = sin(x[4]), c = cos(x[4]);
        dx[0] = p[0] * x[1];
        dx[1] = p[0] * ((1.0 / m) * (p[1] * c - p[2] * s));
        dx[2] = p[0] * x[3];
        dx[3] = p[0] * ((1.0 / m) * (-mg + p[1] * s + p[2] * c));
        dx[4] = p[0] * x[5];
        dx[5] = p[0] * ((r / J) * p[1]);
        dx[6] = (2.0 * (p[1] * p[1]) + p[2] * p[2]) + mu * p[0];
    }
};
using DuctedFan7 = AutoJvp<DuctedFanRhs>;

// lib/OptimalControlBenchmarks/src/problems/hang_glider.jl:9-66; states (x, y, vx, vy), p = (T, cL)
struct HangGliderRhs {
    static constexpr int NX = 4, NXD = 4, NQ = 0, NP = 2;
    template <class T>
    CB_DEV void rhs(const T *x, const T *p, T *dx)
    {
        const double uc = 2.5, rc = 100.0, c0 = 0.034, c1 = 0.069662, S = 14.0, rho = 1.13, m = 100.0, g = 9.81;
        const T rr = x[0] / rc - 2.5;
        const T r = rr * rr;
        const T Uup = uc * (1.0 - r) * exp(-r);
        const T w = x[3] - Uup;
        const T v2 = x[2] * x[2] + w * w;
        const T v = sqrt(v2);
        const T Dr = (0.5 * rho * S) * (c0 + c1 * (p[1] * p[1])) * v2;
        const T L = (0.5 * rho * S) * p[1] * v2;
        dx[0] = p[0] * x[2];
        dx[1] = p[0] * x[3];
        dx[2] = (p[0] * (-1.0)) * (L * w + Dr * x[2]) / (m * v);
        dx[3] = p[0] * ((L * x[2] - Dr * w) / (m * v) - g);
    }
};
using HangGlider4 = AutoJvp<HangGliderRhs>;

}  // namespace cb200
