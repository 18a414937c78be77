// K1/K2: batched integration of every (interval, candidate) shooting unit with forward sensitivities.
//
// Replaces mythreadmap over intervals (src/Corleone.jl:22-24, src/multiple_shooting.jl:118-130) of the
// piece loop _sequential_solve (src/single_shooting.jl:421-471), and the ForwardDiff sweep through it
// (src/dynprob.jl:150,161), for B candidates at once.
//
// Mapping: one thread = (candidate b, interval i, direction group r).  The thread integrates the state x
// and G sensitivity directions d = r*G .. r*G+G-1 of its unit, everything in registers and fully
// unrolled.  Directions are the local variables of the interval in reference order (u0, p, controls)
// (src/single_shooting.jl:190-200), so `sens` is written directly in Jacobian-block order.  b is the
// fastest thread index: with the SoA layout [variable][candidate] every global load and store of a warp
// is one coalesced 256-byte transaction, and (i, r) is warp-uniform whenever B is a multiple of 32.
//
// Stepping: fixed-step Tsit5 (7-stage FSAL; k1 re-evaluated at each piece start because every control
// piece is a fresh `solve` in the reference, src/single_shooting.jl:443-453) or classical RK4; forward
// sensitivities = the same RK scheme applied to the variational system S' = f_x S + f_p (what
// ForwardDiff through a fixed-step solve computes).
//
// What keeps the FP64 pipe busy (DESIGN.md "kernel"):
//   * the step body is one straight-line block: parameter directions enter through per-piece weights
//     (Model::Force), not branches;
//   * the products h*a_ij are kernel-parameter constants when the step size is uniform over the layer
//     (template UNI), so every stage combination is a pure DFMA chain with constant-bank operands;
//   * quadrature states (and their sensitivities) never feed a right-hand side: they are accumulated
//     stage by stage (y_q += h b_j k_j) and need neither stage inputs nor k storage;
//   * the FSAL evaluation after the last step of a piece is skipped (its value is never used).
#pragma once
#include "common.cuh"
#include "tableau.cuh"
#include "fim.cuh"
#include <cstdlib>

namespace cb200 {

// kind-major defect `pos` of candidate column b: to the local block and -- fused all-gather -- to every peer's window
// (P2P stores over NVLink; tail.cuh has the completion protocol)
__device__ __forceinline__ void store_defect(const ShootOut &o, long long pos, long long b, double dd)
{
    const long long q = pos * o.ldk + b;
    o.dk[q] = dd;
    for (int p = 0; p < o.n_dk_peers; p++)
        if (p != o.dk_self) (o.dk_peers[p] + o.dk_peer_off)[q] = dd;
}

// rhs of [x; S_1..S_G] at dynamic stage values td -> dynamic derivatives kd, quadrature derivatives kq
template <class Mdl, int G>
__device__ __forceinline__ void rhs_all(const double *td, const double *p, const typename Mdl::Force *F, double *kd,
                                        double *kq)
{
    constexpr int NXD = Mdl::NXD, NQ = Mdl::NQ, NX = Mdl::NX;
    typename Mdl::Ctx c;
    double out[NX];
    Mdl::f(td, p, out, c);
#pragma unroll
    for (int k = 0; k < NXD; k++) kd[k] = out[k];
#pragma unroll
    for (int k = 0; k < NQ; k++) kq[k] = out[NXD + k];
    if constexpr (G > 0) {
#pragma unroll
        for (int g = 0; g < G; g++) {
            double o2[NX];
            Mdl::jvp(td, p, c, td + NXD * (1 + g), F[g], o2);
#pragma unroll
            for (int k = 0; k < NXD; k++) kd[NXD * (1 + g) + k] = o2[k];
#pragma unroll
            for (int k = 0; k < NQ; k++) kq[NQ * (1 + g) + k] = o2[NXD + k];
        }
    }
}

template <class Mdl, int G, int METHOD, bool SQ = true>
__device__ __forceinline__ void integrate_piece(double *yd, double *yq, const double *p, const typename Mdl::Force *F,
                                                const double *cf, int K)
{
    constexpr int ND = Mdl::NXD * (1 + G), NQ = (Mdl::NQ * (1 + G) > 0) ? Mdl::NQ * (1 + G) : 1;
    constexpr int NQR = Mdl::NQ * (1 + G);
    constexpr int Q0 = SQ ? 0 : Mdl::NQ;   // direction groups other than the first do not need the state's quadratures
    double kq[NQ];
    if constexpr (METHOD == 0) {
        // Accumulator form of the 7-stage scheme: as soon as a stage derivative k_j is known it is folded into
        // the pending stage inputs s_i = y + h sum_{l<=j} a_il k_l (i > j) and into the new state (row 7), in the
        // same left-to-right order as the textbook sums -- bitwise the same result, but only the five pending
        // inputs and one k are live instead of six k's, y and a stage input.
        double k[ND], s2[ND], s3[ND], s4[ND], s5[ND], s6[ND];
        rhs_all<Mdl, G>(yd, p, F, k, kq);
        for (int s = 0; s < K; s++) {
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[15], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                s2[c] = fma(cf[0], k[c], yd[c]);
                s3[c] = fma(cf[1], k[c], yd[c]);
                s4[c] = fma(cf[3], k[c], yd[c]);
                s5[c] = fma(cf[6], k[c], yd[c]);
                s6[c] = fma(cf[10], k[c], yd[c]);
                yd[c] = fma(cf[15], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(s2, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[16], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                s3[c] = fma(cf[2], k[c], s3[c]);
                s4[c] = fma(cf[4], k[c], s4[c]);
                s5[c] = fma(cf[7], k[c], s5[c]);
                s6[c] = fma(cf[11], k[c], s6[c]);
                yd[c] = fma(cf[16], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(s3, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[17], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                s4[c] = fma(cf[5], k[c], s4[c]);
                s5[c] = fma(cf[8], k[c], s5[c]);
                s6[c] = fma(cf[12], k[c], s6[c]);
                yd[c] = fma(cf[17], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(s4, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[18], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                s5[c] = fma(cf[9], k[c], s5[c]);
                s6[c] = fma(cf[13], k[c], s6[c]);
                yd[c] = fma(cf[18], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(s5, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[19], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                s6[c] = fma(cf[14], k[c], s6[c]);
                yd[c] = fma(cf[19], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(s6, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[20], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) yd[c] = fma(cf[20], k[c], yd[c]);
            if (s + 1 < K)  // FSAL: k7 of this step is k1 of the next; after the last step it is never used
                rhs_all<Mdl, G>(yd, p, F, k, kq);
        }
    } else {
        double k[ND], acc[ND], td[ND];
        for (int s = 0; s < K; s++) {
            rhs_all<Mdl, G>(yd, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[2], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                acc[c] = fma(cf[2], k[c], yd[c]);
                td[c] = fma(cf[0], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(td, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[3], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                acc[c] = fma(cf[3], k[c], acc[c]);
                td[c] = fma(cf[0], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(td, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[3], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                acc[c] = fma(cf[3], k[c], acc[c]);
                td[c] = fma(cf[1], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(td, p, F, k, kq);
#pragma unroll
            for (int c = Q0; c < NQR; c++) yq[c] = fma(cf[2], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) yd[c] = fma(cf[2], k[c], acc[c]);
        }
    }
}

// ---- saveat: samples strictly inside a control piece come from the dense output of the step that contains them
// (problem.kwargs saveat, lib/CorleoneOED/src/oed.jl:106-113; OrdinaryDiffEq interpolates, it does not cut steps).
// Tsit5's interpolant (Tsitouras 2011, the 4th-order "free" continuous extension OrdinaryDiffEq uses for Tsit5):
//   y(t + theta dt) = y + dt sum_i b_i(theta) k_i,  b_1 = theta (r11 + theta (r12 + theta (r13 + theta r14))),
//   b_i = theta^2 (ri2 + theta (ri3 + theta ri4)), i = 2..7;  b_i(1) = a_7i (checked in tests/test_saveat_host.py)
__device__ __forceinline__ void tsit5_interp_weights(double th, double *bw)
{
    constexpr double R[7][4] = {{1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
                                {0.0, 0.13169999999999998, -0.2234, 0.1017},
                                {0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253},
                                {0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902},
                                {0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928},
                                {0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661},
                                {0.0, 1.5, -4.0, 2.5}};
    const double th2 = th * th;
    bw[0] = th * fma(th, fma(th, fma(th, R[0][3], R[0][2]), R[0][1]), R[0][0]);
#pragma unroll
    for (int i = 1; i < 7; i++) bw[i] = th2 * fma(th, fma(th, R[i][3], R[i][2]), R[i][1]);
}

template <class Mdl>
__device__ __forceinline__ void save_sample(const DevLayer &L, const ShootOut &o, long long sample, int nrow, long long b,
                                         const double *yd, const double *yq, const double *p);
template <class Mdl, int G>
__device__ __forceinline__ void save_sens_sample(const ShootOut &o, long long row0, long long b, const int *dmap,
                                                 const double *yd, const double *yq);

// no saveat: the integrators compile their sampling code away
struct NoSampler {
    static constexpr bool ON = false;
    int n = 0;
    __device__ __forceinline__ double time(int) const { return 0.0; }
    __device__ __forceinline__ void emit(int, const double *, const double *) const {}
};
// the saveat times of one piece and where their samples go
template <class Mdl, int G, bool TS>
struct PieceSampler {
    static constexpr bool ON = true;
    const DevLayer *L;
    const ShootOut *o;
    const double *t;      // the piece's saveat times
    int n;
    long long b;
    long long samp0;      // merged-trajectory index of the first of them
    long long tsrow0;     // traj_sens row of the first of them
    int ns, nrow;
    const int *dmap;
    const double *p;
    bool write_traj, write_tsens;
    __device__ __forceinline__ double time(int m) const { return __ldg(t + m); }
    __device__ __forceinline__ void emit(int m, const double *zd, const double *zq) const
    {
        if (write_traj) save_sample<Mdl>(*L, *o, samp0 + m, nrow, b, zd, zq, p);
        if constexpr (G > 0 && TS) {
            if (write_tsens) save_sens_sample<Mdl, G>(*o, tsrow0 + (long long)m * ns * Mdl::NX, b, dmap, zd, zq);
        }
    }
};

// samples of the pending saveat times up to tnew from the step (tprev, dt) with start values (y0d, y0q) and stage
// derivatives kd[7][ND], kq[7][NQ]; returns the new count of emitted samples
template <int ND, int NQR, int NQA, class SMP>
__device__ __forceinline__ int emit_pending(const SMP &smp, int m, double tprev, double dt, double tnew, const double *y0d,
                                            const double *y0q, const double (*kd)[ND], const double (*kq)[NQA])
{
    while (m < smp.n && smp.time(m) <= tnew) {
        double bw[7], zd[ND], zq[NQA];
        tsit5_interp_weights((smp.time(m) - tprev) / dt, bw);
#pragma unroll
        for (int c = 0; c < ND; c++) {
            double acc = kd[0][c] * bw[0];
#pragma unroll
            for (int j = 1; j < 7; j++) acc = fma(kd[j][c], bw[j], acc);
            zd[c] = fma(dt, acc, y0d[c]);
        }
#pragma unroll
        for (int c = 0; c < NQR; c++) {
            double acc = kq[0][c] * bw[0];
#pragma unroll
            for (int j = 1; j < 7; j++) acc = fma(kq[j][c], bw[j], acc);
            zq[c] = fma(dt, acc, y0q[c]);
        }
        smp.emit(m, zd, zq);
        m++;
    }
    return m;
}

// fixed-step Tsit5 over one piece WITH saveat samples: the textbook form that keeps every stage derivative of the step
// (the accumulator form of integrate_piece folds them away).  Same arithmetic per stage, so the end state agrees with
// integrate_piece to rounding.  Not a throughput path (the discrete OED layers).
template <class Mdl, int G, class SMP>
__device__ __noinline__ void integrate_piece_dense(double *yd, double *yq, const double *p, const typename Mdl::Force *F,
                                                   double t0, double t1, int K, const SMP &smp)
{
    using namespace ts5;
    constexpr int ND = Mdl::NXD * (1 + G), NQR = Mdl::NQ * (1 + G), NQA = NQR > 0 ? NQR : 1;
    constexpr double A[7][6] = {{0, 0, 0, 0, 0, 0},         {a21, 0, 0, 0, 0, 0},       {a31, a32, 0, 0, 0, 0},
                                {a41, a42, a43, 0, 0, 0},   {a51, a52, a53, a54, 0, 0}, {a61, a62, a63, a64, a65, 0},
                                {a71, a72, a73, a74, a75, a76}};
    double kd[7][ND], kq[7][NQA], td[ND], nq[NQA];
    const double h = (t1 - t0) / (double)K;
    int m = 0;
    rhs_all<Mdl, G>(yd, p, F, kd[0], kq[0]);
    for (int s = 0; s < K; s++) {
#pragma unroll
        for (int st = 1; st < 7; st++) {
#pragma unroll
            for (int c = 0; c < ND; c++) {
                double acc = yd[c];
#pragma unroll
                for (int j = 0; j < st; j++) acc = fma(h * A[st][j], kd[j][c], acc);
                td[c] = acc;
            }
            rhs_all<Mdl, G>(td, p, F, kd[st], kq[st]);
        }
#pragma unroll
        for (int c = 0; c < NQR; c++) {
            double acc = yq[c];
#pragma unroll
            for (int j = 0; j < 6; j++) acc = fma(h * A[6][j], kq[j][c], acc);
            nq[c] = acc;
        }
        const double tprev = t0 + s * h, tnew = (s + 1 == K) ? t1 : t0 + (s + 1) * h;
        m = emit_pending<ND, NQR, NQA>(smp, m, tprev, h, tnew, yd, yq, kd, kq);
#pragma unroll
        for (int c = 0; c < ND; c++) {
            yd[c] = td[c];   // the 7th stage input is the new state
            kd[0][c] = kd[6][c];
        }
#pragma unroll
        for (int c = 0; c < NQR; c++) {
            yq[c] = nq[c];
            kq[0][c] = kq[6][c];
        }
    }
}

// ---- method 2: adaptive Tsit5, the reference's default mode (`solve(problem, Tsit5(); ...)` per control piece with the
// problem's abstol / reltol, src/single_shooting.jl:443-453).  OrdinaryDiffEq's stepping restated from its published
// algorithm (has the long form and the pin on the reference's golden vectors): Hairer's initial-step
// heuristic, embedded 4th-order error estimate in the RMS norm, PI controller with the Float32 `fastpow`, steps clipped
// to the end of the piece; every piece starts over.  The step sequence is decided by the STATE components only; the
// thread's sensitivity directions ride along on it (all direction groups of a unit therefore take identical steps).
// Not a throughput path: the k's are kept (registers / local memory) and lanes of a warp may take different step counts.
__device__ __forceinline__ float cb_fastlog2(float x)
{
    const float a = 0.338953f, b = 2.198599f, c = 1.523692f;
    const unsigned ux = __float_as_uint(x);
    const unsigned e = (ux & 0x7F800000u) >> 23;
    const bool greater = (ux & 0x00400000u) != 0u;   // significand > 1.5
    const float signif = __fsub_rn(__uint_as_float((ux & 0x007FFFFFu) | (greater ? 0x3f000000u : 0x3f800000u)), 1.0f);
    const float fexp = __fsub_rn((float)e, greater ? 126.0f : 127.0f);
    // no FMA contraction: Julia evaluates (a*s + b)*s/(s + c) with separate roundings
    return __fadd_rn(fexp, __fdiv_rn(__fmul_rn(__fadd_rn(__fmul_rn(a, signif), b), signif), __fadd_rn(signif, c)));
}
__device__ __forceinline__ double cb_fastpow(double x, double y)
{
    // exp2 of the Float32 product, rounded to Float32 (through double: correctly rounded, as Julia's exp2(::Float32) is to < 1 ulp)
    return x == 0.0 ? 0.0 : (double)(float)exp2((double)__fmul_rn((float)y, cb_fastlog2((float)x)));
}

template <class Mdl, int G, class SMP = NoSampler>
__device__ __noinline__ void integrate_piece_adaptive(double *yd, double *yq, const double *p, const typename Mdl::Force *F,
                                                      double t0, double t1, double abstol, double reltol, const SMP &smp = SMP())
{
    int n_emitted = 0;
    (void)n_emitted;
    using namespace ts5;
    constexpr int NXD = Mdl::NXD, NQS = Mdl::NQ, NX = Mdl::NX;
    constexpr int ND = NXD * (1 + G), NQR = NQS * (1 + G), NQA = NQR > 0 ? NQR : 1, NQS1 = NQS > 0 ? NQS : 1;
    constexpr double A[7][6] = {{0, 0, 0, 0, 0, 0},         {a21, 0, 0, 0, 0, 0},       {a31, a32, 0, 0, 0, 0},
                                {a41, a42, a43, 0, 0, 0},   {a51, a52, a53, a54, 0, 0}, {a61, a62, a63, a64, a65, 0},
                                {a71, a72, a73, a74, a75, a76}};
    constexpr double BT[7] = {bt1, bt2, bt3, bt4, bt5, bt6, bt7};
    double kd[7][ND], kq[7][NQA], td[ND], nd[ND], nq[NQA];
    const double dtmax = t1 - t0;
    const double beta1 = 7.0 / 50.0, beta2 = 2.0 / 25.0, gamma = 0.9, qmin = 0.2, qmax = 10.0, qoldinit = 1.0e-4;
    double t = t0, dt, qold = qoldinit;
    rhs_all<Mdl, G>(yd, p, F, kd[0], kq[0]);
    {   // initial step on the state components (dynamic ones in yd[0..NXD), quadrature-like ones in yq[0..NQ))
        const double smalldt = 1.0e-6;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int c = 0; c < NX; c++) {
            const double y = c < NXD ? yd[c < NXD ? c : 0] : yq[c >= NXD ? c - NXD : 0];
            const double f = c < NXD ? kd[0][c < NXD ? c : 0] : kq[0][c >= NXD ? c - NXD : 0];
            const double sk = abstol + fabs(y) * reltol;
            s0 += (y / sk) * (y / sk);
            s1 += (f / sk) * (f / sk);
        }
        const double d0 = sqrt(s0 / NX), d1 = sqrt(s1 / NX);
        double dt0 = (d0 < 1.0e-5 || d1 < 1.0e-5) ? smalldt : 0.01 * (d0 / d1);
        dt0 = fmin(dt0, dtmax);
        if (dt0 < 10.0 * 2.220446049250313e-16) {
            dt = smalldt;
        } else {
            double u1[NXD], f1d[NXD], f1q[NQS1];
#pragma unroll
            for (int c = 0; c < NXD; c++) u1[c] = yd[c] + dt0 * kd[0][c];
            rhs_all<Mdl, 0>(u1, p, F, f1d, f1q);
            double s2 = 0.0;
#pragma unroll
            for (int c = 0; c < NX; c++) {
                const double y = c < NXD ? yd[c < NXD ? c : 0] : yq[c >= NXD ? c - NXD : 0];
                const double f0 = c < NXD ? kd[0][c < NXD ? c : 0] : kq[0][c >= NXD ? c - NXD : 0];
                const double f1 = c < NXD ? f1d[c < NXD ? c : 0] : f1q[c >= NXD ? c - NXD : 0];
                const double sk = abstol + fabs(y) * reltol;
                s2 += ((f1 - f0) / sk) * ((f1 - f0) / sk);
            }
            const double d2 = sqrt(s2 / NX) / dt0;
            const double md = fmax(d1, d2);
            const double dt1 = md <= 1.0e-15 ? fmax(smalldt, dt0 * 1.0e-3) : pow(10.0, -(2.0 + log10(md)) / 5.0);
            dt = fmin(fmin(100.0 * dt0, dt1), dtmax);
        }
    }
    for (int iter = 0; t < t1; iter++) {
        dt = fmin(dt, dtmax);
        dt = fmin(dt, t1 - t);
        if (iter >= 100000 || !(dt > 0.0)) {   // OrdinaryDiffEq: retcode MaxIters (maxiters = 1e5) / DtLessThanMin
#pragma unroll
            for (int c = 0; c < NXD; c++) yd[c] = nan("");
            return;
        }
#pragma unroll
        for (int st = 1; st < 6; st++) {
#pragma unroll
            for (int c = 0; c < ND; c++) {
                double acc = A[st][0] * kd[0][c];
#pragma unroll
                for (int j = 1; j < st; j++) acc += A[st][j] * kd[j][c];
                td[c] = yd[c] + dt * acc;
            }
            rhs_all<Mdl, G>(td, p, F, kd[st], kq[st]);
        }
#pragma unroll
        for (int c = 0; c < ND; c++) {
            double acc = A[6][0] * kd[0][c];
#pragma unroll
            for (int j = 1; j < 6; j++) acc += A[6][j] * kd[j][c];
            nd[c] = yd[c] + dt * acc;
        }
#pragma unroll
        for (int c = 0; c < NQR; c++) {
            double acc = A[6][0] * kq[0][c];
#pragma unroll
            for (int j = 1; j < 6; j++) acc += A[6][j] * kq[j][c];
            nq[c] = yq[c] + dt * acc;
        }
        rhs_all<Mdl, G>(nd, p, F, kd[6], kq[6]);
        double se = 0.0;
#pragma unroll
        for (int c = 0; c < NX; c++) {
            double ut = 0.0;
#pragma unroll
            for (int j = 0; j < 7; j++) ut += BT[j] * (c < NXD ? kd[j][c < NXD ? c : 0] : kq[j][c >= NXD ? c - NXD : 0]);
            const double y0 = c < NXD ? yd[c < NXD ? c : 0] : yq[c >= NXD ? c - NXD : 0];
            const double y1 = c < NXD ? nd[c < NXD ? c : 0] : nq[c >= NXD ? c - NXD : 0];
            const double w = dt * ut / (abstol + fmax(fabs(y0), fabs(y1)) * reltol);
            se += w * w;
        }
        const double EEst = sqrt(se / NX);
        if (EEst != EEst) {   // non-finite: OrdinaryDiffEq stops with retcode Unstable; the status flag reports it
#pragma unroll
            for (int c = 0; c < NXD; c++) yd[c] = EEst;
            return;
        }
        double q, q11 = 0.0;
        if (EEst == 0.0) {
            q = 1.0 / qmax;
        } else {
            q11 = cb_fastpow(EEst, beta1);
            q = q11 / cb_fastpow(qold, beta2);
            q = fmax(1.0 / qmax, fmin(1.0 / qmin, q / gamma));
        }
        if (EEst <= 1.0) {
            qold = fmax(EEst, qoldinit);
            const double dtnew = dt / q;
            double tt = t + dt;
            const double big = fabs(t1);
            if (fabs(tt - t1) < 100.0 * (__longlong_as_double(__double_as_longlong(big) + 1) - big)) tt = t1;   // 100 eps(t1)
            if constexpr (SMP::ON)   // saveat times inside the accepted step: dense output, before the step's data is shifted
                n_emitted = emit_pending<ND, NQR, NQA>(smp, n_emitted, t, dt, tt, yd, yq, kd, kq);
            t = tt;
#pragma unroll
            for (int c = 0; c < ND; c++) {
                yd[c] = nd[c];
                kd[0][c] = kd[6][c];
            }
#pragma unroll
            for (int c = 0; c < NQR; c++) {
                yq[c] = nq[c];
                kq[0][c] = kq[6][c];
            }
            dt = fmin(dtmax, dtnew);
        } else {
            dt = dt / fmin(1.0 / qmin, q11 / gamma);
        }
    }
}

// Structural zeros of the sensitivities: the derivative with respect to a control value is identically zero until
// the first piece that uses it, so a piece only integrates the leading `na` directions of the thread's group (the
// host table gives na = 1 + index of the last direction that is already switched on; the variable order (u0, p,
// controls in time order) makes the live directions a prefix).  One instantiation per prefix length, selected by
// a warp-uniform branch.
template <class Mdl, int G, int METHOD, bool SQ, int GA = G>
__device__ __forceinline__ void integrate_active(int na, double *yd, double *yq, const double *p,
                                                 const typename Mdl::Force *F, const double *cf, int K)
{
    if constexpr (GA == 0) {
        integrate_piece<Mdl, 0, METHOD, SQ>(yd, yq, p, F, cf, K);
    } else {
        // instantiated prefix lengths: G, then every second one down to 4, then 3, 2, 1, 0.  The piece runs the
        // SMALLEST instantiated length >= na (a shorter one would drop live directions)
        constexpr int NEXT = GA > 4 ? GA - 2 : GA - 1;
        if (na > NEXT)
            integrate_piece<Mdl, GA, METHOD, SQ>(yd, yq, p, F, cf, K);
        else
            integrate_active<Mdl, G, METHOD, SQ, NEXT>(na, yd, yq, p, F, cf, K);
    }
}

// one merged-trajectory sample: [x(t); controls of the piece in layer row order] (src/single_shooting.jl:456)
template <class Mdl>
__device__ __forceinline__ void save_sample(const DevLayer &L, const ShootOut &o, long long sample, int nrow, long long b,
                                         const double *yd, const double *yq, const double *p)
{
    constexpr int NX = Mdl::NX, NXD = Mdl::NXD, NP = Mdl::NP;
    double *t = o.traj + (sample * nrow) * o.ldo + b;
#pragma unroll
    for (int k = 0; k < NX; k++) t[(long long)k * o.ldo] = (k < NXD) ? yd[k < NXD ? k : 0] : yq[k >= NXD ? k - NXD : 0];
    // control row r of a sample is fed from p slot row_to_p[r]; walking the p slots keeps every register
    // index static (pmap_idx[q] is the row of control slot q)
#pragma unroll
    for (int q = 0; q < NP; q++)
        if (L.pmap_kind[q] == 1) t[(long long)(NX + L.pmap_idx[q]) * o.ldo] = p[q];
}

// sensitivities of the thread's directions at one trajectory sample (WANT_TRAJ_SENS): what the point constraints of
// src/dynprob.jl:30-58 differentiate.  row0 = tsens_off[i] + j * ns * NX.
template <class Mdl, int G>
__device__ __forceinline__ void save_sens_sample(const ShootOut &o, long long row0, long long b, const int *dmap,
                                                 const double *yd, const double *yq)
{
    constexpr int NX = Mdl::NX, NXD = Mdl::NXD, NQ = Mdl::NQ;
#pragma unroll
    for (int g = 0; g < G; g++) {
        if (dmap[g] >= 0) {
#pragma unroll
            for (int k = 0; k < NX; k++)
                o.tsens[(row0 + (long long)dmap[g] * NX + k) * o.ldo + b] =
                    (k < NXD) ? yd[NXD * (1 + g) + (k < NXD ? k : 0)] : yq[NQ * (1 + g) + (k >= NXD ? k - NXD : 0)];
        }
    }
}

// MINB: resident blocks of 128 threads per SM the register allocation must allow (__launch_bounds__); chosen
// per (model, G) in the model's dispatch table from ptxas -v (largest value without spills in the step loop).
// TS: also store the sensitivities at every saved sample (WANT_TRAJ_SENS) -- a separate instantiation, so that the
// common case keeps its register allocation and instruction schedule.
// SV: the layer has saveat times inside pieces (instantiated for the models that ask for it, SaveAtTraits).
template <class Mdl, int G, int METHOD, bool UNI, bool TS, bool SV>
__device__ __forceinline__ void shoot_unit(const DevLayer &L, const double *__restrict__ ps, long long ldb, long long B,
                                           const ShootOut &o)
{
    constexpr int NX = Mdl::NX, NXD = Mdl::NXD, NQ = Mdl::NQ, NP = Mdl::NP;
    constexpr int ND = NXD * (1 + G), NQA = (NQ * (1 + G) > 0) ? NQ * (1 + G) : 1;
    constexpr int G1 = G > 0 ? G : 1;
    constexpr int NCF = METHOD == 0 ? 21 : 4;
    int r, i;
    long long b;
    if (!L.flat) {
        // regular launch: grid.x tiles the candidates, grid.y = interval, grid.z = direction group --
        // no integer divisions, and (i, r) is block-uniform
        // (lanes beyond B redo candidate B-1 without storing, so that warps stay whole for the shuffles)
        // (intervals in descending order: the threads of the LAST interval carry the fused Fisher read-out and the
        // criteria on top of their unit -- scheduled first, they do not form the kernel's tail)
        b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        i = L.i0 + (L.In - 1 - (int)blockIdx.y);
        r = blockIdx.z;
    } else {
        // flat launch for tiny batches (B < 32): lanes of a warp span intervals and direction groups
        const long long nsys = B * L.In;
        const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        if (tid >= nsys * L.R) return;
        r = (int)(tid / nsys);
        const long long rem = tid - r * nsys;
        const int il = (int)(rem / B);
        i = il + L.i0;
        b = rem - (long long)il * B;
    }

    const bool live = b < B;
    if (!live) b = B - 1;
    const int nu0 = __ldg(L.n_u0 + i), nc = __ldg(L.n_c + i);
    const int ns = nu0 + L.ntp + nc;
    const int d0 = r * G;
    if (r > 0 && d0 >= ns) return;
    const int voff = __ldg(L.var_off + i);
    const double *__restrict__ v = ps + (long long)voff * ldb + b;  // v[k*ldb]: local variable k

    // ---- initial state: InitialConditionRemaker (src/single_shooting.jl:260-266), fixval per :359;
    //      sensitivity seeds: e_k for the u0 direction that feeds state k
    // slot g of this thread holds direction dmap[g] of the interval: the host orders the directions by the piece
    // in which they become non-zero (u0 seeds on dynamic states and parameters first, controls in time order,
    // directions that never change -- seeds on quadrature-like states, unused controls -- last), so that the live
    // directions of every piece are a prefix of the group; -1: padding
    int dmap[G1];
#pragma unroll
    for (int g = 0; g < G1; g++) dmap[g] = -1;
    if constexpr (G > 0) {
#pragma unroll
        for (int g = 0; g < G; g++)
            if (d0 + g < ns) dmap[g] = (int)L.dir_perm[voff + d0 + g];
    }
    double yd[ND], yq[NQA];
#pragma unroll
    for (int k = 0; k < NX; k++) {
        const int src = __ldg(L.ic_src + i * NX + k);
        const double val = (src >= 0) ? v[(long long)src * ldb] : __ldg(L.ic_fix + i * NX + k);
        if (k < NXD)
            yd[k] = val;
        else
            yq[k - NXD] = val;
        if constexpr (G > 0) {
#pragma unroll
            for (int g = 0; g < G; g++) {
                const double seed = (src >= 0 && src == dmap[g]) ? 1.0 : 0.0;
                if (k < NXD)
                    yd[NXD * (1 + g) + k] = seed;
                else
                    yq[NQ * (1 + g) + (k - NXD)] = seed;
            }
        }
    }
    // p column each direction perturbs (static per interval: host table, -1 for u0 directions and padding)
    int pcol[G1];
#pragma unroll
    for (int g = 0; g < G1; g++) pcol[g] = -1;
    if constexpr (G > 0) {
#pragma unroll
        for (int g = 0; g < G; g++)
            if (dmap[g] >= 0) pcol[g] = (int)L.dir_pcol[voff + dmap[g]];
    }
    // ---- parameters: p_full = A*p + B*controls_j (src/single_shooting.jl:319-323)
    // G == 0 (the OED augmented systems: short units, many parameter slots): branch-free, every slot issues its load (a
    // control slot reads a dummy table entry), so all the unit's inputs are in flight together -- a uniform branch per slot
    // serialises one memory latency per parameter.  With sensitivity directions the branchy form keeps the tuned
    // register allocation of the step loop (LQR G = 4: 0.115 ms against 0.125 ms branch-free).
    constexpr bool BATCHED = (G == 0);
    double p[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) {
        if constexpr (BATCHED) {
            const bool tun = L.pmap_kind[q] == 0;
            const double val = __ldg(tun ? v + (long long)(nu0 + L.pmap_idx[q]) * ldb : L.ic_fix);
            p[q] = tun ? val : 0.0;
        } else {
            p[q] = (L.pmap_kind[q] == 0) ? v[(long long)(nu0 + L.pmap_idx[q]) * ldb] : 0.0;
        }
    }

    const int M = __ldg(L.M + i);
    const int po = __ldg(L.piece_off + i);
    const int so = __ldg(L.samp_off + i);
    const int nrow = NX + L.nact;
    const bool last_interval = (i == L.I - 1);
    const double *__restrict__ vc = v + (long long)(nu0 + L.ntp) * ldb;

    // controls of a piece: ps.controls[index_grid[:, j]] (src/single_shooting.jl:439).  The index and the
    // value it points to are two dependent global loads; they are issued one piece ahead, together with
    // the piece's forcing mask, so that their latency hides behind the integration of the current piece.
    double pv_nx[NP];
    unsigned mask_nx = 0;
    // (G == 0: first every index, then every value; a parameter slot reads a dummy entry)
    auto load_controls = [&](int piece) {
        if constexpr (BATCHED) {
            int cix[NP];
#pragma unroll
            for (int q = 0; q < NP; q++)
                cix[q] = (L.pmap_kind[q] == 1) ? __ldg(L.cidx + (long long)piece * L.nact + L.pmap_idx[q]) : -1;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                const double val = __ldg(cix[q] >= 0 ? vc + (long long)cix[q] * ldb : L.ic_fix);
                pv_nx[q] = cix[q] >= 0 ? val : 0.0;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NP; q++)
                if (L.pmap_kind[q] == 1) pv_nx[q] = vc[(long long)__ldg(L.cidx + (long long)piece * L.nact + L.pmap_idx[q]) * ldb];
        }
    };
#pragma unroll
    for (int q = 0; q < NP; q++) pv_nx[q] = 0.0;
    load_controls(po);
    if constexpr (G > 0) mask_nx = __ldg(L.act_mask + (long long)po * L.R + r);
    for (int j = 0; j < M; j++) {
#pragma unroll
        for (int q = 0; q < NP; q++) {
            if constexpr (BATCHED) {
                p[q] = (L.pmap_kind[q] == 1) ? pv_nx[q] : p[q];
            } else {
                if (L.pmap_kind[q] == 1) p[q] = pv_nx[q];
            }
        }
        const unsigned mask = mask_nx & 0xffffffu;
        const int na = (int)(mask_nx >> 24);   // live prefix of the direction group during this piece
        if (j + 1 < M) {
            load_controls(po + j + 1);
            if constexpr (G > 0) mask_nx = __ldg(L.act_mask + (long long)(po + j + 1) * L.R + r);
        }
        typename Mdl::Force F[G1];
        if constexpr (G > 0) {
#pragma unroll
            for (int g = 0; g < G; g++) F[g] = Mdl::make_force(pcol[g] | -(int)(((mask >> g) & 1u) ^ 1u), p);  // off -> -1
        }
        if (o.traj && live && r == 0 && j == 0)  // save_start = (j == 1)
            save_sample<Mdl>(L, o, (long long)so, nrow, b, yd, yq, p);
        if constexpr (G > 0 && TS) {
            if (o.tsens && live && j == 0) save_sens_sample<Mdl, G>(o, __ldg(L.tsens_off + i), b, dmap, yd, yq);
        }
        // (index of the piece's end sample within the interval: L.psamp[po + j], j + 1 without saveat -- loaded only where a
        // sample is written, the common path carries no extra load)
        bool sampled = false;
        if constexpr (SV && METHOD != 1) {
            const int nsv = L.sv_n ? __ldg(L.sv_n + po + j) : 0;
            if (nsv > 0) {   // warp-uniform
                const int pse = __ldg(L.psamp + po + j);
                sampled = true;
                PieceSampler<Mdl, G, TS> smp;
                smp.L = &L;
                smp.o = &o;
                smp.t = L.sv_t + __ldg(L.sv_off + po + j);
                smp.n = nsv;
                smp.b = b;
                smp.samp0 = (long long)so + pse - nsv;
                smp.ns = ns;
                smp.nrow = nrow;
                smp.tsrow0 = __ldg(L.tsens_off + i) + (long long)(pse - nsv) * ns * NX;
                smp.dmap = dmap;
                smp.p = p;
                smp.write_traj = o.traj && live && r == 0;
                smp.write_tsens = o.tsens && live;
                if constexpr (METHOD == 2)
                    integrate_piece_adaptive<Mdl, G>(yd, yq, p, F, __ldg(L.pt0 + po + j), __ldg(L.pt1 + po + j), L.abstol, L.reltol, smp);
                else
                    integrate_piece_dense<Mdl, G>(yd, yq, p, F, __ldg(L.pt0 + po + j), __ldg(L.pt1 + po + j), L.K, smp);
            }
        }
        if (sampled) {
            // done above
        } else if constexpr (METHOD == 2) {
            integrate_piece_adaptive<Mdl, G>(yd, yq, p, F, __ldg(L.pt0 + po + j), __ldg(L.pt1 + po + j), L.abstol, L.reltol);
        } else if constexpr (UNI) {
            if (r == 0)   // block-uniform: only the first direction group reports the state (and its quadratures)
                integrate_active<Mdl, G, METHOD, true>(na, yd, yq, p, F, L.hA, L.K);
            else
                integrate_active<Mdl, G, METHOD, false>(na, yd, yq, p, F, L.hA, L.K);
        } else {
            double cf[NCF];
            const double h = __ldg(L.ph + po + j);
            if constexpr (METHOD == 0)
                tsit5_coeffs(h, cf);
            else
                rk4_coeffs(h, cf);
            integrate_active<Mdl, G, (METHOD == 2 ? 0 : METHOD), true>(na, yd, yq, p, F, cf, L.K);
        }
        // save_end; the end sample of a non-final interval is dropped by the merge (multiple_shooting.jl:178-180)
        if (o.traj && live && r == 0 && (j + 1 < M || last_interval))
            save_sample<Mdl>(L, o, (long long)(so + (SV ? __ldg(L.psamp + po + j) : j + 1)), nrow, b, yd, yq, p);
        if constexpr (G > 0 && TS) {
            if (o.tsens && live && (j + 1 < M || last_interval))
                save_sens_sample<Mdl, G>(o, __ldg(L.tsens_off + i) + (long long)(SV ? __ldg(L.psamp + po + j) : j + 1) * ns * NX, b,
                                         dmap, yd, yq);
        }
    }
    // ---- results.  xs(k): end state component k (compile-time k)
#define CB200_XS(k) ((k) < NXD ? yd[(k) < NXD ? (k) : 0] : yq[(k) >= NXD ? (k)-NXD : 0])
#define CB200_SS(g, k) ((k) < NXD ? yd[NXD * (1 + (g)) + ((k) < NXD ? (k) : 0)] : yq[NQ * (1 + (g)) + ((k) >= NXD ? (k)-NXD : 0)])
    if (r == 0 && live) {
        if (o.status) {  // failure detection: the reference never inspects the solver's return code
            bool ok = true;
#pragma unroll
            for (int k = 0; k < NX; k++) ok = ok && isfinite(CB200_XS(k));
            if (!ok) atomicOr(o.status + b, 1);
        }
        if (o.xend) {
#pragma unroll
            for (int k = 0; k < NX; k++) o.xend[((long long)i * NX + k) * o.ldo + b] = CB200_XS(k);
        }
        if (o.objq) {  // objective = one state row of the last sample (src/dynprob.jl:70-71)
            double val = 0.0;
#pragma unroll
            for (int k = 0; k < NX; k++) val = (k == o.obj_state) ? CB200_XS(k) : val;
            o.objq[(long long)i * o.ldo + b] = val;
        }
        // shooting defects of the boundary (i, i+1), both orderings (src/multiple_shooting.jl:150-163, 229-295):
        // state matchings straight from the end state in registers, before any quadrature accumulation
        if ((o.dk || o.dst) && i + 1 < L.I) {
            const int vn = __ldg(L.var_off + i + 1);
            // all node values of the next interval first (NX independent loads in flight: one memory latency, not NX),
            // then the stores
            int pk[NX];
            double nxt[NX];
#pragma unroll
            for (int k = 0; k < NX; k++) {
                pk[k] = __ldg(L.dpos_kind + i * NX + k);
                const int src = __ldg(L.ic_src + (i + 1) * NX + k);
                nxt[k] = (pk[k] >= 0 && src >= 0) ? __ldg(ps + (long long)(vn + src) * ldb + b) : __ldg(L.ic_fix + (i + 1) * NX + k);
            }
#pragma unroll
            for (int k = 0; k < NX; k++) {
                if (pk[k] >= 0) {
                    const double dd = CB200_XS(k) - nxt[k];
                    if (o.dk) store_defect(o, pk[k], b, dd);
                    if (o.dst) o.dst[(long long)__ldg(L.dpos_stage + i * NX + k) * o.ldo + b] = dd;
                }
            }
            const int e1 = __ldg(L.oth_off + i + 1);
            for (int e = __ldg(L.oth_off + i); e < e1; e++) {  // parameter and control matchings
                const MatchEntry t = L.oth[e];
                const double dd = ps[(long long)t.a * ldb + b] - ps[(long long)t.b * ldb + b];
                if (o.dk) store_defect(o, t.pos_kind, b, dd);
                if (o.dst) o.dst[(long long)t.pos_stage * o.ldo + b] = dd;
            }
        }
    }
    // K4 fused: F = F_init + Symmetric(F_triu(t_end)) of the LAST interval (lib/CorleoneOED/src/oed.jl:388-398), its sum
    // over the candidates and the criteria, straight from the end state in registers (no second pass over xend)
    if constexpr (FimTraits<Mdl>::N > 0) {
        if (o.fim_fused && r == 0 && i == L.I - 1) {
            constexpr int FN = FimTraits<Mdl>::N, FO = FimTraits<Mdl>::OFF;
            double A[FN][FN];
#pragma unroll
            for (int c = 0; c < FN; c++)
#pragma unroll
                for (int rr = 0; rr < FN; rr++) {
                    const int lo = rr < c ? rr : c, hi = rr < c ? c : rr;
                    const int q = hi * (hi + 1) / 2 + lo;   // column-major triu position (augmentation.jl:239)
                    A[rr][c] = CB200_XS(FO + q) + (o.finit ? __ldg(o.finit + rr + FN * c) : 0.0);
                }
            fim_emit<FN>(A, FN, live, b, o.fim, o.ldo, o.fsum, o.crit, !L.flat);
        }
    }
    if constexpr (G > 0) {
        if (o.sens && live) {
            const long long s0 = __ldg(L.sens_off + i);
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int d = dmap[g];
                if (d >= 0) {
#pragma unroll
                    for (int k = 0; k < NX; k++) o.sens[(s0 + (long long)d * NX + k) * o.ldo + b] = CB200_SS(g, k);
                }
            }
        }
        // K5: constraint-Jacobian values.  d defect(i, k) / d (variable d of interval i) = S[k, d]; its slot in
        // cb200_jac_structure order is jpos[i, k] + d, so the Jacobian is written without a gather pass
        if (o.jac && live && i + 1 < L.I) {
#pragma unroll
            for (int k = 0; k < NX; k++) {
                const int jp = __ldg(L.jpos + i * NX + k);
                if (jp >= 0) {
#pragma unroll
                    for (int g = 0; g < G; g++)
                        if (dmap[g] >= 0) o.jac[(long long)(jp + dmap[g]) * o.ldo + b] = CB200_SS(g, k);
                }
            }
        }
        // objective gradient: variable voff + d is direction d of this interval, so every thread owns the
        // gradient entries of its directions (zero when the interval does not reach the objective)
        if (o.grad || o.grad_sum || o.gradc) {
            const bool inc = o.obj_all || i == L.I - 1;
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int d = dmap[g];
                if (d >= 0) {
                    double val = 0.0;
#pragma unroll
                    for (int k = 0; k < NX; k++) val = (inc && k == o.obj_state) ? CB200_SS(g, k) : val;
                    if (o.grad && live) o.grad[(long long)(voff + d) * o.ldo + b] = val;
                    if (o.gradc && live && voff + d >= o.gradc_first) o.gradc[(long long)(voff + d - o.gradc_first) * o.ldo + b] = val;
                    if (o.grad_sum) {
                        if (!live) val = 0.0;
                        if (!L.flat) {  // whole warp shares (i, r): shuffle tree, one atomic per warp
#pragma unroll
                            for (int w = 16; w > 0; w >>= 1) val += __shfl_down_sync(0xffffffffu, val, w);
                            if ((threadIdx.x & 31) == 0) atomicAdd(o.grad_sum + voff + d, val);
                        } else {
                            atomicAdd(o.grad_sum + voff + d, val);
                        }
                    }
                }
            }
        }
    }
#undef CB200_XS
#undef CB200_SS
}

// Tail of a SMALL evaluation (a few candidates, flat launch; BASELINE config 1): the last block to finish computes the
// objective (the ordered sum over the intervals, src/multiple_shooting.jl:171-177, src/dynprob.jl:70-71) and copies the whole
// result block to the caller's pinned mirror -- the evaluation is ONE kernel instead of kernel + objective kernel + copy node
// (tools/latency_micro.cu: 8.3 us around the kernel instead of 17.7).
__device__ __forceinline__ void small_tail(const DevLayer &L, const ShootOut &o, long long B)
{
    __shared__ int is_last;
    __threadfence();   // this thread's results are visible before the block is counted
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned prev = atomicAdd(o.st_counter, 1u);
        is_last = (prev == total - 1);
        if (is_last) *o.st_counter = 0;   // zero at rest
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (o.st_obj) {
        for (long long b = threadIdx.x; b < B; b += blockDim.x) {
            double v;
            if (o.obj_all) {
                v = __ldcg(o.objq + b);
                for (int i = 1; i < L.I; i++) v = __ldcg(o.objq + (long long)i * o.ldo + b) + v;
            } else {
                v = __ldcg(o.objq + (long long)(L.I - 1) * o.ldo + b);
            }
            o.st_obj[b] = v;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < o.st_n; j += blockDim.x) o.st_dst[j] = __ldcg(o.st_src + j);
    __threadfence_system();
}

// FT: the kernel carries the small-evaluation tail (a separate instantiation: the throughput kernels stay as they are)
template <class Mdl, int G, int METHOD, bool UNI, int MINB, bool TS, bool SV = false, bool FT = false>
__global__ void __launch_bounds__(128, MINB)
shoot_kernel(const __grid_constant__ DevLayer L, const double *__restrict__ ps, long long ldb, long long B, const ShootOut o)
{
    shoot_unit<Mdl, G, METHOD, UNI, TS, SV>(L, ps, ldb, B, o);
    if constexpr (FT) small_tail(L, o, B);
}

// Forward initialisation sweep (src/node_initialization.jl:152-170): one thread per candidate walks the intervals
// in order, overwrites the node values ps.interval_i.u0 with the leading entries of the previous interval's end
// sample (unless the interval is listed in fixed_indices) and integrates on.  As in the reference the sweep calls
// the layer WITHOUT the shooting flag, so non-tunable states restart from problem.u0 (table ic_fix0), not from 0.
template <class Mdl, int METHOD>
__global__ void __launch_bounds__(128)
forward_kernel(const __grid_constant__ DevLayer L, double *__restrict__ ps, long long ldb, long long B,
               const unsigned char *__restrict__ fixed)
{
    constexpr int NX = Mdl::NX, NXD = Mdl::NXD, NQ = Mdl::NQ, NP = Mdl::NP;
    constexpr int NQA = NQ > 0 ? NQ : 1, NCF = METHOD == 0 ? 21 : 4;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    double carry[NX];
#pragma unroll
    for (int k = 0; k < NX; k++) carry[k] = 0.0;
    for (int i = 0; i < L.I; i++) {
        const int nu0 = __ldg(L.n_u0 + i), voff = __ldg(L.var_off + i);
        double *v = ps + (long long)voff * ldb + b;
        const bool keep = (i == 0) || fixed[i];
        if (!keep) {  // sps.u0 .= u0[eachindex(sps.u0)]
#pragma unroll
            for (int m = 0; m < NX; m++)
                if (m < nu0) v[(long long)m * ldb] = carry[m];
        }
        double yd[NXD], yq[NQA];
#pragma unroll
        for (int k = 0; k < NX; k++) {
            const int src = __ldg(L.ic_src + i * NX + k);
            double val = __ldg(L.ic_fix0 + i * NX + k);
            if (src >= 0) {
                if (keep) {
                    val = v[(long long)src * ldb];
                } else {
#pragma unroll
                    for (int m = 0; m < NX; m++) val = (m == src) ? carry[m] : val;
                }
            }
            if (k < NXD)
                yd[k < NXD ? k : 0] = val;
            else
                yq[k >= NXD ? k - NXD : 0] = val;
        }
        double p[NP];
#pragma unroll
        for (int q = 0; q < NP; q++) p[q] = (L.pmap_kind[q] == 0) ? v[(long long)(nu0 + L.pmap_idx[q]) * ldb] : 0.0;
        const int M = __ldg(L.M + i), po = __ldg(L.piece_off + i);
        const double *vc = v + (long long)(nu0 + L.ntp) * ldb;
        for (int j = 0; j < M; j++) {
#pragma unroll
            for (int q = 0; q < NP; q++)
                if (L.pmap_kind[q] == 1)
                    p[q] = vc[(long long)__ldg(L.cidx + (long long)(po + j) * L.nact + L.pmap_idx[q]) * ldb];
            double cf[NCF];
            if constexpr (METHOD != 2) {
                if (L.uniform_h) {
#pragma unroll
                    for (int c = 0; c < NCF; c++) cf[c] = L.hA[c];
                } else {
                    const double h = __ldg(L.ph + po + j);
                    if constexpr (METHOD == 0)
                        tsit5_coeffs(h, cf);
                    else
                        rk4_coeffs(h, cf);
                }
            }
            typename Mdl::Force F[1];
            if constexpr (METHOD == 2)
                integrate_piece_adaptive<Mdl, 0>(yd, yq, p, F, __ldg(L.pt0 + po + j), __ldg(L.pt1 + po + j), L.abstol, L.reltol);
            else
                integrate_piece<Mdl, 0, METHOD>(yd, yq, p, F, cf, L.K);
        }
#pragma unroll
        for (int k = 0; k < NX; k++) carry[k] = (k < NXD) ? yd[k < NXD ? k : 0] : yq[k >= NXD ? k - NXD : 0];
    }
}

template <class Mdl, bool ADAPTIVE = false>
int launch_forward(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
                   cudaStream_t s)
{
    if (B <= 0) return 0;
    const unsigned grid = (unsigned)((B + 63) / 64);
    if (method == 0)
        forward_kernel<Mdl, 0><<<grid, 64, 0, s>>>(L, ps, ldb, B, fixed);
    else if (method == 1)
        forward_kernel<Mdl, 1><<<grid, 64, 0, s>>>(L, ps, ldb, B, fixed);
    else {
        if constexpr (ADAPTIVE)
            forward_kernel<Mdl, 2><<<grid, 64, 0, s>>>(L, ps, ldb, B, fixed);
        else
            return -1000;   // the adaptive instantiation of this model is not compiled in
    }
    return (int)cudaGetLastError();
}

// threads per block of the shooting kernel (<= 128, the __launch_bounds__); small blocks shorten the tail
// wave of register-heavy instantiations.  CB200_BLOCK overrides for tuning.
inline int launch_block_size()
{
    static const int bs = [] {
        const char *e = getenv("CB200_BLOCK");
        const int v = e ? atoi(e) : 0;
        return (v == 32 || v == 64 || v == 96 || v == 128) ? v : 64;
    }();
    return bs;
}

template <class Mdl, int G, int METHOD, int MINB>
int launch_one(const DevLayer &L, const double *ps, long long ldb, long long B, const ShootOut &o, cudaStream_t s)
{
    const long long nthreads = B * L.In * (long long)L.R;
    if (nthreads <= 0) return 0;
    const int bs = launch_block_size();
    DevLayer L2 = L;
    dim3 grid;
    L2.flat = (B < 32 || L.In > 65535 || L.R > 65535) ? 1 : 0;
    if (L2.flat)
        grid = dim3((unsigned)((nthreads + bs - 1) / bs), 1, 1);
    else
        grid = dim3((unsigned)((B + bs - 1) / bs), (unsigned)L.In, (unsigned)L.R);
    if (o.st_counter) {   // small evaluation with the fused tail: for the models that carry that instantiation
        if constexpr (SmallTailTraits<Mdl>::ON && METHOD != 2) {
            if (L.sv_n || o.tsens || !L2.flat) return -1001;
            if (L.uniform_h)
                shoot_kernel<Mdl, G, METHOD, true, 1, false, false, true><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
            else
                shoot_kernel<Mdl, G, METHOD, false, 1, false, false, true><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
            return (int)cudaGetLastError();
        } else {
            return -1001;
        }
    }
    if (L.sv_n) {   // saveat samples inside pieces: the dense-output instantiation, for the models that carry one
        if constexpr (SaveAtTraits<Mdl>::ON && METHOD != 1) {
            shoot_kernel<Mdl, G, METHOD, false, 1, (G > 0), true><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
            return (int)cudaGetLastError();
        } else {
            return -1000;
        }
    }
    if constexpr (G > 0) {
        if (o.tsens) {
            if constexpr (METHOD != 2) {
                if (L.uniform_h) {
                    shoot_kernel<Mdl, G, METHOD, true, 1, true><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
                    return (int)cudaGetLastError();
                }
            }
            shoot_kernel<Mdl, G, METHOD, false, 1, true><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
            return (int)cudaGetLastError();
        }
    }
    if constexpr (METHOD != 2) {
        if (L.uniform_h) {
            shoot_kernel<Mdl, G, METHOD, true, MINB, false><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
            return (int)cudaGetLastError();
        }
    }
    shoot_kernel<Mdl, G, METHOD, false, MINB, false><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
    return (int)cudaGetLastError();
}

#ifdef CB200_MINB   // tuning builds: one value for every instantiation
#define CB200_MB(v) CB200_MINB
#else
#define CB200_MB(v) v
#endif
#define CB200_DISPATCH_G(Mdl, Gv, MBv)                                                      \
    case Gv:                                                                                \
        if (method > 1) return -1000;   /* no adaptive instantiation of this model */       \
        return method == 0 ? launch_one<Mdl, Gv, 0, CB200_MB(MBv)>(L, ps, ldb, B, o, s)   \
                           : launch_one<Mdl, Gv, 1, CB200_MB(MBv)>(L, ps, ldb, B, o, s);
// ... plus method 2, adaptive Tsit5 (128-register budget: 0.537 vs 0.556 ms at 255 registers, profiles/README.md)
#define CB200_DISPATCH_GA(Mdl, Gv, MBv)                                                     \
    case Gv:                                                                                \
        return method == 0   ? launch_one<Mdl, Gv, 0, CB200_MB(MBv)>(L, ps, ldb, B, o, s) \
               : method == 1 ? launch_one<Mdl, Gv, 1, CB200_MB(MBv)>(L, ps, ldb, B, o, s) \
                             : launch_one<Mdl, Gv, 2, 4>(L, ps, ldb, B, o, s);

}  // namespace cb200
