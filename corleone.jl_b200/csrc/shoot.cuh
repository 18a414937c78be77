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
#include <cstdlib>

namespace cb200 {

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

// Tsit5 coefficient slots: cf[0]=h a21; [1..2]=h a31,a32; [3..5]=h a4.; [6..9]=h a5.; [10..14]=h a6.; [15..20]=h a7.
__host__ __device__ inline void tsit5_coeffs(double h, double *cf)
{
    using namespace ts5;
    const double a[21] = {a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65,
                          a71, a72, a73, a74, a75, a76};
    for (int k = 0; k < 21; k++) cf[k] = h * a[k];
}
// RK4 coefficient slots: cf[0]=h/2, cf[1]=h, cf[2]=h/6, cf[3]=h/3
__host__ __device__ inline void rk4_coeffs(double h, double *cf)
{
    cf[0] = 0.5 * h;
    cf[1] = h;
    cf[2] = h / 6.0;
    cf[3] = h / 3.0;
}

template <class Mdl, int G, int METHOD>
__device__ __forceinline__ void integrate_piece(double *yd, double *yq, const double *p, const typename Mdl::Force *F,
                                                const double *cf, int K)
{
    constexpr int ND = Mdl::NXD * (1 + G), NQ = (Mdl::NQ * (1 + G) > 0) ? Mdl::NQ * (1 + G) : 1;
    constexpr int NQR = Mdl::NQ * (1 + G);
    double kq[NQ];
    if constexpr (METHOD == 0) {
        double k1[ND], k2[ND], k3[ND], k4[ND], k5[ND], k6[ND], td[ND];
        rhs_all<Mdl, G>(yd, p, F, k1, kq);
#pragma unroll
        for (int c = 0; c < NQR; c++) yq[c] = fma(cf[15], kq[c], yq[c]);
        for (int s = 0; s < K; s++) {
#pragma unroll
            for (int c = 0; c < ND; c++) td[c] = fma(cf[0], k1[c], yd[c]);
            rhs_all<Mdl, G>(td, p, F, k2, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[16], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) td[c] = fma(cf[2], k2[c], fma(cf[1], k1[c], yd[c]));
            rhs_all<Mdl, G>(td, p, F, k3, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[17], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) td[c] = fma(cf[5], k3[c], fma(cf[4], k2[c], fma(cf[3], k1[c], yd[c])));
            rhs_all<Mdl, G>(td, p, F, k4, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[18], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++)
                td[c] = fma(cf[9], k4[c], fma(cf[8], k3[c], fma(cf[7], k2[c], fma(cf[6], k1[c], yd[c]))));
            rhs_all<Mdl, G>(td, p, F, k5, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[19], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++)
                td[c] = fma(cf[14], k5[c],
                            fma(cf[13], k4[c], fma(cf[12], k3[c], fma(cf[11], k2[c], fma(cf[10], k1[c], yd[c])))));
            rhs_all<Mdl, G>(td, p, F, k6, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[20], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++)
                yd[c] = fma(cf[20], k6[c],
                            fma(cf[19], k5[c],
                                fma(cf[18], k4[c], fma(cf[17], k3[c], fma(cf[16], k2[c], fma(cf[15], k1[c], yd[c]))))));
            if (s + 1 < K) {  // FSAL: k7 of this step is k1 of the next; after the last step it is never used
                rhs_all<Mdl, G>(yd, p, F, k1, kq);
#pragma unroll
                for (int c = 0; c < NQR; c++) yq[c] = fma(cf[15], kq[c], yq[c]);
            }
        }
    } else {
        double k[ND], acc[ND], td[ND];
        for (int s = 0; s < K; s++) {
            rhs_all<Mdl, G>(yd, p, F, k, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[2], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                acc[c] = fma(cf[2], k[c], yd[c]);
                td[c] = fma(cf[0], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(td, p, F, k, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[3], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                acc[c] = fma(cf[3], k[c], acc[c]);
                td[c] = fma(cf[0], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(td, p, F, k, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[3], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) {
                acc[c] = fma(cf[3], k[c], acc[c]);
                td[c] = fma(cf[1], k[c], yd[c]);
            }
            rhs_all<Mdl, G>(td, p, F, k, kq);
#pragma unroll
            for (int c = 0; c < NQR; c++) yq[c] = fma(cf[2], kq[c], yq[c]);
#pragma unroll
            for (int c = 0; c < ND; c++) yd[c] = fma(cf[2], k[c], acc[c]);
        }
    }
}

template <class Mdl, int G, int METHOD, bool UNI>
__global__ void __launch_bounds__(128)
shoot_kernel(const __grid_constant__ DevLayer L, const double *__restrict__ ps, long long ldb, long long B, const ShootOut o)
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
        b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        i = blockIdx.y;
        r = blockIdx.z;
        if (b >= B) return;
    } else {
        // flat launch for tiny batches (B < 32): lanes of a warp span intervals and direction groups
        const long long nsys = B * L.I;
        const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
        if (tid >= nsys * L.R) return;
        r = (int)(tid / nsys);
        const long long rem = tid - r * nsys;
        i = (int)(rem / B);
        b = rem - (long long)i * B;
    }

    const int nu0 = __ldg(L.n_u0 + i), nc = __ldg(L.n_c + i);
    const int ns = nu0 + L.ntp + nc;
    const int d0 = r * G;
    if (r > 0 && d0 >= ns) return;
    const double *__restrict__ v = ps + (long long)__ldg(L.var_off + i) * ldb + b;  // v[k*ldb]: local variable k

    // ---- initial state: InitialConditionRemaker (src/single_shooting.jl:260-266), fixval per :359;
    //      sensitivity seeds: e_k for the u0 direction that feeds state k
    double yd[ND], yq[NQA];
    int pcol[G1], cloc[G1];
#pragma unroll
    for (int g = 0; g < G1; g++) pcol[g] = cloc[g] = -1;
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
                const double seed = (src >= 0 && src == d0 + g) ? 1.0 : 0.0;
                if (k < NXD)
                    yd[NXD * (1 + g) + k] = seed;
                else
                    yq[NQ * (1 + g) + (k - NXD)] = seed;
            }
        }
    }
    if constexpr (G > 0) {
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int d = d0 + g;
            if (d >= nu0 && d < nu0 + L.ntp)
                pcol[g] = L.tunable_p0[d - nu0];
            else if (d >= nu0 + L.ntp && d < ns)
                cloc[g] = d - nu0 - L.ntp;
        }
    }
    // ---- parameters: p_full = A*p + B*controls_j (src/single_shooting.jl:319-323)
    double p[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) p[q] = (L.pmap_kind[q] == 0) ? v[(long long)(nu0 + L.pmap_idx[q]) * ldb] : 0.0;

    const int M = __ldg(L.M + i);
    const int po = __ldg(L.piece_off + i);
    const int so = __ldg(L.samp_off + i);
    const int nrow = NX + L.nact;
    const bool last_interval = (i == L.I - 1);
    const double *__restrict__ vc = v + (long long)(nu0 + L.ntp) * ldb;

    // controls of a piece: ps.controls[index_grid[:, j]] (src/single_shooting.jl:439).  The index and the
    // value it points to are two dependent global loads; they are issued one piece ahead so that their
    // latency hides behind the integration of the current piece.
    int ci_nx[NP];
    double cv_nx[NP];
#pragma unroll
    for (int rr = 0; rr < NP; rr++) {
        ci_nx[rr] = -1;
        cv_nx[rr] = 0.0;
        if (rr < L.nact) {
            ci_nx[rr] = __ldg(L.cidx + (long long)po * L.nact + rr);
            cv_nx[rr] = vc[(long long)ci_nx[rr] * ldb];
        }
    }
    for (int j = 0; j < M; j++) {
        int pc[G1];
#pragma unroll
        for (int g = 0; g < G1; g++) pc[g] = pcol[g];
        double cval[NP];
#pragma unroll
        for (int rr = 0; rr < NP; rr++) {
            cval[rr] = cv_nx[rr];
            if (rr < L.nact) {
                const int ci = ci_nx[rr];
#pragma unroll
                for (int q = 0; q < NP; q++)
                    if (L.row_to_p[rr] == q) p[q] = cval[rr];
                if constexpr (G > 0) {
#pragma unroll
                    for (int g = 0; g < G; g++)
                        if (cloc[g] == ci) pc[g] = L.row_to_p[rr];
                }
                if (j + 1 < M) {
                    ci_nx[rr] = __ldg(L.cidx + (long long)(po + j + 1) * L.nact + rr);
                    cv_nx[rr] = vc[(long long)ci_nx[rr] * ldb];
                }
            }
        }
        typename Mdl::Force F[G1];
        if constexpr (G > 0) {
#pragma unroll
            for (int g = 0; g < G; g++) F[g] = Mdl::make_force(pc[g], p);
        }
        if (o.traj && r == 0 && j == 0) {  // save_start = (j == 1)
            double *t = o.traj + ((long long)so * nrow) * o.ldo + b;
#pragma unroll
            for (int k = 0; k < NX; k++) t[(long long)k * o.ldo] = (k < NXD) ? yd[k < NXD ? k : 0] : yq[k >= NXD ? k - NXD : 0];
#pragma unroll
            for (int rr = 0; rr < NP; rr++)
                if (rr < L.nact) t[(long long)(NX + rr) * o.ldo] = cval[rr];
        }
        if constexpr (UNI) {
            integrate_piece<Mdl, G, METHOD>(yd, yq, p, F, L.hA, L.K);
        } else {
            double cf[NCF];
            const double h = __ldg(L.ph + po + j);
            if constexpr (METHOD == 0)
                tsit5_coeffs(h, cf);
            else
                rk4_coeffs(h, cf);
            integrate_piece<Mdl, G, METHOD>(yd, yq, p, F, cf, L.K);
        }
        // save_end; the end sample of a non-final interval is dropped by the merge (multiple_shooting.jl:178-180)
        if (o.traj && r == 0 && (j + 1 < M || last_interval)) {
            double *t = o.traj + ((long long)(so + j + 1) * nrow) * o.ldo + b;
#pragma unroll
            for (int k = 0; k < NX; k++) t[(long long)k * o.ldo] = (k < NXD) ? yd[k < NXD ? k : 0] : yq[k >= NXD ? k - NXD : 0];
#pragma unroll
            for (int rr = 0; rr < NP; rr++)
                if (rr < L.nact) t[(long long)(NX + rr) * o.ldo] = cval[rr];
        }
    }
    if (o.xend && r == 0) {
#pragma unroll
        for (int k = 0; k < NX; k++)
            o.xend[((long long)i * NX + k) * o.ldo + b] = (k < NXD) ? yd[k < NXD ? k : 0] : yq[k >= NXD ? k - NXD : 0];
    }
    if constexpr (G > 0) {
        if (o.sens) {
            const long long s0 = __ldg(L.sens_off + i);
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int d = d0 + g;
                if (d < ns) {
#pragma unroll
                    for (int k = 0; k < NX; k++)
                        o.sens[(s0 + (long long)d * NX + k) * o.ldo + b] =
                            (k < NXD) ? yd[NXD * (1 + g) + (k < NXD ? k : 0)] : yq[NQ * (1 + g) + (k >= NXD ? k - NXD : 0)];
                }
            }
        }
    }
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

template <class Mdl, int G, int METHOD>
int launch_one(const DevLayer &L, const double *ps, long long ldb, long long B, const ShootOut &o, cudaStream_t s)
{
    const long long nthreads = B * L.I * (long long)L.R;
    if (nthreads <= 0) return 0;
    const int bs = launch_block_size();
    DevLayer L2 = L;
    dim3 grid;
    L2.flat = (B < 32 || L.I > 65535 || L.R > 65535) ? 1 : 0;
    if (L2.flat)
        grid = dim3((unsigned)((nthreads + bs - 1) / bs), 1, 1);
    else
        grid = dim3((unsigned)((B + bs - 1) / bs), (unsigned)L.I, (unsigned)L.R);
    if (L.uniform_h)
        shoot_kernel<Mdl, G, METHOD, true><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
    else
        shoot_kernel<Mdl, G, METHOD, false><<<grid, bs, 0, s>>>(L2, ps, ldb, B, o);
    return (int)cudaGetLastError();
}

#define CB200_DISPATCH_G(Mdl, Gv)                                                           \
    case Gv:                                                                                \
        return method == 0 ? launch_one<Mdl, Gv, 0>(L, ps, ldb, B, o, s)                   \
                           : launch_one<Mdl, Gv, 1>(L, ps, ldb, B, o, s);

}  // namespace cb200
