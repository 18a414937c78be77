// K1/K2: batched integration of every (interval, candidate) shooting unit with forward sensitivities.
//
// Replaces mythreadmap over intervals (src/Corleone.jl:22-24, src/multiple_shooting.jl:118-130) of the
// piece loop _sequential_solve (src/single_shooting.jl:421-471), and the ForwardDiff sweep through it
// (src/dynprob.jl:150,161), for B candidates at once.
//
// Mapping: one thread = (candidate b, interval i, direction group r).  The thread integrates the state x
// and G sensitivity directions d = r*G .. r*G+G-1 of its unit, everything in registers (NZ = NX*(1+G)
// doubles per stage vector, fully unrolled).  Directions are the local variables of the interval in
// reference order (u0, p, controls) (src/single_shooting.jl:190-200), so `sens` is written directly in
// Jacobian-block order.  b is the fastest thread index: with the SoA layout [variable][candidate] every
// global load and store of a warp is one coalesced 256-byte transaction, and (i, r) -- hence the
// direction kind and every branch on it -- is warp-uniform whenever B is a multiple of 32.
//
// Stepping: fixed-step Tsit5 (7-stage FSAL; k1 re-evaluated at each piece start because every control
// piece is a fresh `solve` in the reference, src/single_shooting.jl:443-453) or classical RK4; forward
// sensitivities = the same RK scheme applied to the variational system S' = f_x S + f_p (what
// ForwardDiff through a fixed-step solve computes).
#pragma once
#include "common.cuh"
#include "tableau.cuh"

namespace cb200 {

template <class Mdl, int G>
__device__ __forceinline__ void rhs_all(const double *y, const double *p, const int *pcol, double *ky)
{
    constexpr int NX = Mdl::NX;
    Mdl::f(y, p, ky);
    if constexpr (G > 0) {
#pragma unroll
        for (int g = 0; g < G; g++) Mdl::jvp(y, p, y + NX * (1 + g), pcol[g], ky + NX * (1 + g));
    }
}

template <class Mdl, int G, int METHOD>
__device__ __forceinline__ void integrate_piece(double (&y)[Mdl::NX * (1 + G)], const double *p, const int *pcol,
                                                double h, int K)
{
    constexpr int NZ = Mdl::NX * (1 + G);
    if constexpr (METHOD == 0) {
        using namespace ts5;
        double k1[NZ], k2[NZ], k3[NZ], k4[NZ], k5[NZ], k6[NZ], tmp[NZ];
        rhs_all<Mdl, G>(y, p, pcol, k1);
        for (int s = 0; s < K; s++) {
#pragma unroll
            for (int c = 0; c < NZ; c++) tmp[c] = fma(h * a21, k1[c], y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k2);
#pragma unroll
            for (int c = 0; c < NZ; c++) tmp[c] = fma(h, fma(a32, k2[c], a31 * k1[c]), y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k3);
#pragma unroll
            for (int c = 0; c < NZ; c++) tmp[c] = fma(h, fma(a43, k3[c], fma(a42, k2[c], a41 * k1[c])), y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k4);
#pragma unroll
            for (int c = 0; c < NZ; c++)
                tmp[c] = fma(h, fma(a54, k4[c], fma(a53, k3[c], fma(a52, k2[c], a51 * k1[c]))), y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k5);
#pragma unroll
            for (int c = 0; c < NZ; c++)
                tmp[c] = fma(h, fma(a65, k5[c], fma(a64, k4[c], fma(a63, k3[c], fma(a62, k2[c], a61 * k1[c])))),
                             y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k6);
#pragma unroll
            for (int c = 0; c < NZ; c++)
                y[c] = fma(h,
                           fma(a76, k6[c],
                               fma(a75, k5[c], fma(a74, k4[c], fma(a73, k3[c], fma(a72, k2[c], a71 * k1[c]))))),
                           y[c]);
            rhs_all<Mdl, G>(y, p, pcol, k1);  // k7 of this step == k1 of the next (FSAL)
        }
    } else {
        double k1[NZ], k2[NZ], k3[NZ], k4[NZ], tmp[NZ];
        for (int s = 0; s < K; s++) {
            rhs_all<Mdl, G>(y, p, pcol, k1);
#pragma unroll
            for (int c = 0; c < NZ; c++) tmp[c] = fma(0.5 * h, k1[c], y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k2);
#pragma unroll
            for (int c = 0; c < NZ; c++) tmp[c] = fma(0.5 * h, k2[c], y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k3);
#pragma unroll
            for (int c = 0; c < NZ; c++) tmp[c] = fma(h, k3[c], y[c]);
            rhs_all<Mdl, G>(tmp, p, pcol, k4);
#pragma unroll
            for (int c = 0; c < NZ; c++)
                y[c] = fma(h / 6.0, k4[c], fma(h / 3.0, k3[c], fma(h / 3.0, k2[c], fma(h / 6.0, k1[c], y[c]))));
        }
    }
}

template <class Mdl, int G, int METHOD>
__global__ void __launch_bounds__(128)
shoot_kernel(const DevLayer L, const double *__restrict__ ps, long long ldb, long long B, const ShootOut o)
{
    constexpr int NX = Mdl::NX, NP = Mdl::NP, NZ = NX * (1 + G);
    const long long nsys = B * L.I;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (tid >= nsys * L.R) return;
    const int r = (int)(tid / nsys);
    const long long rem = tid - r * nsys;
    const int i = (int)(rem / B);
    const long long b = rem - (long long)i * B;

    const int nu0 = __ldg(L.n_u0 + i), nc = __ldg(L.n_c + i);
    const int ns = nu0 + L.ntp + nc;
    const int d0 = r * G;
    if (r > 0 && d0 >= ns) return;
    const double *__restrict__ v = ps + (long long)__ldg(L.var_off + i) * ldb + b;  // v[k*ldb]: local variable k

    // ---- initial state: InitialConditionRemaker (src/single_shooting.jl:260-266), fixval per :359
    double y[NZ];
#pragma unroll
    for (int k = 0; k < NX; k++) {
        const int src = __ldg(L.ic_src + i * NX + k);
        y[k] = (src >= 0) ? v[(long long)src * ldb] : __ldg(L.ic_fix + i * NX + k);
    }
    // ---- sensitivity seeds and direction kinds
    int pcol[G > 0 ? G : 1];
    int cloc[G > 0 ? G : 1];
    if constexpr (G > 0) {
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int d = d0 + g;
            pcol[g] = -1;
            cloc[g] = -1;
#pragma unroll
            for (int k = 0; k < NX; k++)
                y[NX * (1 + g) + k] = (d < nu0 && __ldg(L.ic_src + i * NX + k) == d) ? 1.0 : 0.0;
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

    for (int j = 0; j < M; j++) {
        // controls of this piece: ps.controls[index_grid[:, j]] (src/single_shooting.jl:439)
        int pc[G > 0 ? G : 1];
#pragma unroll
        for (int g = 0; g < (G > 0 ? G : 1); g++) pc[g] = (G > 0) ? pcol[g] : -1;
        double cval[MAX_CTRL];
#pragma unroll
        for (int rr = 0; rr < MAX_CTRL; rr++) {
            if (rr < L.nact) {
                const int ci = __ldg(L.cidx + (long long)(po + j) * L.nact + rr);
                cval[rr] = vc[(long long)ci * ldb];
                if constexpr (G > 0) {
#pragma unroll
                    for (int g = 0; g < G; g++)
                        if (cloc[g] == ci) pc[g] = L.row_to_p[rr];
                }
            } else {
                cval[rr] = 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < NP; q++)
            if (L.pmap_kind[q] == 1) {
                double val = 0.0;
#pragma unroll
                for (int rr = 0; rr < MAX_CTRL; rr++)
                    if (L.pmap_idx[q] == rr) val = cval[rr];
                p[q] = val;
            }
        if (o.traj && r == 0 && j == 0) {  // save_start = (j == 1)
            double *t = o.traj + ((long long)so * nrow) * o.ldo + b;
#pragma unroll
            for (int k = 0; k < NX; k++) t[(long long)k * o.ldo] = y[k];
#pragma unroll
            for (int rr = 0; rr < MAX_CTRL; rr++)
                if (rr < L.nact) t[(long long)(NX + rr) * o.ldo] = cval[rr];
        }
        const double t0 = __ldg(L.pt0 + po + j), t1 = __ldg(L.pt1 + po + j);
        integrate_piece<Mdl, G, METHOD>(y, p, pc, (t1 - t0) / (double)L.K, L.K);
        // save_end; the end sample of a non-final interval is dropped by the merge (multiple_shooting.jl:178-180)
        if (o.traj && r == 0 && (j + 1 < M || last_interval)) {
            double *t = o.traj + ((long long)(so + j + 1) * nrow) * o.ldo + b;
#pragma unroll
            for (int k = 0; k < NX; k++) t[(long long)k * o.ldo] = y[k];
#pragma unroll
            for (int rr = 0; rr < MAX_CTRL; rr++)
                if (rr < L.nact) t[(long long)(NX + rr) * o.ldo] = cval[rr];
        }
    }
    if (o.xend && r == 0) {
#pragma unroll
        for (int k = 0; k < NX; k++) o.xend[((long long)i * NX + k) * o.ldo + b] = y[k];
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
                        o.sens[(s0 + (long long)d * NX + k) * o.ldo + b] = y[NX * (1 + g) + k];
                }
            }
        }
    }
}

template <class Mdl, int G, int METHOD>
int launch_one(const DevLayer &L, const double *ps, long long ldb, long long B, const ShootOut &o, cudaStream_t s)
{
    const long long nthreads = B * L.I * (long long)L.R;
    if (nthreads <= 0) return 0;
    const long long nblocks = (nthreads + 127) / 128;
    shoot_kernel<Mdl, G, METHOD><<<(unsigned)nblocks, 128, 0, s>>>(L, ps, ldb, B, o);
    return (int)cudaGetLastError();
}

#define CB200_DISPATCH_G(Mdl, Gv)                                                           \
    case Gv:                                                                                \
        return method == 0 ? launch_one<Mdl, Gv, 0>(L, ps, ldb, B, o, s)                   \
                           : launch_one<Mdl, Gv, 1>(L, ps, ldb, B, o, s);

}  // namespace cb200
