// Fisher-information read-out shared by the stand-alone kernels (api.cu) and the shooting kernels' fused epilogue.
#pragma once
#include <cuda_runtime.h>

namespace cb200 {
constexpr int CB200_MAX_FIM = 8;
// common tail of the read-outs: A = the full symmetric matrix of this thread's candidate.
// With `crit` the six optimality criteria of lib/CorleoneOED/src/criteria.jl:33-63 are evaluated per candidate on
// Symmetric(F) (rows of crit: A = tr(inv F), D = 1/det F, E = max eig(inv F), FisherA = -tr F, FisherD = -det F,
// FisherE = -min eig F) from the eigenvalues of F (cyclic Jacobi, n <= CB200_MAX_FIM).
template <int NMAX>
__device__ __forceinline__ void fim_emit(double (&A)[NMAX][NMAX], int n, bool live, long long b,
                                         double *__restrict__ fim, long long ldo, double *__restrict__ fsum,
                                         double *__restrict__ crit, bool warp_sum = true)
{
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) {
            double v = live ? A[r][c] : 0.0;
            if (live && fim) fim[(long long)(r + n * c) * ldo + b] = v;
            if (fsum) {
                if (warp_sum) {   // whole warps call this: shuffle tree, one atomic per warp and entry
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                    if ((threadIdx.x & 31) == 0) atomicAdd(fsum + r + n * c, v);
                } else if (live) {
                    atomicAdd(fsum + r + n * c, v);
                }
            }
        }
    if (!crit || !live) return;
    double tr = 0.0;
    for (int k = 0; k < n; k++) tr += A[k][k];
    for (int sweep = 0; sweep < 40; sweep++) {   // cyclic Jacobi on the symmetric matrix; eigenvalues end on the diagonal
        double off = 0.0, dia = 0.0;
        for (int p = 0; p < n; p++) {
            dia += A[p][p] * A[p][p];
            for (int q = p + 1; q < n; q++) off += A[p][q] * A[p][q];
        }
        if (off <= 1e-32 * dia || off == 0.0) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < n; k++) {   // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {   // A <- J' A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
            }
    }
    double det = 1.0, trinv = 0.0, lmin = A[0][0];
    for (int k = 0; k < n; k++) {
        const double l = A[k][k];
        det *= l;
        trinv += 1.0 / l;
        lmin = fmin(lmin, l);
    }
    // max eig(inv F) = 1 / (smallest eigenvalue of F) for positive definite F; for an indefinite F the largest
    // reciprocal is taken, as eigvals(inv(F)) would give
    double einv = 1.0 / A[0][0];
    for (int k = 1; k < n; k++) einv = fmax(einv, 1.0 / A[k][k]);
    crit[0 * ldo + b] = trinv;
    crit[1 * ldo + b] = 1.0 / det;
    crit[2 * ldo + b] = einv;
    crit[3 * ldo + b] = -tr;
    crit[4 * ldo + b] = -det;
    crit[5 * ldo + b] = -lmin;
}

}  // namespace cb200
