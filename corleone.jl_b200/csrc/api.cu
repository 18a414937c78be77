// C ABI of libcorleone_b200.so (include/corleone_b200.h): handle, table upload, evaluation pipeline,
// epilogue kernels (defects / objective / gradient / Fisher information), layout conversion.
#include "../../include/corleone_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges around every C-ABI call (no cost without a profiler attached)

#include "handle.cuh"
#include "tableau.cuh"
#include "tail.cuh"
#include "fim.cuh"

namespace {
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
struct D20Slot {
    int refs = 0;
    double data[32 * 32 + 32];
};
std::mutex g_d20_mu;
D20Slot g_d20[4][64];   // per dense model (dense_slot) and device ordinal
}  // namespace

using namespace cb200;

// the dense family: slot of the model's __constant__ block, states; -1: not a dense model
static int dense_slot(int model, int *nx = nullptr)
{
    int s = -1, n = 0;
    switch (model) {
    case CB200_MODEL_DENSE20: s = 0; n = 20; break;
    case CB200_MODEL_DENSE16: s = 1; n = 16; break;
    case CB200_MODEL_DENSE24: s = 2; n = 24; break;
    case CB200_MODEL_DENSE32: s = 3; n = 32; break;
    default: break;
    }
    if (nx) *nx = n;
    return s;
}

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
int cb200_fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// ------------------------------------------------------------------ kernels
// objective = last(getsym(traj))  (src/dynprob.jl:70-71): one row of the last merged sample.  col[i*ldc + b] is
// the end value of that state on interval i (written by the shooting kernel); a quadrature row carries the
// running sum over intervals in interval order (src/multiple_shooting.jl:171-177).
__global__ void objective_kernel(const double *__restrict__ col, long long ldc, const double *__restrict__ ps,
                                 long long ldb, long long B, int I, int has_state, int is_quad, int ctrl_var,
                                 double *__restrict__ obj, double *__restrict__ obj_sum, const TailArgs tail)
{
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    double v = 0.0;
    if (b < B) {
        if (has_state) {
            if (is_quad) {
                v = col[b];  // q_prev = us[1][end][q]; then += across intervals in order
                for (int i = 1; i < I; i++) v = col[(long long)i * ldc + b] + v;
            } else {
                v = col[(long long)(I - 1) * ldc + b];
            }
        } else {
            v = ps[(long long)ctrl_var * ldb + b];
        }
        if (obj) obj[b] = v;
    }
    if (obj_sum) {  // whole warps reach this point: warp-shuffle tree, one atomic per warp
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(obj_sum, v);
    }
    tail_maybe(tail);   // the last block hands the cross-candidate sums on (all-reduced over the ranks) and closes the epoch
}

__global__ void grad_kernel(const GradEntry *__restrict__ tab, int n, const double *__restrict__ sens,
                            long long lds, long long B, double *__restrict__ grad, long long ldo,
                            double *__restrict__ grad_sum)
{
    // grid.x tiles the candidates, grid.y strides the entries: a warp never straddles two entries, so
    // the per-entry sum over candidates is a warp-shuffle tree + one atomicAdd per warp
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (int e = blockIdx.y; e < n; e += gridDim.y) {
        const GradEntry t = tab[e];
        double v = 0.0;
        if (b < B) {
            v = t.src >= 0 ? sens[t.src * lds + b] : 1.0;
            if (grad) grad[(long long)t.var * ldo + b] = v;
        }
        if (grad_sum) {
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(grad_sum + t.var, v);
        }
    }
}

// quadrature accumulation on the merged trajectory (src/multiple_shooting.jl:171-177)
__global__ void traj_quad_kernel(double *__restrict__ traj, long long ldo, const double *__restrict__ xend,
                                 long long ldx, long long B, int nx, int nrow, int I, const int *__restrict__ samp_off,
                                 const int *__restrict__ nsm, const int *__restrict__ quad, int nquad)
{
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int q = 0; q < nquad; q++) {
        const int k = quad[q];
        double run = xend[(long long)k * ldx + b];
        for (int i = 1; i < I; i++) {
            const int nsmp = nsm[i];
            for (int j = 0; j < nsmp; j++) traj[((long long)(samp_off[i] + j) * nrow + k) * ldo + b] += run;
            run += xend[((long long)i * nx + k) * ldx + b];
        }
    }
}

// K4 read-out: F = F_init + Symmetric(F_triu(t_end)) per candidate (lib/CorleoneOED/src/oed.jl:388-398,
// augmentation.jl:233) and its sum over the candidates of this launch: warp-shuffle + one atomicAdd per
// warp and matrix entry.
// With `crit` the six optimality criteria of lib/CorleoneOED/src/criteria.jl:33-63 are evaluated per candidate on
// Symmetric(F) (rows of crit: A = tr(inv F), D = 1/det F, E = max eig(inv F), FisherA = -tr F, FisherD = -det F,
// FisherE = -min eig F) from the eigenvalues of F (cyclic Jacobi, n <= CB200_MAX_FIM).
// fim_emit: fim.cuh
__global__ void fim_kernel(const double *__restrict__ xend, long long ldx, long long B, int nx, int I, int f_off,
                           int n, const double *__restrict__ finit, double *__restrict__ fim, long long ldo,
                           double *__restrict__ fsum, double *__restrict__ crit, const TailArgs tail)
{
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = b < B;
    double A[CB200_MAX_FIM][CB200_MAX_FIM];
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) {
            const int a = r < c ? r : c, bb = r < c ? c : r;  // upper-triangle entry (a, bb), a <= bb
            const int q = bb * (bb + 1) / 2 + a;             // column-major triu position
            A[r][c] = live ? xend[((long long)(I - 1) * nx + f_off + q) * ldx + b] + (finit ? finit[r + n * c] : 0.0) : 0.0;
        }
    fim_emit(A, n, live, b, fim, ldo, fsum, crit);
    tail_maybe(tail);
}

// K4 read-out of the sample-based variants (lib/CorleoneOED/src/oed.jl): the information is collected from the saved
// samples of the merged trajectory instead of one end state.  g_k(s)[a] = G[k + bx*a] at sample s (observed = the
// leading n_obs states).  rows: one per (interval, observation-grid row); row_sample = its sample, row_var[row*n_obs+k]
// = flat ps index of the weight of observed function k there, -1 = not measured (WeightedObservation, oed.jl:2-16).
//   mode 1  Discrete, no measurements (:446-461): F = sum over ALL samples of s s', s = sum_k g_k (augmentation.jl:182-186)
//   mode 2  DiscreteSampled (:409-432): F = sum_rows sum_k (w_k g_k)'(w_k g_k)   -- the weight enters squared (:8-10)
//   mode 3  ContinuousSampled, fixed (:352-362): F = sum_rows sum_k w_k (F_k(sample+1) - F_k(sample))
__global__ void fim_samples_kernel(const double *__restrict__ traj, long long ldt, long long B, int nrow, int n_samples,
                                   const double *__restrict__ ps, long long ldb, int mode, int bx, int n, int n_obs,
                                   int f_off, int n_rows, const int *__restrict__ row_sample,
                                   const int *__restrict__ row_var, const double *__restrict__ finit,
                                   double *__restrict__ fim, long long ldo, double *__restrict__ fsum,
                                   double *__restrict__ crit, const TailArgs tail)
{
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = b < B;
    double A[CB200_MAX_FIM][CB200_MAX_FIM];
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++) A[r][c] = (live && finit) ? finit[r + n * c] : 0.0;
    if (live) {
        const int ntri = n * (n + 1) / 2;
        const int count = mode == 1 ? n_samples : n_rows;
        for (int row = 0; row < count; row++) {
            const int smp = mode == 1 ? row : row_sample[row];
            const double *u = traj + (long long)nrow * smp * ldt + b;   // state r of the sample: u[r * ldt]
            if (mode == 1) {
                double sg[CB200_MAX_FIM];
                for (int a = 0; a < n; a++) {
                    double acc = 0.0;
                    for (int k = 0; k < n_obs; k++) acc += u[(long long)(bx + k + bx * a) * ldt];
                    sg[a] = acc;
                }
                for (int c = 0; c < n; c++)
                    for (int r = 0; r < n; r++) A[r][c] += sg[r] * sg[c];
                continue;
            }
            for (int k = 0; k < n_obs; k++) {
                const int var = row_var[row * n_obs + k];
                if (var < 0) continue;
                const double w = ps[(long long)var * ldb + b];
                if (mode == 2) {
                    double g[CB200_MAX_FIM];
                    for (int a = 0; a < n; a++) g[a] = w * u[(long long)(bx + k + bx * a) * ldt];
                    for (int c = 0; c < n; c++)
                        for (int r = 0; r < n; r++) A[r][c] += g[r] * g[c];
                } else {
                    const double *un = u + (long long)nrow * ldt;   // the next sample
                    for (int c = 0; c < n; c++)
                        for (int r = 0; r < n; r++) {
                            const int lo = r < c ? r : c, hi = r < c ? c : r;
                            const long long q = f_off + k * ntri + hi * (hi + 1) / 2 + lo;
                            A[r][c] += w * (un[q * ldt] - u[q * ldt]);
                        }
                }
            }
        }
    }
    fim_emit(A, n, live, b, fim, ldo, fsum, crit);
    tail_maybe(tail);
}


// tiled transpose: in[r*ld_in + c] (nr x nc) -> out[c*ld_out + r]
__global__ void transpose_kernel(const double *__restrict__ in, long long ld_in, double *__restrict__ out,
                                 long long ld_out, long long nr, long long nc)
{
    __shared__ double tile[32][33];
    const long long c0 = blockIdx.x * 32LL, r0 = blockIdx.y * 32LL;
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
        const long long r = r0 + rr, c = c0 + threadIdx.x;
        if (r < nr && c < nc) tile[rr][threadIdx.x] = in[r * ld_in + c];
    }
    __syncthreads();
    for (int cc = threadIdx.y; cc < 32; cc += 8) {
        const long long c = c0 + cc, r = r0 + threadIdx.x;
        if (r < nr && c < nc) out[c * ld_out + r] = tile[threadIdx.x][cc];
    }
}

// FP64 peak probe: 8 independent DFMA chains per thread
__global__ void dfma_probe_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c);
        a1 = fma(a1, m, c);
        a2 = fma(a2, m, c);
        a3 = fma(a3, m, c);
        a4 = fma(a4, m, c);
        a5 = fma(a5, m, c);
        a6 = fma(a6, m, c);
        a7 = fma(a7, m, c);
    }
    out[blockIdx.x * (long long)blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// FP64 tensor-path probe: 4 independent mma.sync.m8n8k4.f64 accumulator chains per warp.  On B200 DMMA and DFMA
// share one FP64 datapath (tools/dmma_micro.cu: the mixed rate never exceeds either alone); DMMA reaches the
// nominal 64 FMA/clk/SM, DFMA ~58, so the higher of the two probes is the FP64 roof.
__global__ void dmma_probe_kernel(double *out, int iters, double seed)
{
    double c[8];
    for (int i = 0; i < 8; i++) c[i] = seed + 1e-3 * threadIdx.x + i;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[2 * i]), "+d"(c[2 * i + 1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
    for (int i = 0; i < 8; i++) s += c[i];
    out[blockIdx.x * (long long)blockDim.x + threadIdx.x] = s;
}

static int default_G(int model)
{
    switch (model) {
    case CB200_MODEL_LQR: return 4;   // profiles/: G=4 at 4 blocks/SM beats G=8 at 2 (register-bound occupancy)
    case CB200_MODEL_LOTKA3: return 2;
    case CB200_MODEL_LOTKA2: return 2;
    case CB200_MODEL_LIN1D: return 4;
    case CB200_MODEL_EGERSTEDT: return 2;
    case CB200_MODEL_VDP3: return 2;
    case CB200_MODEL_CARTPOLE5: return 1;
    case CB200_MODEL_ROCKET3: return 2;
    case CB200_MODEL_QUADROTOR7: return 1;
    case CB200_MODEL_THREETANK4: return 2;
    case CB200_MODEL_BATCHREACTOR2: return 2;
    case CB200_MODEL_LOTKACOMP3: return 2;
    case CB200_MODEL_F8AIRCRAFT4: return 2;
    case CB200_MODEL_DOUBLEOSC6: return 1;
    case CB200_MODEL_DENSE20: case CB200_MODEL_DENSE16: case CB200_MODEL_DENSE24: case CB200_MODEL_DENSE32: return 1;
    case CB200_MODEL_LOTKA2_OED: return 1;   // second-order sensitivities of [x; G; F]: 12 dynamic + 6 quadrature per thread (G = 2 spills)
    case CB200_MODEL_LIN1D_OED: return 2;
    case CB200_MODEL_LOTKA2_OED_CU: return 1;
    case CB200_MODEL_LOTKA2_OED_CF: return 1;
    case CB200_MODEL_LOTKA2_OED_DS: return 1;
    case CB200_MODEL_LOTKA2_OED_DU: return 1;
    case CB200_MODEL_LIN1D_OED_CF: return 2;
    case CB200_MODEL_LIN1D_OED_DS: return 2;
    case CB200_MODEL_BRYSONDENHAM3: return 2;
    case CB200_MODEL_CATALYSTMIXING2: return 2;
    case CB200_MODEL_CUSHIONED3: return 2;
    case CB200_MODEL_DENBIGH3: return 2;
    case CB200_MODEL_DIELECTROPHORETIC3: return 2;
    case CB200_MODEL_ELECTRICCAR4: return 2;
    case CB200_MODEL_FULLER3: return 2;
    case CB200_MODEL_JACKSON3: return 2;
    case CB200_MODEL_MOUNTAINCAR3: return 2;
    case CB200_MODEL_PARTICLESTEERING5: return 1;
    case CB200_MODEL_RAOMEASE2: return 2;
    case CB200_MODEL_ROBBINS4: return 2;
    case CB200_MODEL_MOONLANDING3: return 2;
    case CB200_MODEL_LOTKASHARED4: return 2;
    case CB200_MODEL_HANGINGCHAIN3: return 2;
    case CB200_MODEL_TUBULARREACTOR2: return 2;
    case CB200_MODEL_OCEAN4: return 2;
    case CB200_MODEL_DUCTEDFAN7: return 1;
    case CB200_MODEL_HANGGLIDER4: return 2;
    default: return 0;
    }
}

static int model_dims(int model, int *nx, int *np, int *nxd = nullptr)
{
    int d = 0;   // dynamic states: the leading states that feed a right-hand side (the rest are quadrature-like)
    switch (model) {
    case CB200_MODEL_LOTKA3: *nx = 3; *np = 3; d = 2; break;
    case CB200_MODEL_LQR: *nx = 2; *np = 3; d = 1; break;
    case CB200_MODEL_LOTKA2: *nx = 2; *np = 3; d = 2; break;
    case CB200_MODEL_LOTKA2_OED: *nx = 9; *np = 5; d = 6; break;
    case CB200_MODEL_LIN1D: *nx = 1; *np = 1; d = 1; break;
    case CB200_MODEL_LIN1D_OED: *nx = 3; *np = 2; d = 2; break;
    case CB200_MODEL_DENSE20: *nx = 20; *np = 1; d = 20; break;
    case CB200_MODEL_DENSE16: *nx = 16; *np = 1; d = 16; break;
    case CB200_MODEL_DENSE24: *nx = 24; *np = 1; d = 24; break;
    case CB200_MODEL_DENSE32: *nx = 32; *np = 1; d = 32; break;
    case CB200_MODEL_EGERSTEDT: *nx = 3; *np = 3; d = 2; break;
    case CB200_MODEL_VDP3: *nx = 3; *np = 1; d = 2; break;
    case CB200_MODEL_CARTPOLE5: *nx = 5; *np = 1; d = 4; break;
    case CB200_MODEL_ROCKET3: *nx = 3; *np = 2; d = 3; break;
    case CB200_MODEL_QUADROTOR7: *nx = 7; *np = 4; d = 6; break;
    case CB200_MODEL_THREETANK4: *nx = 4; *np = 3; d = 3; break;
    case CB200_MODEL_BATCHREACTOR2: *nx = 2; *np = 1; d = 2; break;
    case CB200_MODEL_LOTKACOMP3: *nx = 3; *np = 1; d = 2; break;
    case CB200_MODEL_F8AIRCRAFT4: *nx = 4; *np = 2; d = 3; break;
    case CB200_MODEL_DOUBLEOSC6: *nx = 6; *np = 1; d = 5; break;
    case CB200_MODEL_LOTKA2_OED_CU: *nx = 9; *np = 3; d = 6; break;
    case CB200_MODEL_LOTKA2_OED_CF: *nx = 12; *np = 5; d = 6; break;
    case CB200_MODEL_LOTKA2_OED_DS: *nx = 6; *np = 3; d = 6; break;
    case CB200_MODEL_LOTKA2_OED_DU: *nx = 6; *np = 3; d = 6; break;
    case CB200_MODEL_LIN1D_OED_CF: *nx = 3; *np = 2; d = 2; break;
    case CB200_MODEL_LIN1D_OED_DS: *nx = 2; *np = 1; d = 2; break;
    case CB200_MODEL_BRYSONDENHAM3: *nx = 3; *np = 1; d = 2; break;
    case CB200_MODEL_CATALYSTMIXING2: *nx = 2; *np = 1; d = 2; break;
    case CB200_MODEL_CUSHIONED3: *nx = 3; *np = 2; d = 2; break;
    case CB200_MODEL_DENBIGH3: *nx = 3; *np = 1; d = 2; break;
    case CB200_MODEL_DIELECTROPHORETIC3: *nx = 3; *np = 2; d = 2; break;
    case CB200_MODEL_ELECTRICCAR4: *nx = 4; *np = 1; d = 2; break;
    case CB200_MODEL_FULLER3: *nx = 3; *np = 1; d = 2; break;
    case CB200_MODEL_JACKSON3: *nx = 3; *np = 1; d = 2; break;
    case CB200_MODEL_MOUNTAINCAR3: *nx = 3; *np = 2; d = 2; break;
    case CB200_MODEL_PARTICLESTEERING5: *nx = 5; *np = 2; d = 4; break;
    case CB200_MODEL_RAOMEASE2: *nx = 2; *np = 1; d = 1; break;
    case CB200_MODEL_ROBBINS4: *nx = 4; *np = 1; d = 3; break;
    case CB200_MODEL_MOONLANDING3: *nx = 3; *np = 2; d = 3; break;
    case CB200_MODEL_LOTKASHARED4: *nx = 4; *np = 1; d = 3; break;
    case CB200_MODEL_HANGINGCHAIN3: *nx = 3; *np = 1; d = 1; break;
    case CB200_MODEL_TUBULARREACTOR2: *nx = 2; *np = 1; d = 1; break;
    case CB200_MODEL_OCEAN4: *nx = 4; *np = 2; d = 3; break;
    case CB200_MODEL_DUCTEDFAN7: *nx = 7; *np = 3; d = 6; break;
    case CB200_MODEL_HANGGLIDER4: *nx = 4; *np = 2; d = 4; break;
    default: return -1;
    }
    if (nxd) *nxd = d;
    return 0;
}

static int launch_model(int model, int G, int method, const DevLayer &L, const double *ps, long long ldb,
                        long long B, const ShootOut &o, cudaStream_t s)
{
    switch (model) {
    case CB200_MODEL_LOTKA3: return launch_lotka3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LQR: return launch_lqr(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKA2: return launch_lotka2(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKA2_OED: return launch_lotka2_oed(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LIN1D: return launch_lin1d(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LIN1D_OED: return launch_lin1d_oed(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DENSE20: return launch_dense20(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DENSE16: return launch_dense16(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DENSE24: return launch_dense24(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DENSE32: return launch_dense32(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_EGERSTEDT: return launch_egerstedt(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_VDP3: return launch_vdp3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_CARTPOLE5: return launch_cartpole5(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_ROCKET3: return launch_rocket3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_QUADROTOR7: return launch_quadrotor7(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_THREETANK4: return launch_threetank4(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_BATCHREACTOR2: return launch_batchreactor2(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKACOMP3: return launch_lotkacomp3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_F8AIRCRAFT4: return launch_f8aircraft4(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DOUBLEOSC6: return launch_doubleosc6(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKA2_OED_CU: return launch_lotka2_oed_cu(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKA2_OED_CF: return launch_lotka2_oed_cf(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKA2_OED_DS: return launch_lotka2_oed_ds(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKA2_OED_DU: return launch_lotka2_oed_du(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LIN1D_OED_CF: return launch_lin1d_oed_cf(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LIN1D_OED_DS: return launch_lin1d_oed_ds(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_BRYSONDENHAM3: return launch_brysondenham3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_CATALYSTMIXING2: return launch_catalystmixing2(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_CUSHIONED3: return launch_cushioned3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DENBIGH3: return launch_denbigh3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DIELECTROPHORETIC3: return launch_dielectrophoretic3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_ELECTRICCAR4: return launch_electriccar4(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_FULLER3: return launch_fuller3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_JACKSON3: return launch_jackson3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_MOUNTAINCAR3: return launch_mountaincar3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_PARTICLESTEERING5: return launch_particlesteering5(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_RAOMEASE2: return launch_raomease2(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_ROBBINS4: return launch_robbins4(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_MOONLANDING3: return launch_moonlanding3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_LOTKASHARED4: return launch_lotkashared4(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_HANGINGCHAIN3: return launch_hangingchain3(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_TUBULARREACTOR2: return launch_tubularreactor2(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_OCEAN4: return launch_ocean4(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_DUCTEDFAN7: return launch_ductedfan7(G, method, L, ps, ldb, B, o, s);
    case CB200_MODEL_HANGGLIDER4: return launch_hangglider4(G, method, L, ps, ldb, B, o, s);
    default: return -1000;
    }
}

static int forward_model(int model, int method, const DevLayer &L, double *ps, long long ldb, long long B,
                         const unsigned char *fixed, cudaStream_t s)
{
    switch (model) {
    case CB200_MODEL_LOTKA3: return forward_lotka3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LQR: return forward_lqr(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKA2: return forward_lotka2(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKA2_OED: return forward_lotka2_oed(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LIN1D: return forward_lin1d(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LIN1D_OED: return forward_lin1d_oed(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DENSE20: return forward_dense20(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DENSE16: return forward_dense16(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DENSE24: return forward_dense24(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DENSE32: return forward_dense32(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_EGERSTEDT: return forward_egerstedt(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_VDP3: return forward_vdp3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_CARTPOLE5: return forward_cartpole5(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_ROCKET3: return forward_rocket3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_QUADROTOR7: return forward_quadrotor7(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_THREETANK4: return forward_threetank4(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_BATCHREACTOR2: return forward_batchreactor2(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKACOMP3: return forward_lotkacomp3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_F8AIRCRAFT4: return forward_f8aircraft4(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DOUBLEOSC6: return forward_doubleosc6(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKA2_OED_CU: return forward_lotka2_oed_cu(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKA2_OED_CF: return forward_lotka2_oed_cf(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKA2_OED_DS: return forward_lotka2_oed_ds(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKA2_OED_DU: return forward_lotka2_oed_du(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LIN1D_OED_CF: return forward_lin1d_oed_cf(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LIN1D_OED_DS: return forward_lin1d_oed_ds(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_BRYSONDENHAM3: return forward_brysondenham3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_CATALYSTMIXING2: return forward_catalystmixing2(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_CUSHIONED3: return forward_cushioned3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DENBIGH3: return forward_denbigh3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DIELECTROPHORETIC3: return forward_dielectrophoretic3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_ELECTRICCAR4: return forward_electriccar4(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_FULLER3: return forward_fuller3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_JACKSON3: return forward_jackson3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_MOUNTAINCAR3: return forward_mountaincar3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_PARTICLESTEERING5: return forward_particlesteering5(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_RAOMEASE2: return forward_raomease2(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_ROBBINS4: return forward_robbins4(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_MOONLANDING3: return forward_moonlanding3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_LOTKASHARED4: return forward_lotkashared4(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_HANGINGCHAIN3: return forward_hangingchain3(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_TUBULARREACTOR2: return forward_tubularreactor2(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_OCEAN4: return forward_ocean4(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_DUCTEDFAN7: return forward_ductedfan7(method, L, ps, ldb, B, fixed, s);
    case CB200_MODEL_HANGGLIDER4: return forward_hangglider4(method, L, ps, ldb, B, fixed, s);
    default: return -1000;
    }
}

extern "C" int cb200_version(void) { return CB200_VERSION; }
extern "C" const char *cb200_last_error(void) { return g_err.c_str(); }

template <class T>
static size_t push(std::vector<char> &blob, const std::vector<T> &v)
{
    size_t off = (blob.size() + 15) & ~size_t(15);
    blob.resize(off + sizeof(T) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) memcpy(blob.data() + off, v.data(), sizeof(T) * v.size());
    return off;
}

extern "C" int cb200_create(const cb200_desc *d, cb200_handle **out)
{
    NvtxRange nvtx_range("cb200_create");
    if (!d || !out) return fail(CB200_ERR_ARG, "null argument");
    *out = nullptr;
    if (d->struct_size != (int32_t)sizeof(cb200_desc))
        return fail(CB200_ERR_ARG, "cb200_desc.struct_size %d != %d", d->struct_size, (int)sizeof(cb200_desc));
    int mnx = 0, mnp = 0, mnxd = 0;
    if (model_dims(d->model, &mnx, &mnp, &mnxd)) return fail(CB200_ERR_UNSUPPORTED, "unknown model id %d", d->model);
    if (d->nx != mnx || d->np_all != mnp)
        return fail(CB200_ERR_ARG, "model %d expects nx=%d, np=%d; got nx=%d, np=%d", d->model, mnx, mnp, d->nx,
                    d->np_all);
    if (d->method != CB200_TSIT5 && d->method != CB200_RK4 && d->method != CB200_TSIT5_ADAPTIVE)
        return fail(CB200_ERR_UNSUPPORTED, "unknown method");
    if (d->steps_per_piece < 1) return fail(CB200_ERR_ARG, "steps_per_piece must be >= 1");
    if (d->method == CB200_TSIT5_ADAPTIVE) {
        if (!(d->abstol > 0.0) || !(d->reltol > 0.0)) return fail(CB200_ERR_ARG, "adaptive Tsit5 needs abstol > 0 and reltol > 0");
    }
    if (d->n_intervals < 1) return fail(CB200_ERR_ARG, "n_intervals must be >= 1");
    if (!d->u0 || !d->n_pieces || !d->tspans || !d->n_u0 || !d->n_local_controls || !d->is_shooting || !d->ic_sorting ||
        !d->ic_n_constants || !d->n_shooting_indices || (d->n_controls > 0 && (!d->index_grid || !d->control_indices)) ||
        (d->n_tunable_p > 0 && !d->tunable_p) || (d->n_quadrature > 0 && !d->quadrature_indices))
        return fail(CB200_ERR_ARG, "a required table pointer of cb200_desc is NULL");
    {   // counts a C-ABI caller can get wrong: everything that is dereferenced or used as an index below
        long long n_si = 0, n_ic = 0;
        for (int i = 0; i < d->n_intervals; i++) {
            if (d->n_u0[i] < 0 || d->n_u0[i] > d->nx) return fail(CB200_ERR_ARG, "interval %d: n_u0 outside 0..nx", i + 1);
            if (d->n_local_controls[i] < 0) return fail(CB200_ERR_ARG, "interval %d: negative n_local_controls", i + 1);
            if (d->n_shooting_indices[i] < 0 || d->ic_n_constants[i] < 0)
                return fail(CB200_ERR_ARG, "interval %d: negative table length", i + 1);
            n_si += d->n_shooting_indices[i];
            n_ic += d->ic_n_constants[i];
        }
        if ((n_si > 0 && !d->shooting_indices) || (n_ic > 0 && !d->ic_constants))
            return fail(CB200_ERR_ARG, "shooting_indices / ic_constants is NULL although its per-interval counts are not zero");
        for (long long q = 0; q < n_si; q++)
            if (d->shooting_indices[q] < 1 || d->shooting_indices[q] > d->nx + d->n_controls)
                return fail(CB200_ERR_ARG, "shooting_indices[%lld] = %lld outside 1..nx + n_controls", q, (long long)d->shooting_indices[q]);
        for (long long q = 0; q < n_ic; q++)
            if (d->ic_constants[q] < 1 || d->ic_constants[q] > d->nx)
                return fail(CB200_ERR_ARG, "ic_constants[%lld] outside 1..nx", q);
        if (d->fim_offset > 0 && (d->fim_n < 1 || d->fim_n > 8))   // 8 = CB200_MAX_FIM of the read-out kernels
            return fail(CB200_ERR_ARG, "fim_n = %d with a Fisher block (fim_offset > 0): need 1 <= fim_n <= 8", d->fim_n);
    }
    if (d->n_controls < 0 || d->n_controls > MAX_CTRL) return fail(CB200_ERR_ARG, "n_controls out of range");
    if (d->np_all > MAX_NP) return fail(CB200_ERR_ARG, "np_all too large");
    if (d->n_tunable_p + d->n_controls != d->np_all)
        return fail(CB200_ERR_ARG, "n_tunable_p + n_controls must equal np_all (every p entry is a parameter or a control)");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(CB200_ERR_CUDA, "no CUDA device (%s); libcorleone_b200 has no CPU fallback",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    if (d->device < 0 || d->device >= ndev) return fail(CB200_ERR_ARG, "device ordinal %d out of range", d->device);
    CUDA_TRY(cudaSetDevice(d->device));

    cb200_handle *h = new (std::nothrow) cb200_handle();
    if (!h) return fail(CB200_ERR_NOMEM, "out of host memory");
    h->device = d->device;
    h->model = d->model;
    h->method = d->method;
    DevLayer &L = h->L;
    const int nx = d->nx, I = d->n_intervals, nact = d->n_controls, ntp = d->n_tunable_p;
    L.nx = nx;
    L.np_all = d->np_all;
    L.nact = nact;
    L.ntp = ntp;
    L.I = I;
    L.i0 = 0;
    L.In = I;
    L.K = d->steps_per_piece;
    L.abstol = d->abstol;
    L.reltol = d->reltol;
    L.R = 1;
    h->G_sens = default_G(d->model);
    if (const char *env = getenv("CB200_DIRS_PER_THREAD")) {
        int g = atoi(env);
        if (g > 0 && g <= 24) h->G_sens = g;
    }

    // scatter maps A, B (src/single_shooting.jl:306-318): column ids in ascending p index
    {
        int pid = 0, cid = 0;
        for (int q = 1; q <= d->np_all; q++) {
            bool is_c = false;
            for (int r = 0; r < nact; r++) is_c |= (d->control_indices[r] == q);
            if (is_c) {
                L.pmap_kind[q - 1] = 1;
                L.pmap_idx[q - 1] = cid;
                L.row_to_p[cid] = q - 1;
                cid++;
            } else {
                if (pid >= ntp || d->tunable_p[pid] != q) {
                    delete h;
                    return fail(CB200_ERR_ARG, "tunable_p is not setdiff(eachindex(p), control_indices)");
                }
                L.pmap_kind[q - 1] = 0;
                L.pmap_idx[q - 1] = pid;
                L.tunable_p0[pid] = q - 1;
                pid++;
            }
        }
        if (cid != nact) {
            delete h;
            return fail(CB200_ERR_ARG, "control_indices must be distinct and within 1..np_all");
        }
    }
    h->nquad = d->n_quadrature;
    for (int k = 0; k < d->n_quadrature; k++) h->quad0.push_back((int)d->quadrature_indices[k] - 1);
    h->fim_mode = d->fim_mode;
    if (d->fim_mode < 0 || d->fim_mode > CB200_FIM_FIXED_CONTINUOUS) {
        delete h;
        return fail(CB200_ERR_ARG, "unknown fim_mode %d", d->fim_mode);
    }
    if (d->fim_mode == CB200_FIM_END || d->fim_mode == CB200_FIM_FIXED_CONTINUOUS) {
        if (d->fim_offset > 0) {
            h->fim_off = d->fim_offset - 1;
            h->fim_n = d->fim_n;
            const int blocks = d->fim_mode == CB200_FIM_FIXED_CONTINUOUS ? d->n_observed : 1;
            if (h->fim_off + blocks * h->fim_n * (h->fim_n + 1) / 2 > nx) {
                delete h;
                return fail(CB200_ERR_ARG, "Fisher block exceeds the state vector");
            }
        } else if (d->fim_mode == CB200_FIM_FIXED_CONTINUOUS) {
            delete h;
            return fail(CB200_ERR_ARG, "fim_mode FIXED_CONTINUOUS needs fim_offset");
        }
    }
    if (d->fim_mode != CB200_FIM_END) {
        h->fim_n = d->fim_n;
        h->fim_bx = d->fim_base_nx;
        h->n_obs = d->n_observed;
        if (h->fim_off < 0) h->fim_off = 0;   // "has a Fisher read-out"
        if (h->fim_n < 1 || h->fim_n > CB200_MAX_FIM || h->fim_bx < 1 || h->n_obs < 1 || h->n_obs > h->fim_bx ||
            h->fim_bx * (1 + h->fim_n) > nx) {
            delete h;
            return fail(CB200_ERR_ARG, "sample-based Fisher read-out: need 1 <= fim_n <= %d, 1 <= n_observed <= fim_base_nx and [x; G] within the state", CB200_MAX_FIM);
        }
        if (d->fim_mode != CB200_FIM_DISCRETE && (!d->n_observation_rows || !d->observation_grid)) {
            delete h;
            return fail(CB200_ERR_ARG, "fim_mode %d needs observation_grid", d->fim_mode);
        }
        if (d->fim_mode == CB200_FIM_FIXED_CONTINUOUS && I != 1) {
            delete h;
            return fail(CB200_ERR_UNSUPPORTED, "the fixed continuous read-out is defined for single shooting only (oed.jl:352-362)");
        }
    }

    // saveat (problem.kwargs): sorted times; a piece saves those strictly inside it from the dense output
    if (d->n_saveat < 0 || (d->n_saveat > 0 && !d->saveat)) {
        delete h;
        return fail(CB200_ERR_ARG, "n_saveat = %d without saveat times", d->n_saveat);
    }
    for (int k = 1; k < d->n_saveat; k++)
        if (!(d->saveat[k] > d->saveat[k - 1])) {
            delete h;
            return fail(CB200_ERR_ARG, "saveat must be strictly ascending");
        }
    std::vector<int> psamp, nsm, sv_n, sv_off;
    std::vector<double> sv_t;
    // per-interval tables
    std::vector<double> pt0, pt1, ph, ic_fix, ic_fix0;
    int po = 0, vo = 0, so = 0, co = 0, sio = 0;
    long long seo = 0;
    long long igo = 0;
    std::vector<int> sidx_off(I + 1, 0);
    for (int i = 0; i < I; i++) {
        const int M = d->n_pieces[i];
        if (M < 1) {
            delete h;
            return fail(CB200_ERR_ARG, "interval %d has no pieces", i + 1);
        }
        h->h_M.push_back(M);
        h->h_piece_off.push_back(po);
        h->h_var_off.push_back(vo);
        h->h_n_u0.push_back(d->n_u0[i]);
        h->h_n_c.push_back(d->n_local_controls[i]);
        h->h_samp_off.push_back(so);
        h->h_sens_off.push_back(seo);
        h->h_tsens_off.push_back(h->n_tsens);
        const int nsi = d->n_u0[i] + ntp + d->n_local_controls[i];
        int samp_i = 0;   // samples of this interval after its start sample
        h->ns.push_back(nsi);
        h->max_ns = std::max(h->max_ns, nsi);
        for (int j = 0; j < M; j++) {
            pt0.push_back(d->tspans[2 * (size_t)(po + j)]);
            pt1.push_back(d->tspans[2 * (size_t)(po + j) + 1]);
            ph.push_back((pt1.back() - pt0.back()) / (double)d->steps_per_piece);
            sv_off.push_back((int)sv_t.size());
            int nin = 0;
            for (int k = 0; k < d->n_saveat; k++)
                if (d->saveat[k] > pt0.back() && d->saveat[k] < pt1.back()) {
                    sv_t.push_back(d->saveat[k]);
                    nin++;
                }
            sv_n.push_back(nin);
            if (nin > 0) h->has_saveat = true;
            if (j == 0) h->h_sample_t.push_back(pt0.back());   // every interval keeps its start sample, the end sample of all but the last is dropped
            for (int k = 0; k < nin; k++) h->h_sample_t.push_back(sv_t[sv_t.size() - (size_t)nin + (size_t)k]);
            if (j + 1 < M || i == I - 1) h->h_sample_t.push_back(pt1.back());
            samp_i += nin + 1;
            psamp.push_back(samp_i);
            for (int r = 0; r < nact; r++) {
                const long long ci = d->index_grid[igo + r + (long long)nact * j] - 1;
                if (ci < 0 || ci >= d->n_local_controls[i]) {
                    delete h;
                    return fail(CB200_ERR_ARG, "index_grid entry out of range in interval %d", i + 1);
                }
                h->h_cidx.push_back((int)ci);
            }
        }
        // InitialConditionRemaker (src/single_shooting.jl:260-266)
        const int ncon = d->ic_n_constants[i], nu0 = d->n_u0[i];
        for (int k = 0; k < nx; k++) {
            int src = -1;
            double fix = 0.0;
            if (nu0 == 0) {
                fix = d->is_shooting[i] ? 0.0 : d->u0[k];
            } else {
                if (ncon + nu0 != nx) {
                    delete h;
                    return fail(CB200_ERR_ARG, "interval %d: constants + tunables != nx", i + 1);
                }
                const int pos = (int)d->ic_sorting[(size_t)i * nx + k] - 1;
                if (pos < 0 || pos >= nx) {
                    delete h;
                    return fail(CB200_ERR_ARG, "ic_sorting out of range");
                }
                if (pos < ncon)
                    fix = d->is_shooting[i] ? 0.0 : d->u0[d->ic_constants[co + pos] - 1];
                else
                    src = pos - ncon;
            }
            h->h_ic_src.push_back(src);
            ic_fix.push_back(fix);
            // the forward sweep calls the layer without the shooting flag: fixval = problem.u0 (node_initialization.jl:165)
            ic_fix0.push_back(src >= 0 ? 0.0 : (nu0 == 0 ? d->u0[k] : d->u0[d->ic_constants[co + ((int)d->ic_sorting[(size_t)i * nx + k] - 1)] - 1]));
        }
        co += ncon;
        sidx_off[i + 1] = sidx_off[i] + d->n_shooting_indices[i];
        sio += d->n_shooting_indices[i];
        nsm.push_back(samp_i + (i == I - 1 ? 1 : 0));
        h->n_tsens += (long long)nsm.back() * nsi * nx;
        po += M;
        igo += (long long)nact * M;
        vo += nsi;
        so += samp_i;
        seo += (long long)nx * nsi;
    }
    if (h->has_saveat && d->method == CB200_RK4) {
        delete h;
        return fail(CB200_ERR_UNSUPPORTED, "saveat times inside a control piece need Tsit5's dense output (method 0 or 2)");
    }
    if (h->has_saveat && dense_slot(d->model) >= 0) {
        delete h;
        return fail(CB200_ERR_UNSUPPORTED, "saveat is not available for the dense tensor-path model");
    }
    (void)sio;
    L.nvars = vo;
    L.max_M = *std::max_element(h->h_M.begin(), h->h_M.end());
    h->n_sens = seo;
    h->n_samples = so + 1;

    // direction tables of the sensitivity kernel: the p column every local variable perturbs and, per piece
    // and direction group, the mask of directions whose forcing term f_p / f_u is switched on
    std::vector<signed char> dir_pcol((size_t)std::max(1, L.nvars), (signed char)-1);
    std::vector<int> dir_perm((size_t)std::max(1, L.nvars), -1);
    const int Gs = h->G_sens;
    const int Rmask = Gs > 0 ? std::max(1, (h->max_ns + Gs - 1) / Gs) : 1;
    std::vector<unsigned> act_mask((size_t)po * Rmask, 0u);
    // the same table for one direction per thread: latency-bound evaluations (a few candidates, flat launch) spread the
    // directions over more threads when the model has a G = 1 kernel
    const int Rmask1 = std::max(1, h->max_ns);
    std::vector<unsigned> act_mask1(Gs > 1 ? (size_t)po * Rmask1 : 0, 0u);
    for (int i = 0; i < I; i++) {
        const int nu0 = h->h_n_u0[i], nc = h->h_n_c[i], M = h->h_M[i], pof = h->h_piece_off[i], vof = h->h_var_off[i];
        const int nsi = h->ns[i];
        std::vector<int> crow((size_t)std::max(1, nc), -1);
        for (int j = 0; j < M; j++)
            for (int r = 0; r < nact; r++) crow[h->h_cidx[(size_t)(pof + j) * nact + r]] = r;
        // first piece in which a direction is non-zero: u0 seeds on dynamic states and parameters from the start,
        // a control value from the first piece that uses it, never (= M) for seeds on quadrature-like states (their
        // sensitivity is the constant seed) and for local controls no piece refers to
        std::vector<int> first((size_t)std::max(1, nsi), M);
        for (int dd = 0; dd < nsi; dd++) {
            int pc = -1;
            if (dd < nu0) {
                for (int k = 0; k < nx; k++)
                    if (h->h_ic_src[(size_t)i * nx + k] == dd && k < mnxd) first[dd] = 0;
            } else if (dd < nu0 + ntp) {
                pc = L.tunable_p0[dd - nu0];
                first[dd] = 0;
            } else if (crow[dd - nu0 - ntp] >= 0) {
                pc = L.row_to_p[crow[dd - nu0 - ntp]];
                for (int j = M - 1; j >= 0; j--)
                    if (h->h_cidx[(size_t)(pof + j) * nact + crow[dd - nu0 - ntp]] == dd - nu0 - ntp) first[dd] = j;
            }
            dir_pcol[(size_t)vof + dd] = (signed char)pc;
        }
        std::vector<int> order((size_t)nsi);
        for (int dd = 0; dd < nsi; dd++) order[dd] = dd;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return first[a] < first[b]; });
        for (int sl = 0; sl < nsi; sl++) {
            const int dd = order[sl];
            dir_perm[(size_t)vof + sl] = dd;
            if (Gs <= 0) continue;
            const int pc = dir_pcol[(size_t)vof + dd];
            for (int j = 0; j < M; j++) {
                unsigned &w = act_mask[(size_t)(pof + j) * Rmask + sl / Gs];
                const bool forced = pc >= 0 && ((dd < nu0 + ntp) ||
                                    h->h_cidx[(size_t)(pof + j) * nact + crow[dd - nu0 - ntp]] == dd - nu0 - ntp);
                if (forced) w |= 1u << (sl % Gs);
                if (j >= first[dd]) {   // bits 24..31: live prefix length of the group during the piece
                    const unsigned na = (unsigned)(sl % Gs) + 1u;
                    if ((w >> 24) < na) w = (w & 0xffffffu) | (na << 24);
                }
                if (!act_mask1.empty())
                    act_mask1[(size_t)(pof + j) * Rmask1 + sl] = (forced ? 1u : 0u) | (j >= first[dd] ? 1u << 24 : 0u);
            }
        }
    }

    // defect table + Jacobian structure (src/multiple_shooting.jl:150-163, 281-295)
    {
        int n_u = 0, n_p = 0, n_c = 0;
        for (int i = 1; i < I; i++) {
            for (int q = sidx_off[i]; q < sidx_off[i + 1]; q++) (d->shooting_indices[q] <= nx ? n_u : n_c)++;
            n_p += ntp;
        }
        h->jpos.assign((size_t)I * nx, -1);
        h->dpos_kind.assign((size_t)I * nx, -1);
        h->dpos_stage.assign((size_t)I * nx, -1);
        h->oth_off.assign((size_t)I + 1, 0);
        int ou = 0, op = n_u, oc = n_u + n_p, os = 0;
        auto ctrl_var = [&](int iv, int piece, int r) {
            return h->h_var_off[iv] + h->h_n_u0[iv] + ntp + h->h_cidx[(size_t)(h->h_piece_off[iv] + piece) * nact + r];
        };
        for (int i = 1; i < I; i++) {
            for (int pass = 0; pass < 3; pass++) {
                if (pass == 1) {
                    for (int k = 0; k < ntp; k++) {
                        DefectEntry e{1, h->h_var_off[i - 1] + h->h_n_u0[i - 1] + k, 1,
                                      h->h_var_off[i] + h->h_n_u0[i] + k, 0.0, op, os};
                        h->defects.push_back(e);
                        h->oth.push_back(MatchEntry{e.a_idx, e.b_idx, op, os});
                        h->jac_rows.push_back(op + 1); h->jac_cols.push_back(e.a_idx + 1); h->jac_src.push_back(-1);
                        h->jac_rows.push_back(op + 1); h->jac_cols.push_back(e.b_idx + 1); h->jac_src.push_back(-2);
                        op++;
                        os++;
                    }
                    continue;
                }
                for (int q = sidx_off[i]; q < sidx_off[i + 1]; q++) {
                    const int rI = (int)d->shooting_indices[q] - 1;
                    if ((pass == 0) != (rI < nx)) continue;
                    DefectEntry e{};
                    int &o = (pass == 0) ? ou : oc;
                    e.pos_kind = o;
                    e.pos_stage = os;
                    if (rI < nx) {
                        e.a_kind = 0;
                        e.a_idx = (i - 1) * nx + rI;
                        h->dpos_kind[(size_t)(i - 1) * nx + rI] = o;
                        h->dpos_stage[(size_t)(i - 1) * nx + rI] = os;
                        const int src = h->h_ic_src[(size_t)i * nx + rI];
                        if (src >= 0) {
                            e.b_kind = 1;
                            e.b_idx = h->h_var_off[i] + src;
                        } else {
                            e.b_kind = 2;
                            e.b_const = ic_fix[(size_t)i * nx + rI];
                        }
                        h->jpos[(size_t)(i - 1) * nx + rI] = (int)h->jac_rows.size();
                        for (int dd = 0; dd < h->ns[i - 1]; dd++) {
                            h->jac_rows.push_back(o + 1);
                            h->jac_cols.push_back(h->h_var_off[i - 1] + dd + 1);
                            h->jac_src.push_back(h->h_sens_off[i - 1] + (long long)dd * nx + rI);
                        }
                        if (e.b_kind == 1) {
                            h->jac_rows.push_back(o + 1); h->jac_cols.push_back(e.b_idx + 1); h->jac_src.push_back(-2);
                        }
                    } else {
                        const int r = rI - nx;
                        if (r >= nact) {
                            delete h;
                            return fail(CB200_ERR_ARG, "shooting index beyond [states; controls]");
                        }
                        e.a_kind = 1;
                        e.a_idx = ctrl_var(i - 1, h->h_M[i - 1] - 1, r);
                        e.b_kind = 1;
                        e.b_idx = ctrl_var(i, 0, r);
                        h->oth.push_back(MatchEntry{e.a_idx, e.b_idx, o, os});
                        h->jac_rows.push_back(o + 1); h->jac_cols.push_back(e.a_idx + 1); h->jac_src.push_back(-1);
                        h->jac_rows.push_back(o + 1); h->jac_cols.push_back(e.b_idx + 1); h->jac_src.push_back(-2);
                    }
                    h->defects.push_back(e);
                    o++;
                    os++;
                }
            }
            h->oth_off[i] = (int)h->oth.size();   // boundary i-1 owns oth[oth_off[i-1], oth_off[i])
        }
        for (int i = std::max(I, 1); i <= I; i++) h->oth_off[i] = (int)h->oth.size();
    }

    // triplet order: the sensitivity-valued entries first (their values are what cb200_eval delivers as jac_vals,
    // the kernel writes them straight to their slot), then the constant +-1 entries of the matching conditions
    {
        const size_t nnz = h->jac_src.size();
        std::vector<long long> newpos(nnz), r2, c2, s2;
        size_t nv = 0;
        for (size_t k = 0; k < nnz; k++)
            if (h->jac_src[k] >= 0) newpos[k] = (long long)nv++;
        h->jac_nnz_var = (long long)nv;
        for (size_t k = 0; k < nnz; k++)
            if (h->jac_src[k] < 0) newpos[k] = (long long)nv++;
        r2.resize(nnz); c2.resize(nnz); s2.resize(nnz);
        for (size_t k = 0; k < nnz; k++) {
            r2[newpos[k]] = h->jac_rows[k];
            c2[newpos[k]] = h->jac_cols[k];
            s2[newpos[k]] = h->jac_src[k];
        }
        for (int &jp : h->jpos)
            if (jp >= 0) jp = (int)newpos[jp];   // a state's ns entries stay contiguous: only constants moved out
        h->jac_rows.swap(r2); h->jac_cols.swap(c2); h->jac_src.swap(s2);
    }
    // observation rows of the sample-based Fisher read-outs
    std::vector<int> obs_sample, obs_var;
    if (h->fim_mode == CB200_FIM_DISCRETE_SAMPLED || h->fim_mode == CB200_FIM_FIXED_CONTINUOUS) {
        long long go = 0;
        for (int i = 0; i < I; i++) {
            const int rows = d->n_observation_rows[i];
            // rows <-> samples of the interval in the merged trajectory (DISCRETE_SAMPLED) / its pieces (FIXED_CONTINUOUS)
            const int max_rows = h->fim_mode == CB200_FIM_DISCRETE_SAMPLED ? nsm[(size_t)i] : h->h_M[i];
            if (rows < 0 || rows > max_rows) {
                delete h;
                return fail(CB200_ERR_ARG, "interval %d: %d observation rows, at most %d samples", i + 1, rows, max_rows);
            }
            for (int j = 0; j < rows; j++) {
                obs_sample.push_back(h->h_samp_off[i] + j);
                for (int k = 0; k < h->n_obs; k++) {
                    const long long idx = d->observation_grid[go++];
                    if (idx < 0 || idx > h->h_n_c[i]) {
                        delete h;
                        return fail(CB200_ERR_ARG, "interval %d: observation_grid entry %lld outside 0..%d", i + 1, idx, h->h_n_c[i]);
                    }
                    obs_var.push_back(idx == 0 ? -1 : h->h_var_off[i] + h->h_n_u0[i] + ntp + (int)idx - 1);
                }
            }
        }
        h->n_obs_rows = (int)obs_sample.size();
    }
    // upload
    std::vector<char> blob;
    const size_t o_obs_s = push(blob, obs_sample), o_obs_v = push(blob, obs_var);
    const size_t o_M = push(blob, h->h_M), o_po = push(blob, h->h_piece_off), o_t0 = push(blob, pt0),
                 o_t1 = push(blob, pt1), o_ph = push(blob, ph), o_ci = push(blob, h->h_cidx), o_vo = push(blob, h->h_var_off),
                 o_nu = push(blob, h->h_n_u0), o_nc = push(blob, h->h_n_c), o_is = push(blob, h->h_ic_src),
                 o_if = push(blob, ic_fix), o_if0 = push(blob, ic_fix0), o_se = push(blob, h->h_sens_off), o_tse = push(blob, h->h_tsens_off), o_so = push(blob, h->h_samp_off),
                 o_psamp = push(blob, psamp), o_nsm = push(blob, nsm), o_svn = push(blob, sv_n), o_svo = push(blob, sv_off), o_svt = push(blob, sv_t),
                 o_q = push(blob, h->quad0), o_dp = push(blob, dir_pcol), o_dperm = push(blob, dir_perm),
                 o_am = push(blob, act_mask), o_am1 = push(blob, act_mask1), o_dk = push(blob, h->dpos_kind), o_dst = push(blob, h->dpos_stage),
                 o_oo = push(blob, h->oth_off), o_ot = push(blob, h->oth), o_jp = push(blob, h->jpos);
    std::vector<double> mdata;
    if (d->model_data && d->model_data_len > 0) mdata.assign(d->model_data, d->model_data + d->model_data_len);
    const size_t o_md = push(blob, mdata);
    const size_t o_gr = (blob.size() + 15) & ~size_t(15);
    h->grad_cap = std::max(1, L.nvars);
    blob.resize(o_gr + sizeof(GradEntry) * (size_t)h->grad_cap);
    auto bail = [&](const char *what, cudaError_t e) {
        int rc = fail(CB200_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        cb200_destroy(h);
        return rc;
    };
    cudaError_t e = cudaMalloc(&h->tables, blob.size());
    if (e != cudaSuccess) return bail("cudaMalloc(tables)", e);
    e = cudaMemcpy(h->tables, blob.data(), blob.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail("cudaMemcpy(tables)", e);
    char *base = (char *)h->tables;
    h->d_obs_sample = (int *)(base + o_obs_s);
    h->d_obs_var = (int *)(base + o_obs_v);
    L.M = (const int *)(base + o_M);
    L.piece_off = (const int *)(base + o_po);
    L.pt0 = (const double *)(base + o_t0);
    L.pt1 = (const double *)(base + o_t1);
    L.ph = (const double *)(base + o_ph);
    // One step size for the whole layer when every piece has the same h = (t1 - t0)/K up to the rounding of
    // the grid points (a few ulp of the
    // grid points t0, t1 themselves): the reference's `adaptive = false, dt = h` is a single dt for
    // all pieces too.  The coefficient products h*a_ij then are kernel-parameter constants.
    double h_ref = ph[0];
    {
        std::vector<double> srt(ph);
        std::nth_element(srt.begin(), srt.begin() + srt.size() / 2, srt.end());
        h_ref = srt[srt.size() / 2];
    }
    L.uniform_h = 1;
    for (size_t q = 0; q < ph.size(); q++) {
        // t1 - t0 carries the rounding of the grid points themselves: a few ulp of max(|t0|, |t1|)
        const double tol = 4.0 * 2.220446049250313e-16 * std::max(fabs(pt0[q]), fabs(pt1[q])) / (double)d->steps_per_piece;
        L.uniform_h &= (fabs(ph[q] - h_ref) <= tol + 8.0 * 2.220446049250313e-16 * fabs(h_ref));
    }
    if (getenv("CB200_EXACT_PIECE_STEPS")) {
        L.uniform_h = 1;
        for (double hh : ph) L.uniform_h &= (hh == ph[0]);
        h_ref = ph[0];
    }
    for (int k = 0; k < 21; k++) L.hA[k] = 0.0;
    if (d->method == CB200_RK4) {
        rk4_coeffs(h_ref, L.hA);
    } else {
        tsit5_coeffs(h_ref, L.hA);
    }
    L.cidx = (const int *)(base + o_ci);
    L.var_off = (const int *)(base + o_vo);
    L.n_u0 = (const int *)(base + o_nu);
    L.n_c = (const int *)(base + o_nc);
    L.ic_src = (const int *)(base + o_is);
    L.ic_fix = (const double *)(base + o_if);
    L.ic_fix0 = (const double *)(base + o_if0);
    L.sens_off = (const long long *)(base + o_se);
    L.tsens_off = (const long long *)(base + o_tse);
    L.samp_off = (const int *)(base + o_so);
    L.psamp = (const int *)(base + o_psamp);
    L.nsm = (const int *)(base + o_nsm);
    L.sv_n = h->has_saveat ? (const int *)(base + o_svn) : nullptr;
    L.sv_off = (const int *)(base + o_svo);
    L.sv_t = (const double *)(base + o_svt);
    L.dir_pcol = (const signed char *)(base + o_dp);
    L.dir_perm = (const int *)(base + o_dperm);
    L.act_mask = (const unsigned *)(base + o_am);
    h->act_mask1 = act_mask1.empty() ? nullptr : (const unsigned *)(base + o_am1);
    L.dpos_kind = (const int *)(base + o_dk);
    L.dpos_stage = (const int *)(base + o_dst);
    L.oth_off = (const int *)(base + o_oo);
    L.oth = (const MatchEntry *)(base + o_ot);
    L.jpos = (const int *)(base + o_jp);
    L.model_data = mdata.empty() ? nullptr : (const double *)(base + o_md);
    h->d_quad = (int *)(base + o_q);
    h->d_grad = (GradEntry *)(base + o_gr);

    int dnx = 0;
    const int dslot = dense_slot(d->model, &dnx);
    if (dslot >= 0) {
        const long long want_len = (long long)dnx * dnx + dnx;
        if (!d->model_data || d->model_data_len != want_len) {
            cb200_destroy(h);
            return fail(CB200_ERR_ARG, "the dense %d-state model needs model_data of %lld doubles (W row-major, then b)", dnx, want_len);
        }
        {   // the thread-per-unit kernels read W, b from ONE __constant__ block per model and device: live handles of a
            // dense model on a device must share their data (the tensor-path kernel reads the handle's own table and has
            // no such limit)
            std::lock_guard<std::mutex> lk(g_d20_mu);
            D20Slot &slot = g_d20[dslot][d->device & 63];
            if (slot.refs > 0 && memcmp(slot.data, d->model_data, sizeof(double) * (size_t)want_len) != 0) {
                cb200_destroy(h);
                return fail(CB200_ERR_UNSUPPORTED, "another dense %d-state layer with different model_data is alive on device %d", dnx, d->device);
            }
            memcpy(slot.data, d->model_data, sizeof(double) * (size_t)want_len);
            slot.refs++;
            h->d20_ref = true;
        }
        int rc = dslot == 0   ? dense20_set_data(d->model_data, d->model_data_len)
                 : dslot == 1 ? dense16_set_data(d->model_data, d->model_data_len)
                 : dslot == 2 ? dense24_set_data(d->model_data, d->model_data_len)
                              : dense32_set_data(d->model_data, d->model_data_len);
        if (rc != 0) return bail("cudaMemcpyToSymbol(dense model data)", (cudaError_t)rc);
    }
    {   // accumulators of the cross-candidate sums, zero at rest (handle.cuh)
        const size_t n_acc = (size_t)(1 + L.nvars + h->fim_n * h->fim_n);
        e = cudaMalloc((void **)&h->acc, sizeof(double) * n_acc + 16);
        if (e != cudaSuccess) return bail("cudaMalloc(acc)", e);
        e = cudaMemset(h->acc, 0, sizeof(double) * n_acc + 16);
        if (e != cudaSuccess) return bail("cudaMemset(acc)", e);
        h->tail_counter = (unsigned *)(h->acc + n_acc);
        h->graphs_enabled = getenv("CB200_NO_GRAPH") == nullptr;
    }
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return bail("cudaStreamCreate", e);
    h->own_stream = true;
    for (int k = 0; k < cb200_handle::NAUX; k++) {
        e = cudaStreamCreateWithFlags(&h->aux[k], cudaStreamNonBlocking);
        if (e != cudaSuccess) return bail("cudaStreamCreate", e);
        e = cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming);
        if (e != cudaSuccess) return bail("cudaEventCreate", e);
    }
    e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e != cudaSuccess) return bail("cudaEventCreate", e);
    for (int k = 0; k < cb200_handle::EV_RING; k++) {
        e = cudaEventCreate(&h->ev0[k]);
        if (e != cudaSuccess) return bail("cudaEventCreate", e);
        e = cudaEventCreate(&h->ev1[k]);
        if (e != cudaSuccess) return bail("cudaEventCreate", e);
    }
    *out = h;
    return CB200_OK;
}

extern "C" int cb200_destroy(cb200_handle *h)
{
    NvtxRange nvtx_range("cb200_destroy");
    if (!h) return CB200_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    comm_release(h);
    for (SmallGraph &g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
    if (h->acc) cudaFree(h->acc);
    if (h->d20_ref) {
        std::lock_guard<std::mutex> lk(g_d20_mu);
        g_d20[dense_slot(h->model)][h->device & 63].refs--;
        h->d20_ref = false;
    }
    for (DevBuf *b : {&h->ws_ps, &h->ws_in, &h->ws_xend, &h->ws_sens, &h->ws_traj, &h->ws_dk, &h->ws_ds, &h->ws_obj,
                      &h->ws_grad, &h->ws_fim, &h->ws_finit, &h->ws_out, &h->ws_sums, &h->ws_objq, &h->ws_jac, &h->ws_crit, &h->ws_fixed, &h->ws_status, &h->ws_small, &h->ws_gradc, &h->ws_tsens})
        b->release();
    if (h->tables) cudaFree(h->tables);
    if (h->pin) cudaFreeHost(h->pin);
    for (int k = 0; k < cb200_handle::EV_RING; k++) {
        if (h->ev0[k]) cudaEventDestroy(h->ev0[k]);
        if (h->ev1[k]) cudaEventDestroy(h->ev1[k]);
    }
    for (int k = 0; k < cb200_handle::NAUX; k++) {
        if (h->aux[k]) {
            cudaStreamSynchronize(h->aux[k]);
            cudaStreamDestroy(h->aux[k]);
        }
        if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return CB200_OK;
}

extern "C" int cb200_get_sizes(const cb200_handle *h, cb200_sizes *s)
{
    if (!h || !s) return fail(CB200_ERR_ARG, "null argument");
    s->nvars = h->L.nvars;
    s->n_defects = (int64_t)h->defects.size();
    s->n_samples = h->n_samples;
    s->n_row = h->L.nx + h->L.nact;
    s->n_sens = h->n_sens;
    s->n_units = h->L.I;
    s->jac_nnz = (int64_t)h->jac_rows.size();
    s->jac_nnz_var = (int64_t)h->jac_nnz_var;
    s->n_traj_sens = (int64_t)h->n_tsens;
    return CB200_OK;
}

extern "C" int cb200_block_structure(const cb200_handle *h, int64_t *blocks)
{
    if (!h || !blocks) return fail(CB200_ERR_ARG, "null argument");
    for (int i = 0; i < h->L.I; i++) blocks[i] = h->h_var_off[i];
    blocks[h->L.I] = h->L.nvars;
    return CB200_OK;
}

extern "C" int cb200_set_stream(cb200_handle *h, void *stream)
{
    if (!h) return fail(CB200_ERR_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->own_stream && h->stream) {
        cudaStreamSynchronize(h->stream);
        cudaStreamDestroy(h->stream);
    }
    h->stream = (cudaStream_t)stream;
    h->own_stream = false;
    return CB200_OK;
}

extern "C" int cb200_set_interval_range(cb200_handle *h, int32_t first, int32_t last)
{
    if (!h) return fail(CB200_ERR_ARG, "null handle");
    std::lock_guard<std::mutex> lk(h->mu);
    const int I = h->L.I;
    if (first == 1 && last == I) {
        h->range_i0 = 0;
        h->range_n = -1;
        return CB200_OK;
    }
    if (first < 1 || last > I || last < first - 1)
        return fail(CB200_ERR_ARG, "interval range %d:%d outside 1:%d", first, last, I);
    h->range_i0 = first - 1;
    h->range_n = last - first + 1;   // first:first-1 is Julia's empty range
    return CB200_OK;
}

extern "C" int cb200_jac_structure(const cb200_handle *h, int64_t *rows, int64_t *cols, int64_t *src)
{
    if (!h) return fail(CB200_ERR_ARG, "null handle");
    for (size_t k = 0; k < h->jac_rows.size(); k++) {
        if (rows) rows[k] = h->jac_rows[k];
        if (cols) cols[k] = h->jac_cols[k];
        if (src) src[k] = h->jac_src[k];
    }
    return CB200_OK;
}

// Variables the objective row depends on: the contiguous range [first, first + n) of the flat ps vector (1-based
// `first` as Julia indexes): every variable for a quadrature row (it sums over the intervals), the last interval's
// block for any other state row.  Returns n, or a negative status.
extern "C" int64_t cb200_grad_structure(const cb200_handle *h, int32_t objective_row, int64_t *first)
{
    if (!h) return fail(CB200_ERR_ARG, "null handle");
    const int nx = h->L.nx, I = h->L.I;
    if (objective_row < 1 || objective_row > nx) return fail(CB200_ERR_ARG, "objective_row is not a state row");
    bool quad = false;
    for (int q : h->quad0) quad |= (q == objective_row - 1);
    const int f0 = quad ? 0 : h->h_var_off[I - 1];
    if (first) *first = f0 + 1;
    return h->L.nvars - f0;
}

extern "C" float cb200_last_kernel_ms(const cb200_handle *h)
{
    float ms = -1.0f;
    if (cb200_kernel_ms_history(h, &ms, 1) != 1) return -1.0f;
    return ms;
}

extern "C" int cb200_kernel_ms_history(const cb200_handle *h, float *ms, int n)
{
    if (!h || !ms || n <= 0) return 0;
    const long long have = std::min<long long>(h->n_timed, cb200_handle::EV_RING);
    const int m = (int)std::min<long long>(n, have);
    for (int k = 0; k < m; k++) {  // ms[0] = newest
        const int slot = (int)((h->n_timed - 1 - k) % cb200_handle::EV_RING);
        if (cudaEventSynchronize(h->ev1[slot]) != cudaSuccess) return k;
        if (cudaEventElapsedTime(&ms[k], h->ev0[slot], h->ev1[slot]) != cudaSuccess) return k;
    }
    return m;
}

extern "C" int64_t cb200_launch_count(const cb200_handle *h) { return h ? h->launches : 0; }

static inline unsigned nblk(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static int transpose(cb200_handle *h, cudaStream_t st, const double *in, long long ld_in, double *out, long long ld_out,
                     long long nr, long long nc)
{
    if (nr <= 0 || nc <= 0) return 0;
    // grid.y is limited to 65535 blocks: split rows
    const long long max_rows = 65535LL * 32;
    for (long long r0 = 0; r0 < nr; r0 += max_rows) {
        const long long rows = std::min(max_rows, nr - r0);
        dim3 grid((unsigned)((nc + 31) / 32), (unsigned)((rows + 31) / 32));
        transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in + r0 * ld_in, ld_in, out + r0, ld_out, rows, nc);
        h->launches++;
    }
    return (int)cudaGetLastError();
}

// what one evaluation produces, resolved from (want, out)
struct EvalPlan {
    bool xend, sens, traj, dk, ds, obj, grad, fim, osum, gsum, fsum, jac, crit, gradc, tsens;
    int gradc_first, gradc_n;   // compact gradient: variables [gradc_first, gradc_first + gradc_n)
    bool fused_grad, need_sens, need_xend, need_traj;
    int obj_state, obj_ctrl_var, obj_quad, objective_row;
    bool fused_fim;             // the shooting kernel does the Fisher read-out itself (OED models, FIM_END)
    bool reduce, gather;        // CB200_COMM_REDUCE / CB200_COMM_GATHER
    long long B_total, b0;      // all-gather: candidates of all ranks, this rank's first
    // small evaluation as ONE kernel (shoot.cuh small_tail): the shooting kernel's last block computes the objective and
    // copies the result block [st_src, st_src + st_n) to the caller's pinned mirror st_dst.  eval_core clears `fuse_small`
    // when the model has no such kernel and runs the usual pipeline
    int sm_reserve;             // chunks of the host pipeline: SMs a persistent kernel leaves to the neighbours' conversions
    mutable bool fuse_small;
    const double *st_src;
    double *st_dst;
    int st_n;
};

// where the kind-major defects of one batch (or chunk) go: the compact block of the evaluation, or this rank's columns
// of the all-gather window together with every peer's window (P2P transport)
struct GatherPlan {
    double *dk = nullptr;       // local destination, [pos * ldk + b]
    long long ldk = 0;
    double *const *peers = nullptr;
    long long peer_off = 0;
    int n_peers = 0, self = 0;
};

// device SoA buffers of one batch (or one chunk of it): [q * ldo + b]
struct DevSet {
    double *xend = nullptr, *sens = nullptr, *traj = nullptr, *dk = nullptr, *ds = nullptr, *obj = nullptr,
           *grad = nullptr, *fim = nullptr, *objq = nullptr, *jac = nullptr, *crit = nullptr, *gradc = nullptr, *tsens = nullptr;
    int *status = nullptr;      // [B] of the batch / chunk
    double *osum = nullptr, *gsum = nullptr, *fsum = nullptr;   // the handle's accumulators (shared by all chunks)
    const double *finit = nullptr;
    GatherPlan gp;
};

// OED models whose end-state Fisher read-out the shooting kernel can do in its epilogue (models.cuh: FIM_N, FIM_OFF)
static bool model_fuses_fim(int model, int *n, int *off)
{
    switch (model) {
    case CB200_MODEL_LOTKA2_OED:
    case CB200_MODEL_LOTKA2_OED_CU: *n = 2; *off = 6; return true;
    case CB200_MODEL_LIN1D_OED: *n = 1; *off = 2; return true;
    default: return false;
    }
}

// The device pipeline of one batch on one stream: K1/K2 with the fused K3 epilogue, then the small kernels.
// ldo: leading dimension of every result (>= B: a chunk writes its columns of the whole batch's buffers).
// tail: carried by the last kernel of the pipeline (nullptr: the caller runs the tail itself, after joining its chunks).
static int eval_core(cb200_handle *h, cudaStream_t st, const double *d_ps, long long d_ldb, long long B, long long ldo,
                     const EvalPlan &pl, const DevSet &d, const TailArgs *tail)
{
    const DevLayer &L0 = h->L;
    const int nx = L0.nx, I = L0.I;
    const long long nrow = nx + L0.nact;
    const long long ndef = (long long)h->defects.size();
    DevLayer L = L0;
    L.sm_reserve = pl.sm_reserve;
    int G = 0;
    if (pl.need_sens) {
        G = h->G_sens;
        if (G <= 0) return fail(CB200_ERR_UNSUPPORTED, "model %d has no sensitivity kernel", h->model);
        L.R = (h->max_ns + G - 1) / G;
        if (L.R < 1) L.R = 1;
    }
    const bool partial = h->range_n >= 0 && h->range_n < I;
    if (partial) {
        // a sub-range of the intervals (interval sharding across devices): everything the other intervals would
        // have written stays zero, so that the per-device results ASSEMBLE BY SUMMATION (one all-reduce)
        L.i0 = h->range_i0;
        L.In = h->range_n;
        const struct { double *p; long long n; } z[] = {
            {d.xend, (long long)nx * I}, {d.sens, h->n_sens}, {d.traj, nrow * h->n_samples}, {d.gp.dk, ndef}, {d.ds, ndef},
            {d.grad, L0.nvars}, {d.objq, I}, {d.jac, h->jac_nnz_var}, {d.gradc, pl.gradc_n}, {d.tsens, h->n_tsens}};
        for (const auto &e : z)
            if (e.p && e.n > 0) {
                const long long ld = (e.p == d.gp.dk) ? d.gp.ldk : ldo;
                CUDA_TRY(cudaMemset2DAsync(e.p, sizeof(double) * (size_t)ld, 0, sizeof(double) * (size_t)B, (size_t)e.n, st));
            }
        if (L.In == 0) {   // nothing to integrate on this device: the outputs are the zeros
            if (pl.obj && d.obj) CUDA_TRY(cudaMemsetAsync(d.obj, 0, sizeof(double) * (size_t)B, st));
            if (tail && (tail->nseg > 0 || tail->barrier)) {
                tail_kernel<<<1, 256, 0, st>>>(*tail);
                h->launches++;
            }
            return CB200_OK;
        }
    }
    ShootOut so{};
    so.xend = d.xend;
    so.sens = d.sens;
    so.traj = d.traj;
    so.tsens = d.tsens;
    so.ldo = ldo;
    so.dk = ndef > 0 ? d.gp.dk : nullptr;
    so.ldk = d.gp.ldk > 0 ? d.gp.ldk : ldo;
    so.dk_peers = d.gp.peers;
    so.dk_peer_off = d.gp.peer_off;
    so.n_dk_peers = (ndef > 0 && d.gp.dk) ? d.gp.n_peers : 0;
    so.dk_self = d.gp.self;
    so.dst = ndef > 0 ? d.ds : nullptr;
    so.objq = d.objq;
    so.grad = pl.fused_grad ? d.grad : nullptr;
    so.gradc = d.gradc;
    so.gradc_first = pl.gradc_first;
    so.grad_sum = pl.fused_grad ? d.gsum : nullptr;
    so.jac = d.jac;
    so.status = d.status;
    so.obj_state = pl.obj_state;
    so.obj_all = pl.obj_quad;
    if (pl.fused_fim) {
        so.fim_fused = 1;
        so.fim = d.fim;
        so.fsum = d.fsum;
        so.crit = d.crit;
        so.finit = d.finit;
    }
    const int ev_slot = (int)(h->n_timed % cb200_handle::EV_RING);
    CUDA_TRY(cudaEventRecord(h->ev0[ev_slot], st));
    auto launch = [&](const ShootOut &so_) {
        int r_ = -1000;
        if (G > 1 && B < 32 && h->act_mask1 && !h->no_g1) {   // latency-bound: one direction per thread, if the model has that kernel
            DevLayer L1 = L;
            L1.R = std::max(1, h->max_ns);
            L1.act_mask = h->act_mask1;
            r_ = launch_model(h->model, 1, h->method, L1, d_ps, d_ldb, B, so_, st);
            if (r_ == -1000) h->no_g1 = true;
        }
        if (r_ == -1000) r_ = launch_model(h->model, G, h->method, L, d_ps, d_ldb, B, so_, st);
        return r_;
    };
    if (pl.fuse_small && !h->no_small_tail) {
        so.st_counter = h->tail_counter;
        so.st_obj = ((pl.obj) && pl.obj_state >= 0) ? d.obj : nullptr;
        so.st_src = pl.st_src;
        so.st_dst = pl.st_dst;
        so.st_n = pl.st_n;
    } else {
        pl.fuse_small = false;
    }
    int rc = launch(so);
    if (rc == -1001) {   // the model has no kernel with the small-evaluation tail: the usual pipeline
        h->no_small_tail = true;
        pl.fuse_small = false;
        so.st_counter = nullptr;
        rc = launch(so);
    }
    if (rc == -1000)
        return fail(CB200_ERR_UNSUPPORTED, "model %d: no kernel for %d directions/thread, method %d", h->model, G,
                    h->method);
    if (rc) return fail(CB200_ERR_CUDA, "shoot kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    h->launches++;
    CUDA_TRY(cudaEventRecord(h->ev1[ev_slot], st));
    h->n_timed++;

    // ---- K3 remainder: control-row gradient, trajectory quadratures, FIM, objective sum over intervals (last: it
    //      carries the tail)
    // a control-row objective lives on the last interval: under interval sharding only its owner reports it
    const bool owns_last = !partial || (L.i0 + L.In == I);
    if ((pl.grad || pl.gsum) && !pl.fused_grad) {  // objective = a control row: the gradient is a unit vector
        if (d.grad) CUDA_TRY(cudaMemset2DAsync(d.grad, sizeof(double) * (size_t)ldo, 0, sizeof(double) * (size_t)B, (size_t)L0.nvars, st));
        if (owns_last) {
            grad_kernel<<<dim3(nblk(B, 128), 1), 128, 0, st>>>(h->d_grad, 1, nullptr, B, B, d.grad, ldo, d.gsum);
            h->launches++;
        }
    }
    if (pl.need_traj && h->nquad > 0 && I > 1) {
        traj_quad_kernel<<<nblk(B, 128), 128, 0, st>>>(d.traj, ldo, d.xend, ldo, B, nx, (int)nrow, I, L.samp_off, L.nsm,
                                                     h->d_quad, h->nquad);
        h->launches++;
    }
    const bool want_tail = tail && tail->counter && (tail->nseg > 0 || tail->barrier);
    TailArgs none{};
    const bool run_obj = (pl.obj || pl.osum) && !(pl.obj_state < 0 && !owns_last);
    const bool run_fim = (pl.fim || pl.fsum || pl.crit) && !pl.fused_fim;
    bool tail_done = false;
    if (run_fim) {
        const TailArgs &ta = (want_tail && !run_obj) ? *tail : none;
        tail_done = want_tail && !run_obj;
        if (h->fim_mode == CB200_FIM_END)
            fim_kernel<<<nblk(B, 128), 128, 0, st>>>(d.xend, ldo, B, nx, I, h->fim_off, h->fim_n, d.finit, d.fim, ldo, d.fsum, d.crit, ta);
        else
            fim_samples_kernel<<<nblk(B, 128), 128, 0, st>>>(d.traj, ldo, B, (int)nrow, h->n_samples, d_ps, d_ldb, h->fim_mode,
                                                           h->fim_bx, h->fim_n, h->n_obs, h->fim_off, h->n_obs_rows,
                                                           h->d_obs_sample, h->d_obs_var, d.finit, d.fim, ldo, d.fsum, d.crit, ta);
        h->launches++;
    }
    if (pl.fuse_small) {
        // objective and copy-out were the shooting kernel's tail
    } else if ((pl.obj || pl.osum) && !run_obj) {
        if (d.obj) CUDA_TRY(cudaMemsetAsync(d.obj, 0, sizeof(double) * (size_t)B, st));
    } else if (run_obj) {
        objective_kernel<<<nblk(B, 256), 256, 0, st>>>(d.objq, ldo, d_ps, d_ldb, B, I, pl.obj_state >= 0, pl.obj_quad,
                                                     pl.obj_ctrl_var, d.obj, d.osum, want_tail ? *tail : none);
        tail_done = want_tail;
        h->launches++;
    }
    if (want_tail && !tail_done) {
        tail_kernel<<<1, 256, 0, st>>>(*tail);
        h->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return CB200_OK;
}

// copy / transpose a quantity-major device block [n][ld_in] (B live columns) to its destination
static int deliver(cb200_handle *h, cudaStream_t st, const double *src, long long ld_in, long long n, long long B, double *dst,
                   bool dst_soa, bool dst_host)
{
    if (n <= 0 || B <= 0) return CB200_OK;
    const cudaMemcpyKind kind = dst_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (dst_soa) {
        if (ld_in == B)
            CUDA_TRY(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)n * (size_t)B, kind, st));
        else
            CUDA_TRY(cudaMemcpy2DAsync(dst, sizeof(double) * (size_t)B, src, sizeof(double) * (size_t)ld_in, sizeof(double) * (size_t)B,
                                       (size_t)n, kind, st));
        return CB200_OK;
    }
    if (dst_host) return fail(CB200_ERR_ARG, "internal: candidate-major host delivery goes through a staging buffer");
    const int rc = transpose(h, st, src, ld_in, dst, n, n, B);
    if (rc) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    return CB200_OK;
}

extern "C" int cb200_eval(cb200_handle *h, const double *ps, int64_t B, int64_t ldb, uint32_t want, uint32_t flags,
                          const cb200_out *out)
{
    NvtxRange nvtx_range("cb200_eval");
    if (!h || !out) return fail(CB200_ERR_ARG, "null argument");
    if (B < 0) return fail(CB200_ERR_ARG, "negative batch");
    const bool collective = (flags & (CB200_COMM_REDUCE | CB200_COMM_GATHER)) != 0;
    if (B == 0 && !collective) return CB200_OK;
    if (B == 0) return fail(CB200_ERR_UNSUPPORTED, "a collective evaluation needs at least one candidate on every rank");
    if (!ps && h->L.nvars > 0) return fail(CB200_ERR_ARG, "null ps");
    std::lock_guard<std::mutex> lk(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    const DevLayer &L0 = h->L;
    const int nx = L0.nx, I = L0.I, nvars = L0.nvars;
    bool dev = flags & CB200_MEM_DEVICE, ps_soa = flags & CB200_PS_SOA, out_soa = flags & CB200_OUT_SOA;
    if (ps_soa ? (ldb < B) : (ldb < nvars)) return fail(CB200_ERR_ARG, "ldb too small");
    if (B == 1) {   // one candidate (the reference's own use): both layouts coincide, no conversion kernels
        ps_soa = true;
        ldb = 1;
        if (!(flags & CB200_COMM_GATHER)) out_soa = true;
    }
    const long long nrow = nx + L0.nact;
    const long long ndef = (long long)h->defects.size();
    const int n2 = h->fim_n * h->fim_n;
    const uint32_t res = out->resident;
    auto on = [&](uint32_t bit, const void *p) { return (want & bit) && (p || (res & bit)); };

    EvalPlan pl{};
    pl.sens = on(CB200_WANT_SENS, out->sens);
    pl.grad = on(CB200_WANT_GRAD, out->grad);
    pl.traj = on(CB200_WANT_TRAJ, out->traj);
    pl.xend = on(CB200_WANT_XEND, out->xend);
    pl.dk = on(CB200_WANT_DEFECTS, out->defects_kind);
    pl.ds = (want & CB200_WANT_DEFECTS) && out->defects_stage;
    pl.obj = on(CB200_WANT_OBJ, out->objective);
    pl.fim = on(CB200_WANT_FIM, out->fim);
    pl.fsum = (want & CB200_WANT_FIM) && out->fim_sum;
    pl.osum = (want & CB200_WANT_OBJ) && out->objective_sum;
    pl.gsum = (want & CB200_WANT_GRAD) && out->grad_sum;
    pl.jac = on(CB200_WANT_JAC, out->jac_vals);
    pl.crit = on(CB200_WANT_CRIT, out->criteria);
    pl.gradc = on(CB200_WANT_GRAD_COMPACT, out->grad_compact);
    pl.tsens = on(CB200_WANT_TRAJ_SENS, out->traj_sens);
    pl.reduce = (flags & CB200_COMM_REDUCE) != 0;
    pl.gather = (flags & CB200_COMM_GATHER) != 0;
    if (collective && !h->comm) return fail(CB200_ERR_ARG, "CB200_COMM_REDUCE / CB200_COMM_GATHER need a communicator (cb200_comm_init)");
    if (pl.gather) {
        if (!(want & CB200_WANT_DEFECTS) || ndef == 0) return fail(CB200_ERR_ARG, "CB200_COMM_GATHER needs CB200_WANT_DEFECTS on a layer with shooting constraints");
        pl.B_total = out->comm_B_total;
        pl.b0 = out->comm_b0;
    }
    if ((pl.fim || pl.fsum || pl.crit) && h->fim_off < 0) return fail(CB200_ERR_ARG, "layer has no Fisher block (fim_offset == 0)");
    if (pl.crit && h->fim_n > CB200_MAX_FIM) return fail(CB200_ERR_UNSUPPORTED, "criteria need fim_n <= %d", CB200_MAX_FIM);
    const bool partial = h->range_n >= 0 && h->range_n < I;
    if (partial) {
        if (pl.fim || pl.fsum || pl.crit)
            return fail(CB200_ERR_UNSUPPORTED, "the Fisher read-out needs every interval on one device: reset cb200_set_interval_range");
        if (pl.traj && h->nquad > 0 && I > 1)
            return fail(CB200_ERR_UNSUPPORTED, "trajectory samples with quadrature running sums need every interval on one device");
    }
    pl.obj_state = pl.obj_ctrl_var = -1;
    pl.objective_row = out->objective_row;
    if (pl.obj || pl.grad || pl.osum || pl.gsum || pl.gradc) {
        const int row = out->objective_row;
        if (row < 1 || row > nrow) return fail(CB200_ERR_ARG, "objective_row out of range");
        if (row <= nx) {
            pl.obj_state = row - 1;
            for (int q : h->quad0) pl.obj_quad |= (q == pl.obj_state);
        } else {
            const int r = row - nx - 1;
            pl.obj_ctrl_var = h->h_var_off[I - 1] + h->h_n_u0[I - 1] + L0.ntp +
                              h->h_cidx[(size_t)(h->h_piece_off[I - 1] + h->h_M[I - 1] - 1) * L0.nact + r];
        }
    }
    pl.fused_grad = (pl.grad || pl.gsum) && pl.obj_state >= 0;   // gradient written by the shooting kernel
    if (pl.gradc) {
        if (pl.obj_state < 0) return fail(CB200_ERR_UNSUPPORTED, "grad_compact needs a state row as objective (a control row's gradient is a unit vector)");
        pl.gradc_first = pl.obj_quad ? 0 : h->h_var_off[I - 1];   // a quadrature row sums over all intervals
        pl.gradc_n = nvars - pl.gradc_first;
    }
    pl.need_sens = pl.sens || pl.fused_grad || pl.jac || pl.gradc || pl.tsens;   // sensitivity directions are integrated
    const bool any_fim = pl.fim || pl.fsum || pl.crit;
    {
        int fn = 0, fo = 0;
        pl.fused_fim = any_fim && h->fim_mode == CB200_FIM_END && !partial && model_fuses_fim(h->model, &fn, &fo) &&
                       fn == h->fim_n && fo == h->fim_off && !getenv("CB200_NO_FUSED_FIM");
    }
    pl.need_traj = pl.traj || (any_fim && h->fim_mode != CB200_FIM_END);   // sample-based read-outs consume the samples
    pl.need_xend = pl.xend || (any_fim && h->fim_mode == CB200_FIM_END && !pl.fused_fim) || (pl.need_traj && h->nquad > 0);
    const bool need_objq = (pl.obj || pl.osum) && pl.obj_state >= 0;
    if ((pl.grad || pl.gsum) && !pl.fused_grad) {
        h->h_grad.assign(1, GradEntry{pl.obj_ctrl_var, -1});
        CUDA_TRY(cudaMemcpyAsync(h->d_grad, h->h_grad.data(), sizeof(GradEntry), cudaMemcpyHostToDevice, h->stream));
    }
    int rc = comm_check(h);
    if (rc) return rc;

    // per-candidate sizes (doubles) of every SoA result; user == nullptr: the result stays in the workspace
    struct Item { bool on; long long n; DevBuf *ws; double *user; double *DevSet::*slot; uint32_t bit; };
    auto usr = [&](uint32_t bit, double *p) { return (res & bit) ? nullptr : p; };
    const Item items[] = {
        {pl.need_xend, (long long)nx * I, &h->ws_xend, pl.xend ? usr(CB200_WANT_XEND, out->xend) : nullptr, &DevSet::xend, CB200_WANT_XEND},
        {pl.sens, h->n_sens, &h->ws_sens, usr(CB200_WANT_SENS, out->sens), &DevSet::sens, CB200_WANT_SENS},
        {pl.need_traj, nrow * h->n_samples, &h->ws_traj, pl.traj ? usr(CB200_WANT_TRAJ, out->traj) : nullptr, &DevSet::traj, CB200_WANT_TRAJ},
        {pl.dk && !pl.gather, ndef, &h->ws_dk, usr(CB200_WANT_DEFECTS, out->defects_kind), &DevSet::dk, CB200_WANT_DEFECTS},
        {pl.ds, ndef, &h->ws_ds, out->defects_stage, &DevSet::ds, 0u},
        {pl.obj, 1, &h->ws_obj, usr(CB200_WANT_OBJ, out->objective), &DevSet::obj, CB200_WANT_OBJ},
        {pl.grad, nvars, &h->ws_grad, usr(CB200_WANT_GRAD, out->grad), &DevSet::grad, CB200_WANT_GRAD},
        {pl.fim, n2, &h->ws_fim, usr(CB200_WANT_FIM, out->fim), &DevSet::fim, CB200_WANT_FIM},
        {need_objq, I, &h->ws_objq, nullptr, &DevSet::objq, 0u},
        {pl.jac, h->jac_nnz_var, &h->ws_jac, usr(CB200_WANT_JAC, out->jac_vals), &DevSet::jac, CB200_WANT_JAC},
        {pl.crit, 6, &h->ws_crit, usr(CB200_WANT_CRIT, out->criteria), &DevSet::crit, CB200_WANT_CRIT},
        {pl.gradc, pl.gradc_n, &h->ws_gradc, usr(CB200_WANT_GRAD_COMPACT, out->grad_compact), &DevSet::gradc, CB200_WANT_GRAD_COMPACT},
        {pl.tsens, h->n_tsens, &h->ws_tsens, usr(CB200_WANT_TRAJ_SENS, out->traj_sens), &DevSet::tsens, CB200_WANT_TRAJ_SENS},
    };
    h->resident_valid = 0;
    h->last_B = B;
    h->gradc_n_last = pl.gradc_n;

    // ---- cross-candidate sums (expectation-type objective / gradient, FIM): the kernels accumulate into the handle's
    //      block (zero at rest); the tail of the evaluation's last kernel hands them on and clears the block
    if (h->acc_dirty) {
        CUDA_TRY(cudaMemsetAsync(h->acc, 0, sizeof(double) * (size_t)(1 + nvars + n2), h->stream));
        CUDA_TRY(cudaMemsetAsync(h->tail_counter, 0, sizeof(unsigned), h->stream));
        h->acc_dirty = false;
    }
    DevSet shared{};
    shared.osum = pl.osum ? h->acc : nullptr;
    shared.gsum = pl.gsum ? h->acc + 1 : nullptr;
    shared.fsum = pl.fsum ? h->acc + 1 + nvars : nullptr;
    TailArgs tail{};
    tail.counter = h->tail_counter;
    tail.acc = h->acc;
    auto add_seg = [&](bool want_it, int off, int len, double *dst) {
        if (want_it) tail.seg[tail.nseg++] = TailSeg{off, len, dst};
    };
    if (out->fim_init && any_fim) {
        if (h->ws_finit.reserve(sizeof(double) * n2)) return fail(CB200_ERR_NOMEM, "device out of memory");
        CUDA_TRY(cudaMemcpyAsync(h->ws_finit.p, out->fim_init, sizeof(double) * n2,
                                 dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        shared.finit = (const double *)h->ws_finit.p;
    }
    // ---- communicator: open the epoch; all-gathered defects go straight into the windows on the P2P transport
    ShootOut comm_so{};
    double *gather_local = nullptr;   // this rank's columns inside its own window (P2P), [pos * B_total + b]
    if (collective) {
        rc = comm_begin(h, pl.reduce, pl.gather, B, pl.B_total, pl.b0, &tail, &comm_so, &gather_local);
        if (rc) return rc;
    }
    const bool nccl_path = collective && !comm_is_p2p(h);
    TailSeg real_seg[3];
    int n_real = 0;
    double *red_stage = nullptr;
    long long n_stage = 0;
    // set up the tail's destinations: dst_* are device pointers
    auto plan_tail = [&](double *dst_o, double *dst_g, double *dst_f) -> int {
        tail.nseg = 0;
        add_seg(pl.osum, 0, 1, dst_o);
        add_seg(pl.gsum, 1, nvars, dst_g);
        add_seg(pl.fsum, 1 + nvars, n2, dst_f);
        if (nccl_path && pl.reduce && tail.nseg > 0) {   // NCCL transport: the local sums are staged, all-reduced, then delivered
            n_stage = 0;
            for (int s = 0; s < tail.nseg; s++) n_stage += tail.seg[s].len;
            red_stage = comm_red_stage(h, n_stage);
            if (!red_stage) return fail(CB200_ERR_NOMEM, "device out of memory");
            long long pos = 0;
            n_real = tail.nseg;
            for (int s = 0; s < tail.nseg; s++) {
                real_seg[s] = tail.seg[s];
                tail.seg[s].dst = red_stage + pos;
                pos += tail.seg[s].len;
            }
        }
        return CB200_OK;
    };
    // the gather plan of a batch whose first candidate is column c0 of the evaluation
    auto gather_plan = [&](double *compact, long long ld_compact, long long c0) {
        GatherPlan g;
        if (pl.gather && gather_local) {
            g.dk = gather_local + c0;
            g.ldk = pl.B_total;
            g.peers = comm_so.dk_peers;
            g.peer_off = comm_so.dk_peer_off + c0;
            g.n_peers = comm_so.n_dk_peers;
            g.self = comm_so.dk_self;
        } else {
            g.dk = compact ? compact + c0 : nullptr;
            g.ldk = ld_compact;
        }
        return g;
    };
    // after the pipeline: NCCL collectives, then the all-gathered defects to the caller
    auto finish_comm = [&](cudaStream_t st, const double *dk_compact, bool to_host) -> int {
        if (nccl_path) {
            const int r2 = comm_finish_nccl(h, st, red_stage, n_stage, real_seg, n_real, pl.gather ? dk_compact : nullptr, ndef, B);
            if (r2) return r2;
        }
        if (pl.gather && out->defects_gathered) {
            const double *g = comm_gather_ptr(h);
            if (out_soa || !to_host) return deliver(h, st, g, pl.B_total, ndef, pl.B_total, out->defects_gathered, out_soa, to_host);
            // candidate-major host matrix: transpose on the device, one contiguous copy back
            if (h->ws_out.reserve(sizeof(double) * (size_t)ndef * (size_t)pl.B_total)) return fail(CB200_ERR_NOMEM, "device out of memory (staging)");
            const int r3 = transpose(h, st, g, pl.B_total, (double *)h->ws_out.p, ndef, ndef, pl.B_total);
            if (r3) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)r3));
            CUDA_TRY(cudaMemcpyAsync(out->defects_gathered, h->ws_out.p, sizeof(double) * (size_t)ndef * (size_t)pl.B_total,
                                     cudaMemcpyDeviceToHost, st));
        }
        return CB200_OK;
    };
    // compact local defect block: needed without an all-gather, and as the NCCL all-gather's send buffer
    const bool need_dk_compact = pl.gather && (nccl_path || !gather_local);
    h->acc_dirty = pl.osum || pl.gsum || pl.fsum;

    // =================================================================== device pointers: one pass
    if (dev) {
        const bool direct = out_soa;   // kernels write straight into the caller's buffers
        const double *d_ps = ps;
        long long d_ldb = ldb;
        if (nvars > 0 && !ps_soa) {
            if (h->ws_ps.reserve(sizeof(double) * (size_t)nvars * (size_t)B)) return fail(CB200_ERR_NOMEM, "device out of memory (ps)");
            rc = transpose(h, h->stream, ps, ldb, (double *)h->ws_ps.p, B, B, nvars);
            if (rc) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)rc));
            d_ps = (const double *)h->ws_ps.p;
            d_ldb = B;
        }
        DevSet d = shared;
        for (const Item &it : items) {
            if (!it.on) continue;
            if (it.user && direct) {
                d.*(it.slot) = it.user;
            } else {
                if (it.ws->reserve(sizeof(double) * (size_t)it.n * (size_t)B))
                    return fail(CB200_ERR_NOMEM, "device out of memory for result buffers (B=%lld)", (long long)B);
                d.*(it.slot) = (double *)it.ws->p;
                if (!it.user) h->resident_valid |= it.bit;
            }
        }
        if (need_dk_compact) {
            if (h->ws_dk.reserve(sizeof(double) * (size_t)ndef * (size_t)B)) return fail(CB200_ERR_NOMEM, "device out of memory");
            d.dk = (double *)h->ws_dk.p;
        }
        d.gp = gather_plan(d.dk, B, 0);
        if (out->status) {   // int32 per candidate, the caller's device buffer
            d.status = out->status;
            CUDA_TRY(cudaMemsetAsync(d.status, 0, sizeof(int) * (size_t)B, h->stream));
        }
        rc = plan_tail(out->objective_sum, out->grad_sum, out->fim_sum);
        if (rc) return rc;
        rc = eval_core(h, h->stream, d_ps, d_ldb, B, B, pl, d, &tail);
        if (rc) return rc;
        h->acc_dirty = false;
        if (!direct) {
            for (const Item &it : items) {
                if (!it.on || !it.user) continue;
                rc = transpose(h, h->stream, d.*(it.slot), B, it.user, it.n, it.n, B);
                if (rc) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)rc));
            }
        }
        if (pl.gather && gather_local && pl.dk && out->defects_kind && !(res & CB200_WANT_DEFECTS)) {   // the local view as well
            rc = deliver(h, h->stream, gather_local, pl.B_total, ndef, B, out->defects_kind, direct, false);
            if (rc) return rc;
        }
        if (collective) {
            rc = finish_comm(h->stream, d.dk, false);
            if (rc) return rc;
        }
        return CB200_OK;   // enqueued on the handle's stream
    }

    // =================================================================== host pointers, small evaluation
    // (one candidate = the reference's own use, or a small variable-major batch): pageable caller arrays would make
    // every cudaMemcpyAsync a synchronous staged copy, so the inputs go through a pinned mirror with ONE H2D copy,
    // every result lands in ONE device buffer that comes back with ONE D2H copy, and the CPU scatters it.
    if (!pl.gather) {
        size_t out_doubles = 0;
        for (const Item &it : items)
            if (it.on && it.user) out_doubles += (size_t)it.n * (size_t)B;
        const size_t sums_doubles = (size_t)(pl.osum ? 1 : 0) + (pl.gsum ? (size_t)nvars : 0) + (pl.fsum ? (size_t)n2 : 0);
        const size_t in_doubles = (size_t)nvars * (size_t)B;
        const size_t status_doubles = out->status ? ((size_t)B * sizeof(int) + sizeof(double) - 1) / sizeof(double) : 0;
        const size_t tot = sizeof(double) * (out_doubles + sums_doubles + in_doubles + status_doubles);
        if (ps_soa && out_soa && ldb == B && tot <= cb200_handle::PIN_BYTES) {
            if (!h->pin) CUDA_TRY(cudaHostAlloc(&h->pin, cb200_handle::PIN_BYTES, cudaHostAllocMapped | cudaHostAllocPortable));   // the small-evaluation tail writes it from the device
            if (h->ws_small.reserve(tot)) return fail(CB200_ERR_NOMEM, "device out of memory");
            double *hp = (double *)h->pin, *dp = (double *)h->ws_small.p;
            // layout (doubles): [inputs | status ints (zeroed on the host: they travel with the inputs) | results ... | sums]
            if (in_doubles) memcpy(hp, ps, sizeof(double) * in_doubles);
            if (status_doubles) memset(hp + in_doubles, 0, sizeof(double) * status_doubles);
            DevSet d = shared;
            size_t off = in_doubles + status_doubles;
            for (const Item &it : items) {
                if (!it.on) continue;
                if (it.user) {
                    d.*(it.slot) = dp + off;
                    off += (size_t)it.n * (size_t)B;
                } else {
                    if (it.ws->reserve(sizeof(double) * (size_t)it.n * (size_t)B)) return fail(CB200_ERR_NOMEM, "device out of memory");
                    d.*(it.slot) = (double *)it.ws->p;
                    h->resident_valid |= it.bit;
                }
            }
            d.gp = gather_plan(d.dk, B, 0);
            double *dst_o = nullptr, *dst_g = nullptr, *dst_f = nullptr;
            if (pl.osum) { dst_o = dp + off; off += 1; }
            if (pl.gsum) { dst_g = dp + off; off += (size_t)nvars; }
            if (pl.fsum) { dst_f = dp + off; off += (size_t)n2; }
            if (out->status) d.status = (int *)(dp + in_doubles);
            rc = plan_tail(dst_o, dst_g, dst_f);
            if (rc) return rc;
            const size_t back = sizeof(double) * (off - in_doubles);
            // one kernel instead of kernel + objective kernel + copy node when the pipeline is just those: a few candidates
            // (flat launch), a state row as objective (or none), no cross-candidate sums, no read-out / quadrature kernels
            pl.fuse_small = !collective && !partial && B < 32 && sums_doubles == 0 && !any_fim && !pl.tsens &&
                            !(pl.need_traj && h->nquad > 0 && I > 1) && !((pl.grad || pl.gsum) && !pl.fused_grad) &&
                            (!(pl.obj || pl.osum) || pl.obj_state >= 0) && h->method != CB200_TSIT5_ADAPTIVE && !h->has_saveat &&
                            back > 0 && back / sizeof(double) < (1u << 30) && !getenv("CB200_NO_SMALL_TAIL");
            pl.st_src = dp + in_doubles;
            pl.st_dst = hp + in_doubles;
            pl.st_n = (int)(off - in_doubles);
            // the whole evaluation as one CUDA graph (captured once per shape of the request): a single launch instead
            // of two copies and two to four kernels -- the call is latency-bound (BASELINE config 1)
            SmallGraph *sg = nullptr;
            const unsigned gkey = flags | (pl.fused_fim ? 0x80000000u : 0u);
            const bool graphable = h->graphs_enabled && !collective && !partial && !(out->fim_init && any_fim) && !h->comm;
            if (graphable) {
                if (!h->graphs.empty() && h->graphs.front().gen != devbuf_generation()) {   // a workspace moved: the graphs hold stale pointers
                    for (SmallGraph &g : h->graphs)
                        if (g.exec) cudaGraphExecDestroy(g.exec);
                    h->graphs.clear();
                }
                for (SmallGraph &g : h->graphs)
                    if (g.want == want && g.objective_row == out->objective_row && g.B == B && g.has_status == (out->status != nullptr) &&
                        g.tot == tot && g.dp == (void *)dp && g.res == res && g.flags == gkey)
                        sg = &g;
            }
            if (sg) {
                CUDA_TRY(cudaGraphLaunch(sg->exec, h->stream));
                h->launches += sg->launches;
            } else {
                const bool capture = graphable && h->graphs.size() < 16;
                const long long launches0 = h->launches, timed0 = h->n_timed;
                if (capture) CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                cudaError_t ce = cudaSuccess;
                if (in_doubles + status_doubles)
                    ce = cudaMemcpyAsync(dp, hp, sizeof(double) * (in_doubles + status_doubles), cudaMemcpyHostToDevice, h->stream);
                if (ce == cudaSuccess) {
                    rc = eval_core(h, h->stream, in_doubles ? dp : nullptr, B, B, B, pl, d, &tail);
                    if (rc == CB200_OK && back && !pl.fuse_small)   // (fused: the kernel's tail wrote the mirror itself)
                        ce = cudaMemcpyAsync(hp + in_doubles, dp + in_doubles, back, cudaMemcpyDeviceToHost, h->stream);
                }
                if (capture) {
                    cudaGraph_t graph = nullptr;
                    const cudaError_t ee = cudaStreamEndCapture(h->stream, &graph);
                    h->n_timed = timed0;   // event records inside a capture carry no timing
                    if (rc == CB200_OK && ce == cudaSuccess && ee == cudaSuccess && graph) {
                        SmallGraph g;
                        g.want = want; g.objective_row = out->objective_row; g.B = B; g.has_status = out->status != nullptr;
                        g.tot = tot; g.dp = dp; g.res = res; g.flags = gkey; g.launches = h->launches - launches0;
                        g.gen = devbuf_generation();
                        if (cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
                            h->graphs.push_back(g);
                            ce = cudaGraphLaunch(g.exec, h->stream);
                        } else {
                            ce = cudaErrorUnknown;
                        }
                    } else if (ee != cudaSuccess && ce == cudaSuccess) {
                        ce = ee;
                    }
                    if (graph) cudaGraphDestroy(graph);
                    if (ce != cudaSuccess || rc != CB200_OK) {   // capture not possible here: never try again on this handle
                        cudaGetLastError();
                        h->graphs_enabled = false;
                        if (rc == CB200_OK) return fail(CB200_ERR_CUDA, "graph capture of the small evaluation failed: %s", cudaGetErrorString(ce));
                    }
                }
                if (rc) return rc;
                if (ce != cudaSuccess) return fail(CB200_ERR_CUDA, "small evaluation failed: %s", cudaGetErrorString(ce));
            }
            if (collective) {
                rc = finish_comm(h->stream, d.dk, true);
                if (rc) return rc;
                if (nccl_path && n_real) {   // the reduced sums were delivered to dp after the D2H copy above: fetch them
                    const size_t so_off = off - sums_doubles;
                    CUDA_TRY(cudaMemcpyAsync(hp + so_off, dp + so_off, sizeof(double) * sums_doubles, cudaMemcpyDeviceToHost, h->stream));
                }
            }
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            h->acc_dirty = false;
            if (out->status) memcpy(out->status, hp + in_doubles, sizeof(int) * (size_t)B);
            size_t o2 = in_doubles + status_doubles;
            for (const Item &it : items) {
                if (!it.on || !it.user) continue;
                memcpy(it.user, hp + o2, sizeof(double) * (size_t)it.n * (size_t)B);
                o2 += (size_t)it.n * (size_t)B;
            }
            if (pl.osum) { *out->objective_sum = hp[o2]; o2 += 1; }
            if (pl.gsum) { memcpy(out->grad_sum, hp + o2, sizeof(double) * (size_t)nvars); o2 += (size_t)nvars; }
            if (pl.fsum) { memcpy(out->fim_sum, hp + o2, sizeof(double) * (size_t)n2); o2 += (size_t)n2; }
            return comm_check(h);
        }
    }

    // =================================================================== host pointers: chunked, pipelined
    // The batch is cut into chunks of candidates; chunk c runs on stream c % NS: host->device copy, layout
    // conversion, the kernels, conversion back and device->host copies.  With pinned caller buffers the copy
    // engines (one per direction) and the SMs overlap across chunks; the caller's buffers are chunk-contiguous
    // in the candidate-major (Julia) layout, so every copy is one contiguous cudaMemcpyAsync.  On the device every
    // buffer holds the whole batch quantity-major, [q * B + b]; a chunk works on its columns.
    long long Bc = B;
    if (!ps_soa && !out_soa && B >= 1024) {
        Bc = ((B + 7) / 8 + 31) / 32 * 32;   // ~8 chunks, multiples of a warp
        if (Bc < 256) Bc = 256;
    } else if (!ps_soa && !out_soa && B >= 128 && h->n_sens >= 20000 && pl.need_sens) {
        // fewer candidates, but heavy ones (a rank's shard of BASELINE config 5: 512 candidates x 122 400 sensitivity
        // entries, milliseconds of kernel per chunk): the copies of one chunk still hide behind the kernel of the next
        Bc = ((B + 7) / 8 + 31) / 32 * 32;
        if (Bc < 64) Bc = 64;
    }
    if (const char *e = getenv("CB200_HOST_CHUNK")) {
        const long long v = atoll(e);
        if (v > 0 && !ps_soa && !out_soa) Bc = std::min<long long>(B, (v + 31) / 32 * 32);
    }
    const int nchunks = (int)((B + Bc - 1) / Bc);
    // A chunk's persistent kernel (the dense tensor-path kernel: 512 threads x 128 registers = the whole register file of every
    // SM) would keep the layout conversions of the chunks before and after it off the GPU until it ends, and with them their
    // copies: the chunks would run one after the other.  A few SMs stay free for them (tools/e2e_probe.py).
    if (nchunks > 1) {
        static const int reserve_env = getenv("CB200_SM_RESERVE") ? atoi(getenv("CB200_SM_RESERVE")) : -1;
        pl.sm_reserve = reserve_env >= 0 ? reserve_env : 1;   // e2e_probe: 0 -> +24.5 ms on 170.7 ms, 1 -> +5.6, 2 -> +7.0, 3 -> +8.1, 4 -> +9.7, 8 -> +14.5
    }
    if (nvars > 0) {
        const size_t bytes = sizeof(double) * (size_t)nvars * (size_t)B;
        if (h->ws_ps.reserve(bytes)) return fail(CB200_ERR_NOMEM, "device out of memory (ps, %zu bytes)", bytes);
        if (!ps_soa && h->ws_in.reserve(bytes)) return fail(CB200_ERR_NOMEM, "device out of memory (staging)");
    }
    size_t stage_per_cand = 0;
    for (const Item &it : items) {
        if (!it.on) continue;
        if (it.ws->reserve(sizeof(double) * (size_t)it.n * (size_t)B))
            return fail(CB200_ERR_NOMEM, "device out of memory for result buffers (B=%lld)", (long long)B);
        if (it.user && !out_soa) stage_per_cand += (size_t)it.n;
        if (!it.user) h->resident_valid |= it.bit;
    }
    const bool local_dk_from_window = pl.gather && gather_local && pl.dk && out->defects_kind && !(res & CB200_WANT_DEFECTS);
    if (local_dk_from_window && !out_soa) stage_per_cand += (size_t)ndef;
    if (need_dk_compact && h->ws_dk.reserve(sizeof(double) * (size_t)ndef * (size_t)B)) return fail(CB200_ERR_NOMEM, "device out of memory");
    {   // staging for the candidate-major results of the chunks and, after them, of the all-gathered defects
        size_t stage = stage_per_cand * (size_t)B;
        if (pl.gather && out->defects_gathered && !out_soa) stage = std::max(stage, (size_t)ndef * (size_t)pl.B_total);
        if (stage && h->ws_out.reserve(sizeof(double) * stage)) return fail(CB200_ERR_NOMEM, "device out of memory (staging)");
    }
    const size_t sums_doubles = (size_t)(pl.osum ? 1 : 0) + (pl.gsum ? (size_t)nvars : 0) + (pl.fsum ? (size_t)n2 : 0);
    double *dst_o = nullptr, *dst_g = nullptr, *dst_f = nullptr;
    if (sums_doubles) {
        if (h->ws_sums.reserve(sizeof(double) * (size_t)(nvars + 1 + n2))) return fail(CB200_ERR_NOMEM, "device out of memory");
        double *base = (double *)h->ws_sums.p;
        dst_o = pl.osum ? base : nullptr;
        dst_g = pl.gsum ? base + 1 : nullptr;
        dst_f = pl.fsum ? base + 1 + nvars : nullptr;
    }
    rc = plan_tail(dst_o, dst_g, dst_f);
    if (rc) return rc;

    if (out->status) {
        if (h->ws_status.reserve(sizeof(int) * (size_t)B)) return fail(CB200_ERR_NOMEM, "device out of memory");
        CUDA_TRY(cudaMemsetAsync(h->ws_status.p, 0, sizeof(int) * (size_t)B, h->stream));
    }
    cudaStream_t streams[cb200_handle::NAUX + 1];
    int ns_used = 1;
    streams[0] = h->stream;
    if (nchunks > 1) {
        for (int k = 0; k < cb200_handle::NAUX; k++) streams[ns_used++] = h->aux[k];
        CUDA_TRY(cudaEventRecord(h->ev_fork, h->stream));   // memsets / tables above are ordered before the chunks
        for (int k = 1; k < ns_used; k++) CUDA_TRY(cudaStreamWaitEvent(streams[k], h->ev_fork, 0));
    }
    double *dk_compact = need_dk_compact ? (double *)h->ws_dk.p : nullptr;
    for (int c = 0; c < nchunks; c++) {
        const long long b0 = (long long)c * Bc, bn = std::min(Bc, B - b0);
        cudaStream_t st = streams[c % ns_used];
        const double *d_ps = nullptr;
        if (nvars > 0) {
            double *soa = (double *)h->ws_ps.p + b0;   // the chunk's columns of [nvars][B]
            if (ps_soa) {  // host, variable-major (single chunk)
                CUDA_TRY(cudaMemcpy2DAsync(soa, sizeof(double) * (size_t)B, ps + b0, sizeof(double) * (size_t)ldb, sizeof(double) * (size_t)bn,
                                           nvars, cudaMemcpyHostToDevice, st));
            } else {       // host, candidate-major (Julia nvars x B matrix)
                double *stg = (double *)h->ws_in.p + (size_t)b0 * nvars;
                if (ldb == nvars)
                    CUDA_TRY(cudaMemcpyAsync(stg, ps + (size_t)b0 * ldb, sizeof(double) * (size_t)nvars * (size_t)bn,
                                             cudaMemcpyHostToDevice, st));
                else
                    CUDA_TRY(cudaMemcpy2DAsync(stg, sizeof(double) * nvars, ps + (size_t)b0 * ldb, sizeof(double) * ldb,
                                               sizeof(double) * nvars, bn, cudaMemcpyHostToDevice, st));
                rc = transpose(h, st, stg, nvars, soa, B, bn, nvars);
                if (rc) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)rc));
            }
            d_ps = soa;
        }
        DevSet d = shared;
        for (const Item &it : items)
            if (it.on) d.*(it.slot) = (double *)it.ws->p + b0;
        if (dk_compact) d.dk = dk_compact + b0;
        d.gp = gather_plan(d.dk ? d.dk - b0 : nullptr, B, b0);
        if (out->status) d.status = (int *)h->ws_status.p + b0;
        rc = eval_core(h, st, d_ps, B, bn, B, pl, d, nchunks == 1 ? &tail : nullptr);
        if (rc) return rc;
        if (out->status)
            CUDA_TRY(cudaMemcpyAsync(out->status + b0, d.status, sizeof(int) * (size_t)bn, cudaMemcpyDeviceToHost, st));
        size_t stage_off = (size_t)b0 * stage_per_cand;
        auto send_back = [&](const double *src, long long ld_in, long long n, double *user) -> int {
            if (out_soa) {  // quantity-major host output (single chunk)
                CUDA_TRY(cudaMemcpy2DAsync(user + b0, sizeof(double) * (size_t)B, src, sizeof(double) * (size_t)ld_in,
                                           sizeof(double) * (size_t)bn, (size_t)n, cudaMemcpyDeviceToHost, st));
            } else {
                double *stg = (double *)h->ws_out.p + stage_off;
                stage_off += (size_t)n * (size_t)bn;
                const int r2 = transpose(h, st, src, ld_in, stg, n, n, bn);
                if (r2) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)r2));
                CUDA_TRY(cudaMemcpyAsync(user + (size_t)b0 * (size_t)n, stg, sizeof(double) * (size_t)n * (size_t)bn, cudaMemcpyDeviceToHost, st));
            }
            return CB200_OK;
        };
        for (const Item &it : items) {
            if (!it.on || !it.user) continue;
            rc = send_back(d.*(it.slot), B, it.n, it.user);
            if (rc) return rc;
        }
        if (local_dk_from_window) {
            rc = send_back(gather_local + b0, pl.B_total, ndef, out->defects_kind);
            if (rc) return rc;
        }
    }
    for (int k = 1; k < ns_used; k++) {   // join
        CUDA_TRY(cudaEventRecord(h->ev_join[k - 1], streams[k]));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_join[k - 1], 0));
    }
    if (nchunks > 1 && (tail.nseg > 0 || tail.barrier)) {
        tail_kernel<<<1, 256, 0, h->stream>>>(tail);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    if (collective) {
        rc = finish_comm(h->stream, dk_compact, true);
        if (rc) return rc;
    }
    if (pl.osum) CUDA_TRY(cudaMemcpyAsync(out->objective_sum, dst_o, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (pl.gsum)
        CUDA_TRY(cudaMemcpyAsync(out->grad_sum, dst_g, sizeof(double) * (size_t)nvars, cudaMemcpyDeviceToHost, h->stream));
    if (pl.fsum) CUDA_TRY(cudaMemcpyAsync(out->fim_sum, dst_f, sizeof(double) * n2, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->acc_dirty = false;
    return comm_check(h);
}

extern "C" int cb200_sample_times(const cb200_handle *h, double *t)
{
    if (!h || !t) return fail(CB200_ERR_ARG, "null argument");
    for (size_t k = 0; k < h->h_sample_t.size(); k++) t[k] = h->h_sample_t[k];
    return CB200_OK;
}

extern "C" int cb200_resident(const cb200_handle *h, uint32_t want_bit, double **ptr, int64_t *ld)
{
    if (!h || !ptr) return fail(CB200_ERR_ARG, "null argument");
    if (!(h->resident_valid & want_bit)) return fail(CB200_ERR_ARG, "no device-resident result for want bit 0x%x from the last evaluation", want_bit);
    const DevBuf *b = nullptr;
    switch (want_bit) {
    case CB200_WANT_XEND: b = &h->ws_xend; break;
    case CB200_WANT_SENS: b = &h->ws_sens; break;
    case CB200_WANT_TRAJ: b = &h->ws_traj; break;
    case CB200_WANT_DEFECTS: b = &h->ws_dk; break;
    case CB200_WANT_OBJ: b = &h->ws_obj; break;
    case CB200_WANT_GRAD: b = &h->ws_grad; break;
    case CB200_WANT_FIM: b = &h->ws_fim; break;
    case CB200_WANT_JAC: b = &h->ws_jac; break;
    case CB200_WANT_CRIT: b = &h->ws_crit; break;
    case CB200_WANT_GRAD_COMPACT: b = &h->ws_gradc; break;
    case CB200_WANT_TRAJ_SENS: b = &h->ws_tsens; break;
    default: return fail(CB200_ERR_ARG, "want_bit must be exactly one CB200_WANT_* bit");
    }
    *ptr = (double *)b->p;
    if (ld) *ld = h->last_B;
    return CB200_OK;
}

extern "C" int cb200_abi_layout(int32_t which, int64_t *vals, int32_t n)
{
    if (!vals || n <= 0) return 0;
    std::vector<int64_t> v;
    switch (which) {
    case 0:
        v = {(int64_t)sizeof(cb200_desc), (int64_t)offsetof(cb200_desc, u0), (int64_t)offsetof(cb200_desc, model_data_len),
             (int64_t)offsetof(cb200_desc, observation_grid), (int64_t)offsetof(cb200_desc, reltol),
             (int64_t)offsetof(cb200_desc, saveat), (int64_t)offsetof(cb200_desc, n_saveat)};
        break;
    case 1:
        v = {(int64_t)sizeof(cb200_out), (int64_t)offsetof(cb200_out, objective_row), (int64_t)offsetof(cb200_out, jac_vals),
             (int64_t)offsetof(cb200_out, status), (int64_t)offsetof(cb200_out, defects_gathered), (int64_t)offsetof(cb200_out, comm_b0),
             (int64_t)offsetof(cb200_out, resident)};
        break;
    case 2:
        v = {(int64_t)sizeof(cb200_sizes), (int64_t)offsetof(cb200_sizes, n_traj_sens)};
        break;
    case 3:
        v = {(int64_t)sizeof(cb200_comm_info), (int64_t)offsetof(cb200_comm_info, epoch), (int64_t)offsetof(cb200_comm_info, error)};
        break;
    default: return 0;
    }
    const int m = std::min<int>(n, (int)v.size());
    for (int k = 0; k < m; k++) vals[k] = v[(size_t)k];
    return m;
}

// forward_initialization(rng, shooting; ps, fixed_indices) (src/node_initialization.jl:152-170) for B candidates:
// ps is updated in place (only the u0 entries of the intervals change).
extern "C" int cb200_forward_init(cb200_handle *h, double *ps, int64_t B, int64_t ldb, uint32_t flags,
                                  const int32_t *fixed_indices, int32_t n_fixed)
{
    NvtxRange nvtx_range("cb200_forward_init");
    if (!h || (!ps && h->L.nvars > 0)) return fail(CB200_ERR_ARG, "null argument");
    if (B < 0 || n_fixed < 0 || (n_fixed > 0 && !fixed_indices)) return fail(CB200_ERR_ARG, "bad argument");
    if (B == 0 || h->L.nvars == 0) return CB200_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    const int I = h->L.I, nvars = h->L.nvars;
    const bool dev = flags & CB200_MEM_DEVICE, soa = flags & CB200_PS_SOA;
    if (soa ? (ldb < B) : (ldb < nvars)) return fail(CB200_ERR_ARG, "ldb too small");
    std::vector<unsigned char> fx((size_t)I, 0);
    for (int k = 0; k < n_fixed; k++) {
        if (fixed_indices[k] < 1 || fixed_indices[k] > I) return fail(CB200_ERR_ARG, "fixed_indices entry out of 1..I");
        fx[fixed_indices[k] - 1] = 1;
    }
    if (h->ws_fixed.reserve((size_t)I)) return fail(CB200_ERR_NOMEM, "device out of memory");
    CUDA_TRY(cudaMemcpyAsync(h->ws_fixed.p, fx.data(), (size_t)I, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // fx is a stack-lifetime buffer
    double *d_ps = ps;
    long long d_ldb = ldb;
    const size_t bytes = sizeof(double) * (size_t)nvars * (size_t)B;
    if (!(dev && soa)) {
        if (h->ws_ps.reserve(bytes)) return fail(CB200_ERR_NOMEM, "device out of memory (ps)");
        d_ps = (double *)h->ws_ps.p;
        d_ldb = B;
        if (soa) {  // host, variable-major
            CUDA_TRY(cudaMemcpy2DAsync(d_ps, sizeof(double) * B, ps, sizeof(double) * ldb, sizeof(double) * B, nvars,
                                       cudaMemcpyHostToDevice, h->stream));
        } else {
            const double *cm = ps;
            long long ldc = ldb;
            if (!dev) {  // host, candidate-major
                if (h->ws_in.reserve(bytes)) return fail(CB200_ERR_NOMEM, "device out of memory (staging)");
                CUDA_TRY(cudaMemcpy2DAsync(h->ws_in.p, sizeof(double) * nvars, ps, sizeof(double) * ldb,
                                           sizeof(double) * nvars, B, cudaMemcpyHostToDevice, h->stream));
                cm = (const double *)h->ws_in.p;
                ldc = nvars;
            }
            int rc = transpose(h, h->stream, cm, ldc, d_ps, B, B, nvars);
            if (rc) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        }
    }
    int rc = forward_model(h->model, h->method, h->L, d_ps, d_ldb, B, (const unsigned char *)h->ws_fixed.p, h->stream);
    if (rc == -1000) return fail(CB200_ERR_UNSUPPORTED, "model %d has no forward-sweep kernel", h->model);
    if (rc) return fail(CB200_ERR_CUDA, "forward kernel launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    h->launches++;
    if (!(dev && soa)) {
        if (soa) {
            CUDA_TRY(cudaMemcpy2DAsync(ps, sizeof(double) * ldb, d_ps, sizeof(double) * B, sizeof(double) * B, nvars,
                                       cudaMemcpyDeviceToHost, h->stream));
        } else if (dev) {
            rc = transpose(h, h->stream, d_ps, B, ps, ldb, nvars, B);
            if (rc) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        } else {
            rc = transpose(h, h->stream, d_ps, B, (double *)h->ws_in.p, nvars, nvars, B);
            if (rc) return fail(CB200_ERR_CUDA, "transpose launch failed: %s", cudaGetErrorString((cudaError_t)rc));
            CUDA_TRY(cudaMemcpy2DAsync(ps, sizeof(double) * ldb, h->ws_in.p, sizeof(double) * nvars, sizeof(double) * nvars,
                                       B, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    if (!dev) CUDA_TRY(cudaStreamSynchronize(h->stream));
    return CB200_OK;
}

static double probe_fp64(int device, int iters, bool mma)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * (mma ? 2 : 8), threads = 256;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto launch = [&](int n, double seed) {
        if (mma)
            dmma_probe_kernel<<<blocks, threads>>>(out, n, seed);
        else
            dfma_probe_kernel<<<blocks, threads>>>(out, n, seed);
    };
    launch(iters / 4 + 1, 1.0);  // warm-up
    double best = -1.0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        launch(iters, 1.0 + rep);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        // per thread and iteration: 8 DFMA, or 4 DMMA of 8*8*4 FMA shared by the 32 lanes of a warp
        const double fma_per_thread_iter = mma ? 4.0 * 256.0 / 32.0 : 8.0;
        const double flops = 2.0 * fma_per_thread_iter * (double)iters * (double)blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}

extern "C" double cb200_probe_fp64_tflops(int device, int iters) { return probe_fp64(device, iters, false); }
extern "C" double cb200_probe_fp64_mma_tflops(int device, int iters) { return probe_fp64(device, iters, true); }
