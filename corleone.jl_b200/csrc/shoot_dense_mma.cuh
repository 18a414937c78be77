// K2-dense: shooting units of a dense 20-state model  x' = W x + phi(x) + b u  with the FULL sensitivity
// Jacobian, on the FP64 tensor path (mma.sync.m8n8k4.f64).
//
// Here the sensitivity right-hand side is a real dense contraction:  d/dt [x S] = W [x S] + (elementwise terms),
// the same 20x20 matrix W applied to all 1 + ns columns of every unit at every stage (north_star: "tensor cores
// only if the Jacobian-matrix products become a real dense contraction at large state dimension").  On B200 DMMA
// and DFMA share one FP64 datapath (tools/dmma_micro.cu), but DMMA reaches the nominal 64 FMA/clk/SM with 1/8 of
// the issue slots and keeps W in 15 fragment registers per lane instead of constant-bank operands.  (The wider PTX
// shapes m16n8k4 / k8 / k16 .f64 lower to the same DMMA.8x8x4 SASS on sm_100a: nothing to gain there.)
//
// A tile = 8 columns of Y = [x | S_1..S_ns] = the A operand (rows = columns of Y, k = state) of
// D[col][i] = sum_j Y[j][col] W[i][j]; W^T is the B operand.  Lane (g = lane/4, t = lane%4) owns the five states
// 5t..5t+4 of column g in every vector of the integrator: as A fragment it supplies state 5t+kt in k-tile kt, and the
// state order of the three n-tiles is permuted (n-slot 2t+e of tile nt <-> state 5t+2nt+e, the sixth slot is zero
// padding) so that the D fragments it receives are exactly its own five states.  No shuffle, no shared-memory
// transpose: the stage derivative lands in the registers of the lane that needs it.  The only exchange is x -> S tiles
// (the -3 x^2 S term), through shared memory.
//
// Two kernels:
//   * shoot_dense_mma_queue_kernel (the throughput path): persistent blocks, an ITEM = 8 consecutive candidates of one
//     interval = one x tile (8 state columns: every MMA row is live) + ns S tiles (tile tau = direction tau of the 8
//     candidates: direction, piece forcing and table look-ups are warp-uniform, stores touch 8 consecutive candidates
//     per row).  The warps of a block pull tile tasks from a queue in shared memory; the x task of item K+2 is queued
//     among the S tasks of item K and leaves -3 x^2 of all stages in one of three ring slots, so the S tiles never wait
//     for the states, no warp is tied to a role, and the hardware balances the four sub-partitions by itself.
//   * shoot_dense_mma_kernel (lockstep; layers whose stage count does not fit the ring): a block owns 5 candidates of
//     one interval, warp 0 = x tile, warps 1..15 = 120 sensitivity columns, one __syncthreads per stage.
//
// Reference semantics as in shoot.cuh (pieces, FSAL restart per piece, accumulator-form Tsit5, fused K3/K5 epilogue).
#pragma once
#include "common.cuh"
#include "tableau.cuh"
#include <algorithm>
#include <cstdlib>

namespace cb200 {

// NX (template parameter): states, a multiple of 4: 16, 20, 24, 32.  Lane (g, t) owns the S = NX/4 states S t .. S t + S - 1;
// k-tiles KT = S (k-tile kt = states {S t + kt}), n-tiles NT = ceil(S / 2) (n-slot 2 tau + e of tile nt <-> state
// S tau + 2 nt + e; for odd S the last slot is zero padding).
template <int NX>
struct DenseDims {
    static_assert(NX % 4 == 0 && NX >= 8 && NX <= 32, "dense tensor path: NX must be a multiple of 4 in 8..32");
    static constexpr int S = NX / 4, KT = S, NT = (S + 1) / 2;
};
constexpr int DM_CPB = 5;        // candidates per block
constexpr int DM_SW = 15;        // S-tile warps: 15 * 8 = 24 * 5 columns per round
constexpr int DM_DPR = 24;       // directions per round
constexpr int DM_WARPS = DM_SW + 1;
constexpr int DM_THREADS = DM_WARPS * 32;

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NX>
struct DenseTile {
    double Wf[DenseDims<NX>::KT * DenseDims<NX>::NT];   // B fragments of W^T: [kt*NT + nt]
    bool col_on;     // this lane's column is live
    int t;           // lane % 4
};

// the lane's W^T fragments in the permuted state order (see header); loaded once per kernel
template <int NX>
__device__ __forceinline__ void dense_load_w(const DevLayer &L, DenseTile<NX> &T)
{
    constexpr int S = DenseDims<NX>::S, KT = DenseDims<NX>::KT, NT = DenseDims<NX>::NT;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double *W = L.model_data;   // row-major NX x NX, then b[NX]
    T.t = t;
#pragma unroll
    for (int kt = 0; kt < KT; kt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            // B[k = t][n = g]: k-state S t + kt; n-slot g = 2 tau + e -> state S tau + 2 nt + e (beyond S: padding)
            const int tau = g >> 1, e = g & 1;
            const int sn = S * tau + 2 * nt + e;
            T.Wf[kt * NT + nt] = (2 * nt + e >= S) ? 0.0 : __ldg(W + sn * NX + (S * t + kt));
        }
}

// k = W y + elementwise(y; x) for the lane's five states.  `slot` = this lane's five entries of the exchange slot of this
// right-hand-side call ([CPB][20] per call, already offset by candidate and state group).
// The x tile publishes c = -3 x^2 (what the S tiles need).  PIPE = false: x and S tiles of the same unit run in lockstep, one
// __syncthreads per call (every warp of the block calls this the same number of times).  PIPE = true: the x tile ran
// ahead, `slot` already holds its value.  No predicate on idle tiles: a tile without live columns integrates zeros
// (lockstep) or never gets here (pipelined), so the stage body is straight-line.
template <int NX, bool PIPE, bool IS_X>
__device__ __forceinline__ void dense_rhs(const DenseTile<NX> &T, const double *y, const double *binit, double *k, double *slot)
{
    constexpr int S = DenseDims<NX>::S, KT = DenseDims<NX>::KT, NT = DenseDims<NX>::NT;
    double q[S];
    if constexpr (IS_X) {
#pragma unroll
        for (int s = 0; s < S; s++) q[s] = y[s] * y[s];
        if (T.col_on) {
#pragma unroll
            for (int s = 0; s < S; s++) slot[s] = -3.0 * q[s];
        }
    }
    // the accumulators start from the b-term of the piece: b u for the states, b for the running piece's control
    // direction, 0 otherwise (the C operand of the first DMMA of each chain)
    double d[2 * NT];
#pragma unroll
    for (int s = 0; s < 2 * NT; s++) d[s] = s < S ? binit[s < S ? s : 0] : 0.0;
#pragma unroll
    for (int kt = 0; kt < KT; kt++) {
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma884(d[2 * nt], d[2 * nt + 1], y[kt], T.Wf[kt * NT + nt]);
    }
    if constexpr (!PIPE) __syncthreads();
    if constexpr (IS_X) {  // f(x) = W x - x^3 + b u
#pragma unroll
        for (int s = 0; s < S; s++) k[s] = fma(-q[s], y[s], d[s]);
    } else {               // f_x v + f_u = W v - 3 x^2 v (+ b)
#pragma unroll
        for (int s = 0; s < S; s++) k[s] = fma(slot[s], y[s], d[s]);
    }
}

// slot of the next right-hand-side call: two alternating slots (lockstep) or one slot per call of the item (ring);
// sp walks the lane's entries of the slots, sp_alt is the lockstep partner
#define DM_RHS(y)                                      \
    do {                                               \
        dense_rhs<NX, PIPE, IS_X>(T, y, binit, k, sp);     \
        if constexpr (PIPE) {                          \
            sp += slot_stride;                         \
        } else {                                       \
            double *tmp_ = sp;                         \
            sp = sp_alt;                               \
            sp_alt = tmp_;                             \
        }                                              \
    } while (0)

// One column per lane group of the calling warp's tile, across the pieces of interval i: the state x of candidate b
// (IS_X) or its sensitivity with respect to local variable `dir`.  col_on: the column exists (else it integrates
// zeros); store: it writes results (a live candidate); x_results: the x tile writes the unit's results (first
// direction round only).  sp / sp_alt / slot_stride: the lane's entries of the stage-exchange slots (DM_RHS).
template <int NX, int METHOD, bool UNI, bool PIPE, bool IS_X>
__device__ __forceinline__ void dense_column(const DevLayer &L, const double *__restrict__ ps, long long ldb,
                                             const ShootOut &o, int i, int nu0, int ns, int dir, long long b, bool col_on,
                                             bool store, bool x_results, double *sp, double *sp_alt, int slot_stride,
                                             DenseTile<NX> &T)
{
    constexpr int S = DenseDims<NX>::S, NCF = METHOD == 0 ? 21 : 4;
    const int t = threadIdx.x & 3;
    (void)sp_alt;
    (void)slot_stride;
    T.col_on = col_on;
    const int voff = __ldg(L.var_off + i);
    const double *__restrict__ v = ps + (long long)voff * ldb + b;
    // ---- initial values of the lane's five entries
    double yd[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int k = S * t + s;
        const int src = __ldg(L.ic_src + i * NX + k);
        if (IS_X)
            yd[s] = (src >= 0) ? v[(long long)src * ldb] : __ldg(L.ic_fix + i * NX + k);
        else
            yd[s] = (T.col_on && src >= 0 && src == dir) ? 1.0 : 0.0;
    }
    if (!T.col_on) {
#pragma unroll
        for (int s = 0; s < S; s++) yd[s] = 0.0;
    }
    const int cdir = dir - nu0 - L.ntp;   // local control index this column differentiates with respect to (< 0: a u0 direction)
    const int M = __ldg(L.M + i), po = __ldg(L.piece_off + i), so = __ldg(L.samp_off + i);
    const int nrow = NX + L.nact;
    const bool last_interval = (i == L.I - 1);
    const double *__restrict__ vc = v + (long long)(nu0 + L.ntp) * ldb;

    int ci_nx = __ldg(L.cidx + (long long)po * L.nact);
    double u_nx = vc[(long long)ci_nx * ldb];
    for (int j = 0; j < M; j++) {
        const int ci = ci_nx;
        const double u = u_nx;
        if (j + 1 < M) {
            ci_nx = __ldg(L.cidx + (long long)(po + j + 1) * L.nact);
            u_nx = vc[(long long)ci_nx * ldb];
        }
        double binit[S];   // b-term of this piece for the lane's states
        {
            const double w = IS_X ? u : ((T.col_on && cdir == ci) ? 1.0 : 0.0);
#pragma unroll
            for (int s = 0; s < S; s++) binit[s] = w * __ldg(L.model_data + NX * NX + S * t + s);
        }
        if (!IS_X && o.tsens && store && j == 0) {   // sensitivities at the interval's start sample
            const long long row0 = __ldg(L.tsens_off + i) + (long long)dir * NX + S * t;
#pragma unroll
            for (int s = 0; s < S; s++) o.tsens[(row0 + s) * o.ldo + b] = yd[s];
        }
        if (o.traj && IS_X && store && x_results && j == 0) {  // save_start = (j == 1)
            double *tr = o.traj + ((long long)so * nrow) * o.ldo + b;
#pragma unroll
            for (int s = 0; s < S; s++) tr[(long long)(S * t + s) * o.ldo] = yd[s];
            if (t == 0) tr[(long long)NX * o.ldo] = u;
        }
        double cf[NCF];
        if constexpr (UNI) {
#pragma unroll
            for (int c = 0; c < NCF; c++) cf[c] = L.hA[c];
        } else {
            const double h = __ldg(L.ph + po + j);
            if constexpr (METHOD == 0)
                tsit5_coeffs(h, cf);
            else
                rk4_coeffs(h, cf);
        }
        double k[S];
#pragma unroll
        for (int c = 0; c < S; c++) k[c] = 0.0;
        if constexpr (METHOD == 0) {   // accumulator-form Tsit5 (shoot.cuh), five components per lane
            double s2[S], s3[S], s4[S], s5[S], s6[S];
            DM_RHS(yd);
            for (int st = 0; st < L.K; st++) {
#pragma unroll
                for (int c = 0; c < S; c++) {
                    s2[c] = fma(cf[0], k[c], yd[c]);
                    s3[c] = fma(cf[1], k[c], yd[c]);
                    s4[c] = fma(cf[3], k[c], yd[c]);
                    s5[c] = fma(cf[6], k[c], yd[c]);
                    s6[c] = fma(cf[10], k[c], yd[c]);
                    yd[c] = fma(cf[15], k[c], yd[c]);
                }
                DM_RHS(s2);
#pragma unroll
                for (int c = 0; c < S; c++) {
                    s3[c] = fma(cf[2], k[c], s3[c]);
                    s4[c] = fma(cf[4], k[c], s4[c]);
                    s5[c] = fma(cf[7], k[c], s5[c]);
                    s6[c] = fma(cf[11], k[c], s6[c]);
                    yd[c] = fma(cf[16], k[c], yd[c]);
                }
                DM_RHS(s3);
#pragma unroll
                for (int c = 0; c < S; c++) {
                    s4[c] = fma(cf[5], k[c], s4[c]);
                    s5[c] = fma(cf[8], k[c], s5[c]);
                    s6[c] = fma(cf[12], k[c], s6[c]);
                    yd[c] = fma(cf[17], k[c], yd[c]);
                }
                DM_RHS(s4);
#pragma unroll
                for (int c = 0; c < S; c++) {
                    s5[c] = fma(cf[9], k[c], s5[c]);
                    s6[c] = fma(cf[13], k[c], s6[c]);
                    yd[c] = fma(cf[18], k[c], yd[c]);
                }
                DM_RHS(s5);
#pragma unroll
                for (int c = 0; c < S; c++) {
                    s6[c] = fma(cf[14], k[c], s6[c]);
                    yd[c] = fma(cf[19], k[c], yd[c]);
                }
                DM_RHS(s6);
#pragma unroll
                for (int c = 0; c < S; c++) yd[c] = fma(cf[20], k[c], yd[c]);
                if (st + 1 < L.K) DM_RHS(yd);   // FSAL
            }
        } else {   // classical RK4
            double acc[S], td[S];
            for (int st = 0; st < L.K; st++) {
                DM_RHS(yd);
#pragma unroll
                for (int c = 0; c < S; c++) {
                    acc[c] = fma(cf[2], k[c], yd[c]);
                    td[c] = fma(cf[0], k[c], yd[c]);
                }
                DM_RHS(td);
#pragma unroll
                for (int c = 0; c < S; c++) {
                    acc[c] = fma(cf[3], k[c], acc[c]);
                    td[c] = fma(cf[0], k[c], yd[c]);
                }
                DM_RHS(td);
#pragma unroll
                for (int c = 0; c < S; c++) {
                    acc[c] = fma(cf[3], k[c], acc[c]);
                    td[c] = fma(cf[1], k[c], yd[c]);
                }
                DM_RHS(td);
#pragma unroll
                for (int c = 0; c < S; c++) yd[c] = fma(cf[2], k[c], acc[c]);
            }
        }
        // save_end; the end sample of a non-final interval is dropped by the merge (multiple_shooting.jl:178-180)
        if (!IS_X && o.tsens && store && (j + 1 < M || last_interval)) {
            const long long row0 = __ldg(L.tsens_off + i) + ((long long)(j + 1) * ns + dir) * NX + S * t;
#pragma unroll
            for (int s = 0; s < S; s++) o.tsens[(row0 + s) * o.ldo + b] = yd[s];
        }
        if (o.traj && IS_X && store && x_results && (j + 1 < M || last_interval)) {
            double *tr = o.traj + ((long long)(so + j + 1) * nrow) * o.ldo + b;
#pragma unroll
            for (int s = 0; s < S; s++) tr[(long long)(S * t + s) * o.ldo] = yd[s];
            if (t == 0) tr[(long long)NX * o.ldo] = u;
        }
    }

    // ---- results (fused K3 / K5 epilogue, as in shoot.cuh)
    if (!store) return;
    if constexpr (IS_X) {
        if (!x_results) return;
        if (o.status) {
            bool ok = true;
#pragma unroll
            for (int s = 0; s < S; s++) ok = ok && isfinite(yd[s]);
            if (!ok) atomicOr(o.status + b, 1);
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int k = S * t + s;
            if (o.xend) o.xend[((long long)i * NX + k) * o.ldo + b] = yd[s];
            if (o.objq && k == o.obj_state) o.objq[(long long)i * o.ldo + b] = yd[s];
        }
        if ((o.dk || o.dst) && i + 1 < L.I) {
            const int vn = __ldg(L.var_off + i + 1);
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int k = S * t + s;
                const int pk = __ldg(L.dpos_kind + i * NX + k);
                if (pk >= 0) {
                    const int src = __ldg(L.ic_src + (i + 1) * NX + k);
                    const double nxt = (src >= 0) ? ps[(long long)(vn + src) * ldb + b] : __ldg(L.ic_fix + (i + 1) * NX + k);
                    const double dd = yd[s] - nxt;
                    if (o.dk) store_defect(o, pk, b, dd);
                    if (o.dst) o.dst[(long long)__ldg(L.dpos_stage + i * NX + k) * o.ldo + b] = dd;
                }
            }
            if (t == 0) {  // parameter and control matchings: one lane per candidate
                const int e1 = __ldg(L.oth_off + i + 1);
                for (int e = __ldg(L.oth_off + i); e < e1; e++) {
                    const MatchEntry m = L.oth[e];
                    const double dd = ps[(long long)m.a * ldb + b] - ps[(long long)m.b * ldb + b];
                    if (o.dk) store_defect(o, m.pos_kind, b, dd);
                    if (o.dst) o.dst[(long long)m.pos_stage * o.ldo + b] = dd;
                }
            }
        }
        return;
    }
    const long long s0 = __ldg(L.sens_off + i);
    const bool inc = o.obj_all || i == L.I - 1;
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int k = S * t + s;
        if (o.sens) o.sens[(s0 + (long long)dir * NX + k) * o.ldo + b] = yd[s];
        if (o.jac && i + 1 < L.I) {
            const int jp = __ldg(L.jpos + i * NX + k);
            if (jp >= 0) o.jac[(long long)(jp + dir) * o.ldo + b] = yd[s];
        }
        if (k == o.obj_state) {
            const double val = inc ? yd[s] : 0.0;
            if (o.grad) o.grad[(long long)(voff + dir) * o.ldo + b] = val;
            if (o.gradc && voff + dir >= o.gradc_first) o.gradc[(long long)(voff + dir - o.gradc_first) * o.ldo + b] = val;
            if (o.grad_sum) atomicAdd(o.grad_sum + voff + dir, val);
        }
    }
}

#undef DM_RHS

// Lockstep kernel: one block per unit group (5 candidates of one interval, direction round blockIdx.z), x and S tiles
// exchange through two alternating slots; both branches meet at the same barrier once per right-hand-side call.
template <int NX, int METHOD, bool UNI>
__global__ void __launch_bounds__(DM_THREADS, 1)
shoot_dense_mma_kernel(const __grid_constant__ DevLayer L, const double *__restrict__ ps, long long ldb, long long B,
                       const ShootOut o)
{
    __shared__ double sx[2 * DM_CPB * NX];
    DenseTile<NX> T;
    dense_load_w(L, T);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int i = blockIdx.y + L.i0, round = blockIdx.z;
    const long long b0 = (long long)blockIdx.x * DM_CPB;
    const int nu0 = __ldg(L.n_u0 + i), ns = nu0 + L.ntp + __ldg(L.n_c + i);
    if (round > 0 && round * DM_DPR >= ns) return;   // uniform over the block
    const bool is_x = warp == 0;
    const int col = 8 * (warp - 1) + g;                                 // S tiles: column among the 120 of this round
    const int cand = is_x ? (g < DM_CPB ? g : 0) : col % DM_CPB;
    const int dir = round * DM_DPR + col / DM_CPB;                      // S tiles: direction of this lane's column
    long long b = b0 + cand;
    const bool cand_live = b < B;
    if (!cand_live) b = B - 1;
    const bool col_on = is_x ? (g < DM_CPB) : (dir < ns);
    double *sp = sx + cand * NX + DenseDims<NX>::S * t;
    if (is_x)
        dense_column<NX, METHOD, UNI, false, true>(L, ps, ldb, o, i, nu0, ns, 0, b, col_on, col_on && cand_live, round == 0, sp,
                                               sp + DM_CPB * NX, 0, T);
    else
        dense_column<NX, METHOD, UNI, false, false>(L, ps, ldb, o, i, nu0, ns, dir, b, col_on, col_on && cand_live, false, sp,
                                                sp + DM_CPB * NX, 0, T);
}

// ---- queue kernel
constexpr int DQ_CPI = 8;   // candidates per item = rows of the MMA
constexpr int DQ_MAXRING = 3;

__device__ __forceinline__ unsigned dq_peek(const unsigned *p) { return *(const volatile unsigned *)p; }

// Persistent blocks; block `blockIdx.x` owns items blockIdx.x, blockIdx.x + gridDim.x, ... (item -> interval fastest).
// Item K of the block = one x task + n_tiles S tasks.  Task order in the queue: x_0 .. x_{la-1}, then for every item K
// its S tasks with x_{K+la} inserted after the px-th (la = nring - 1): when a warp picks x_{K+la}, the S tasks of item
// K - 1 -- the previous users of its ring slot -- were queued more than a block's worth of warps earlier and have all
// but certainly finished, and the result is needed only a whole item later.  Every wait is on a task EARLIER in the
// queue (already running or done), so the order is deadlock-free.
//   s_full[slot] = K + 1 once x_K has filled the slot;  s_done[slot] counts the finished S tasks on the slot.
template <int NX, int METHOD, bool UNI, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1)
shoot_dense_mma_queue_kernel(const __grid_constant__ DevLayer L, const double *__restrict__ ps, long long ldb, long long B,
                             const ShootOut o, long long n_items, int n_tiles, int nst, int nring)
{
    extern __shared__ double ring[];   // [nring][nst][8 candidates][20]
    __shared__ unsigned q_head, s_full[DQ_MAXRING], s_done[DQ_MAXRING];
    if (threadIdx.x == 0) {
        q_head = 0;
        for (int r = 0; r < DQ_MAXRING; r++) s_full[r] = s_done[r] = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = nst * DQ_CPI * NX;
    const unsigned nk = n_items > blockIdx.x ? (unsigned)((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    const unsigned la = (unsigned)nring - 1u, row = (unsigned)n_tiles + 1u;
    // (nring == 1, layers with very many stages: no look-ahead, the x task of an item opens its row)
    const unsigned px = nring == 1 ? 0u : (unsigned)min(nring == 3 ? 16 : 8, n_tiles);
    const unsigned n_tasks = la + nk * row;
    DenseTile<NX> T;
    dense_load_w(L, T);
    for (;;) {
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(&q_head, 1u);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= n_tasks) break;
        unsigned K;
        int tau = -1;   // -1: x task
        if (tk < la) {
            K = tk;
        } else {
            const unsigned q = tk - la;
            K = q / row;
            const unsigned r = q - K * row;
            if (r == px)
                K += la;
            else
                tau = (int)(r < px ? r : r - 1);
        }
        if (K >= nk) continue;   // x tasks past the last item
        const long long n = blockIdx.x + (long long)K * gridDim.x;
        const int i = (int)(n % L.In) + L.i0;
        long long b = (n / L.In) * DQ_CPI + g;
        const bool cand_live = b < B;
        if (!cand_live) b = B - 1;
        const int slot = (int)(K % (unsigned)nring);
        double *sp = ring + (long long)slot * per + g * NX + DenseDims<NX>::S * t;
        const int nu0 = __ldg(L.n_u0 + i), ns = nu0 + L.ntp + __ldg(L.n_c + i);
        if (tau < 0) {
            const unsigned need = (unsigned)n_tiles * (K / (unsigned)nring);   // S tasks of the slot's earlier items
            while (dq_peek(&s_done[slot]) < need) __nanosleep(64);
            __threadfence_block();
            dense_column<NX, METHOD, UNI, true, true>(L, ps, ldb, o, i, nu0, ns, 0, b, true, cand_live, true, sp, nullptr,
                                                  DQ_CPI * NX, T);
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                *(volatile unsigned *)&s_full[slot] = K + 1u;
            }
        } else {
            // (a tile beyond the interval's directions has nothing to integrate, but it reports `done` only after the item's
            // x task like everybody else: the x task that re-uses the slot counts reports, and a report of THIS item must
            // not stand in for a missing one of the slot's previous item)
            while (dq_peek(&s_full[slot]) < K + 1u) __nanosleep(64);
            __threadfence_block();
            if (tau < ns)
                dense_column<NX, METHOD, UNI, true, false>(L, ps, ldb, o, i, nu0, ns, tau, b, true, cand_live, false, sp, nullptr,
                                                       DQ_CPI * NX, T);
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&s_done[slot], 1u);
            }
        }
    }
}

// returns cudaError_t as int; -1000 if the layer does not fit this kernel.  WARPS: warps per block of the queue kernel
// (the register budget per lane grows with NX: 16 warps x 128 registers up to NX = 20, 12 x 168 for 24, 8 x 255 for 32)
template <int NX, int WARPS>
inline int launch_dense_mma(int method, const DevLayer &L, const double *ps, long long ldb, long long B, const ShootOut &o,
                            int max_ns, int max_pieces, cudaStream_t s)
{
    if (L.nact != 1 || L.ntp != 0 || L.nx != NX || !L.model_data) return -1000;
    if (B <= 0) return 0;
    if (max_ns < 1) max_ns = 1;
    // right-hand-side calls per interval: Tsit5 6K per piece (FSAL after the last step skipped), RK4 4K
    const int nst = max_pieces * (method == 0 ? 6 : 4) * L.K;
    const size_t per = sizeof(double) * (size_t)nst * DQ_CPI * NX;
    static const bool lockstep = getenv("CB200_DENSE20_LOCKSTEP") != nullptr;
    static const int nring_env = getenv("CB200_DENSE20_NRING") ? atoi(getenv("CB200_DENSE20_NRING")) : 0;
    const size_t smem_cap = 225 * 1024;
    int nring = (nring_env >= 1 && nring_env <= 3) ? nring_env : 3;
    while (nring > 1 && nring * per > smem_cap) nring--;
    if (!lockstep && nring * per <= smem_cap) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long n_items = ((B + DQ_CPI - 1) / DQ_CPI) * L.In;
        const unsigned grid = (unsigned)std::min<long long>(n_items, std::max(1, sms - std::max(0, L.sm_reserve)));
        const size_t smem = nring * per;
        auto go = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, WARPS * 32, smem, s>>>(L, ps, ldb, B, o, n_items, max_ns, nst, nring);
        };
        if (method == 0) {
            if (L.uniform_h) go(shoot_dense_mma_queue_kernel<NX, 0, true, WARPS>);
            else go(shoot_dense_mma_queue_kernel<NX, 0, false, WARPS>);
        } else {
            if (L.uniform_h) go(shoot_dense_mma_queue_kernel<NX, 1, true, WARPS>);
            else go(shoot_dense_mma_queue_kernel<NX, 1, false, WARPS>);
        }
        return (int)cudaGetLastError();
    }
    {   // lockstep fallback (stage count beyond even one ring slot): 16 warps x 128 registers -- spills beyond nx = 20
        const long long gx = (B + DM_CPB - 1) / DM_CPB;
        const int rounds = (max_ns + DM_DPR - 1) / DM_DPR;
        if (gx > 2147483647LL || L.In > 65535 || rounds > 65535) return -1000;
        dim3 grid((unsigned)gx, (unsigned)L.In, (unsigned)rounds);
        if (method == 0) {
            if (L.uniform_h)
                shoot_dense_mma_kernel<NX, 0, true><<<grid, DM_THREADS, 0, s>>>(L, ps, ldb, B, o);
            else
                shoot_dense_mma_kernel<NX, 0, false><<<grid, DM_THREADS, 0, s>>>(L, ps, ldb, B, o);
        } else {
            if (L.uniform_h)
                shoot_dense_mma_kernel<NX, 1, true><<<grid, DM_THREADS, 0, s>>>(L, ps, ldb, B, o);
            else
                shoot_dense_mma_kernel<NX, 1, false><<<grid, DM_THREADS, 0, s>>>(L, ps, ldb, B, o);
        }
        return (int)cudaGetLastError();
    }
}

}  // namespace cb200
