// Shared device-side layer description and launch interfaces.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cb200 {

constexpr int MAX_NP = 16;
constexpr int MAX_CTRL = 8;

// Static tables of one layer, resident in HBM (tiny; L1/L2 cached).  Built once in cb200_create from the
// reference's `st` (src/single_shooting.jl:343-351), converted to 0-based.
struct DevLayer {
    int nx, np_all, nact, ntp, I, K;
    int R;            // direction groups per unit = max_i ceil(ns_i / G); 1 when G == 0
    int nvars;
    int max_M;               // largest number of pieces of an interval
    int flat;                // launch mapping: 1 = flat thread index (tiny batches), 0 = grid (b-tiles, i, r)
    int i0, In;              // intervals [i0, i0 + In) are integrated by this launch (cb200_set_interval_range); default 0, I
    int sm_reserve;          // persistent kernels leave this many SMs free (host pipeline: the layout conversions of the
                             // neighbouring chunks need somewhere to run while a chunk's kernel holds every register)
    int uniform_h;           // every piece of the layer has the same step size h = (t1 - t0)/K
    double hA[21];           // uniform_h: h * a_ij (Tsit5, slots as tsit5_coeffs) or the 4 RK4 factors
    double abstol, reltol;   // method 2 (adaptive Tsit5): the problem's tolerances
    int pmap_kind[MAX_NP];   // p_full[q] = kind ? controls_j[idx] : ps.p[idx]   (A, B of single_shooting.jl:306-323)
    int pmap_idx[MAX_NP];
    int tunable_p0[MAX_NP];  // 0-based p index of tunable parameter m
    int row_to_p[MAX_CTRL];  // 0-based p index fed by control row r
    const int *M;            // [I] pieces
    const int *piece_off;    // [I] offset into the piece arrays
    const double *pt0;       // [sum M] piece start
    const double *pt1;       // [sum M] piece end
    const double *ph;        // [sum M] step size of the piece, (t1 - t0)/K
    const int *cidx;         // [sum M * nact] 0-based local control index of (piece, row)
    const int *var_off;      // [I] offset of the interval's block in the flat ps
    const int *n_u0;         // [I]
    const int *n_c;          // [I]
    const int *ic_src;       // [I*nx] >= 0: index into the interval's ps.u0; -1: fixed value
    const double *ic_fix;    // [I*nx]
    const double *ic_fix0;   // [I*nx] the same with fixval = problem.u0 on every interval (forward sweep)
    const long long *sens_off;  // [I] first row of the interval's block in `sens`
    const int *samp_off;     // [I] first merged-trajectory sample of the interval
    const long long *tsens_off;  // [I] first row of the interval's block in `traj_sens`
    const int *psamp;        // [sum M] index of the piece's END sample among the samples of its interval (j + 1 without saveat)
    const int *nsm;          // [I] samples of the interval in the merged trajectory (its end sample dropped unless it is the last)
    const int *sv_n;         // [sum M] saveat times strictly inside the piece -- nullptr: the layer has none anywhere
    const int *sv_off;       // [sum M] offset of the piece's times in sv_t
    const double *sv_t;      // the saveat times, piece after piece
    const double *model_data;  // model constants in HBM (DENSE20: W row-major 20x20, then b[20]) or nullptr
    const signed char *dir_pcol;  // [nvars] 0-based p column that local variable (= direction) perturbs; -1: a u0 entry
    const int *dir_perm;          // [nvars] direction held by slot s of the interval (activation order, see shoot.cuh)
    const unsigned *act_mask;     // [sum M][R] bits 0..23: slot r*G+g is forced during the piece (G = handle's group size);
                                  //            bits 24..31: live prefix length of the group
    // shooting defects of the boundary between interval i and i+1 (src/multiple_shooting.jl:150-163)
    const int *dpos_kind;    // [I*nx] position of the state matching of state k in shooting_constraints, -1: none
    const int *dpos_stage;   // [I*nx] same, in stage_ordered_shooting_constraints
    const int *jpos;         // [I*nx] first cons_j value slot of (boundary i, state k): its ns_i entries follow in
                             //        direction order (cb200_jac_structure order), -1: state not matched
    const int *oth_off;      // [I+1] range of the boundary's parameter / control matchings in `oth`
    const struct MatchEntry *oth;
};

struct MatchEntry {  // defect = ps[a] - ps[b]  (parameter and control matchings)
    int a, b, pos_kind, pos_stage;
};

// ---- multi-GPU data plane (comm.cu): peer-mapped windows of the ranks of one box, one rank per GPU.
constexpr int MAX_RANKS = 16;
constexpr int FLAG_STRIDE = 16;   // one 128-byte line per flag
// Device-resident table of one handle's communicator.  Rank p's window is mapped into this process at slots[p] /
// flags[p] / gather[p] (CUDA IPC between processes, plain peer access between the threads of one process).
struct CommPeers {
    int nranks, rank;
    int s_cap;                                    // doubles per sums slot
    double *slots[MAX_RANKS];                     // [2][nranks][s_cap]: slot (parity, q) of rank p's window is written by rank q
    unsigned long long *flags[MAX_RANKS];         // [2][nranks][FLAG_STRIDE]: flag (parity, q) of rank p is set by rank q
    double *gather[MAX_RANKS];                    // [2][gather_cap]: all-gathered defects, [row][B_total] per parity
    long long gather_cap;                         // doubles per parity buffer
    unsigned long long *err;                      // host-mapped error word of THIS rank (0 = fine)
    long long timeout_ns;                         // a peer that does not show up within this time is reported, not waited for
};
struct TailSeg {
    int off, len;      // segment of the accumulator block
    double *dst;       // where the finished sums go (device pointer)
};
// Tail of the LAST kernel of an evaluation (tail.cuh): the last block to finish hands the cross-candidate sums to their
// destination -- all-reduced over the ranks by a one-shot exchange through the peer windows -- clears the accumulators
// and closes the epoch (signal to / wait for every peer: their all-gathered defects have landed).
struct TailArgs {
    unsigned *counter;           // nullptr: this kernel carries no tail
    double *acc;
    TailSeg seg[3];
    int nseg;
    const CommPeers *peers;      // nullptr: single rank
    unsigned long long epoch;    // >= 1, the same on every rank
    int exchange;                // all-reduce the segments across the ranks
    int barrier;                 // signal + wait even without segments (all-gather completion)
};

struct ShootOut {
    double *xend;  // [(i*nx + k) * ldo + b]
    double *sens;  // [(sens_off[i] + d*nx + k) * ldo + b]
    double *traj;  // [((samp_off[i] + j) * nrow + r) * ldo + b]
    double *tsens; // [(tsens_off[i] + (j * ns_i + d) * nx + k) * ldo + b]: sensitivities at the interval's j-th sample
    long long ldo;
    // fused epilogue (K3): defects in both orderings, the objective state of every interval, objective gradient
    double *dk;        // [pos_kind * ldo + b]
    double *dst;       // [pos_stage * ldo + b]
    double *objq;      // [i * ldo + b] end value of state obj_state on interval i
    double *grad;      // [var * ldo + b]
    double *gradc;     // [(var - gradc_first) * ldo + b] compact gradient: only the variables the objective depends on
    int gradc_first;   // first variable of that range (the range ends at nvars)
    double *grad_sum;  // [var] sum over the candidates of this launch (atomicAdd of warp partial sums)
    double *jac;       // [slot * ldo + b] values of d(shooting_constraints)/d(ps) in cb200_jac_structure order
    int *status;       // [b] failure flags (atomicOr): bit 0 = a non-finite end state on some interval
    int obj_state;     // 0-based state row of the objective, -1: none
    int obj_all;       // 1: every interval contributes to the objective (quadrature state), 0: last interval only
    long long ldk;     // leading dimension of dk (= ldo, or the candidates of ALL ranks when the defects are all-gathered)
    // fused all-gather of the kind-major defects: the same value also goes to rank p's window dk_peers[p] + dk_peer_off
    // (P2P stores over NVLink), p != dk_self; n_dk_peers == 0: none
    double *const *dk_peers;
    long long dk_peer_off;
    int n_dk_peers, dk_self;
    // fused Fisher read-out of the OED models (FIM_END): threads of the last interval emit F = F_init + Symmetric(F_triu)
    double *fim;       // [(r + n c) * ldo + b] or nullptr
    double *fsum;      // [n*n] accumulators (warp shuffle + atomicAdd) or nullptr
    double *crit;      // [6][ldo] or nullptr
    const double *finit;
    int fim_fused;     // 1: the kernel does the read-out
    // small-evaluation tail (shoot.cuh small_tail): the last block computes the objective and copies the result block out
    unsigned *st_counter;   // nullptr: the kernel carries no such tail
    double *st_obj;         // [B] objective inside the result block, or nullptr
    const double *st_src;   // device result block
    double *st_dst;         // the caller's pinned mirror (device-accessible host memory)
    int st_n;               // doubles to copy
};

// launchers, one per model translation unit; return cudaError_t as int, or -1000 if (G, method) is not
// instantiated for the model
int launch_lotka3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
                  const ShootOut &o, cudaStream_t s);
int launch_lotka2(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
                  const ShootOut &o, cudaStream_t s);
int launch_lqr(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
               const ShootOut &o, cudaStream_t s);
int launch_lin1d(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
                 const ShootOut &o, cudaStream_t s);
int launch_egerstedt(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
                     const ShootOut &o, cudaStream_t s);
int launch_lotka2_oed(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
                      const ShootOut &o, cudaStream_t s);
int launch_lin1d_oed(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
                     const ShootOut &o, cudaStream_t s);
int launch_dense20(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
                   const ShootOut &o, cudaStream_t s);
int forward_lotka3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int forward_lotka2(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int forward_lqr(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int forward_lin1d(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int forward_egerstedt(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int forward_lotka2_oed(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int forward_lin1d_oed(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int forward_dense20(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_vdp3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_vdp3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_cartpole5(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_cartpole5(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_rocket3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_rocket3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_quadrotor7(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_quadrotor7(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_threetank4(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_threetank4(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_batchreactor2(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_batchreactor2(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lotkacomp3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lotkacomp3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_f8aircraft4(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_f8aircraft4(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_doubleosc6(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_doubleosc6(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lotka2_oed_cu(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lotka2_oed_cu(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lotka2_oed_cf(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lotka2_oed_cf(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lotka2_oed_ds(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lotka2_oed_ds(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lotka2_oed_du(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lotka2_oed_du(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lin1d_oed_cf(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lin1d_oed_cf(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lin1d_oed_ds(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lin1d_oed_ds(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_brysondenham3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_brysondenham3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_catalystmixing2(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_catalystmixing2(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_cushioned3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_cushioned3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_denbigh3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_denbigh3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_dielectrophoretic3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_dielectrophoretic3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_electriccar4(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_electriccar4(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_fuller3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_fuller3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_jackson3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_jackson3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_mountaincar3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_mountaincar3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_particlesteering5(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_particlesteering5(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_raomease2(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_raomease2(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_robbins4(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_robbins4(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_moonlanding3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_moonlanding3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_lotkashared4(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_lotkashared4(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_hangingchain3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_hangingchain3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_tubularreactor2(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_tubularreactor2(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_ocean4(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_ocean4(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_ductedfan7(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_ductedfan7(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int launch_hangglider4(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s);
int forward_hangglider4(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s);
int dense20_set_data(const double *host_data, long long n);
#define CB200_DECL_DENSE(NXV)                                                                                             \
    int launch_dense##NXV(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B, const ShootOut &o, \
                          cudaStream_t s);                                                                                \
    int forward_dense##NXV(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,   \
                           cudaStream_t s);                                                                               \
    int dense##NXV##_set_data(const double *host_data, long long n);
CB200_DECL_DENSE(16)
CB200_DECL_DENSE(24)
CB200_DECL_DENSE(32)
#undef CB200_DECL_DENSE

}  // namespace cb200
