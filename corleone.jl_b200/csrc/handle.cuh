// Private to the library: the opaque handle behind include/corleone_b200.h, shared by api.cu (evaluation) and
// comm.cu (multi-GPU data plane).
#pragma once
#include "../../include/corleone_b200.h"

#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

// ------------------------------------------------------------------ errors
// thread-local message behind cb200_last_error(); returns `code` so that `return cb200_fail(...)` reads well
int cb200_fail(int code, const char *fmt, ...);
#define fail cb200_fail
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(CB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                       \
    } while (0)

// ------------------------------------------------------------------ epilogue tables
struct DefectEntry {
    int a_kind;   // 0: xend row, 1: ps variable
    int a_idx;
    int b_kind;   // 1: ps variable, 2: constant
    int b_idx;
    double b_const;
    int pos_kind;   // position in shooting_constraints (kind-major)
    int pos_stage;  // position in stage_ordered_shooting_constraints
};

struct GradEntry {   // grad[var] = sens[src] (src >= 0) or 1.0 (src == -1)
    int var;
    long long src;
};

// ------------------------------------------------------------------ handle
// bumped by every (re)allocation of a workspace: captured CUDA graphs hold raw workspace pointers and are discarded
// when the generation they were captured in is over
inline unsigned long long &devbuf_generation()
{
    static unsigned long long gen = 0;
    return gen;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        devbuf_generation()++;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) return (int)e;
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct CommState;   // comm.cu

// CUDA graph of one small evaluation (B = 1 through host pointers: the reference's own use), keyed by what it computes
struct SmallGraph {
    unsigned want = 0;
    int objective_row = 0;
    long long B = 0;
    bool has_status = false, has_finit = false;
    size_t tot = 0;
    void *dp = nullptr;          // ws_small.p the graph was captured with
    unsigned res = 0, flags = 0;
    long long launches = 0;      // kernels in the graph
    unsigned long long gen = 0;  // workspace generation at capture
    cudaGraphExec_t exec = nullptr;
};

struct cb200_handle {
    std::mutex mu;
    int device = 0;
    int model = 0, method = 0;
    cb200::DevLayer L{};
    int nquad = 0;
    std::vector<int> quad0;           // 0-based quadrature states
    std::vector<int> ns;              // directions per interval
    std::vector<int> h_M, h_var_off, h_n_u0, h_n_c, h_samp_off, h_piece_off;
    std::vector<long long> h_sens_off, h_tsens_off;
    long long n_tsens = 0;
    std::vector<double> h_sample_t;   // times of the merged trajectory's samples
    bool has_saveat = false;
    std::vector<int> h_cidx, h_ic_src;
    std::vector<DefectEntry> defects;
    std::vector<int> dpos_kind, dpos_stage, oth_off, jpos;
    long long jac_nnz_var = 0;         // leading entries of the triplet list whose values come from sens
    std::vector<cb200::MatchEntry> oth;
    std::vector<long long> jac_rows, jac_cols, jac_src;
    long long n_sens = 0;
    int n_samples = 0;
    int max_ns = 0;
    int fim_off = -1, fim_n = 0;
    int fim_mode = 0, fim_bx = 0, n_obs = 0, n_obs_rows = 0;   // sample-based read-outs
    int *d_obs_sample = nullptr, *d_obs_var = nullptr;
    int range_i0 = 0, range_n = -1;   // cb200_set_interval_range: intervals [i0, i0 + n); n < 0 = all
    bool d20_ref = false;             // holds a reference on the device's DENSE20 constant block
    int G_sens = 2;                   // sensitivity directions per thread
    const unsigned *act_mask1 = nullptr;   // DevLayer::act_mask for one direction per thread (flat launches), or nullptr
    bool no_g1 = false;               // the model has no G = 1 kernel
    bool no_small_tail = false;       // the model has no kernel with the small-evaluation tail
    void *tables = nullptr;           // one device allocation holding every table
    int *d_quad = nullptr;
    GradEntry *d_grad = nullptr;      // gradient gather table of the cached objective row
    int grad_cap = 0;
    std::vector<GradEntry> h_grad;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    static constexpr int NAUX = 2;        // extra streams of the chunked host path (copy/compute overlap)
    cudaStream_t aux[NAUX] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[NAUX] = {};
    static constexpr int EV_RING = 256;   // event pairs around the integration kernel of the last 256 evals
    cudaEvent_t ev0[EV_RING] = {}, ev1[EV_RING] = {};
    long long n_timed = 0;
    long long launches = 0;
    DevBuf ws_ps, ws_in, ws_xend, ws_sens, ws_traj, ws_dk, ws_ds, ws_obj, ws_grad, ws_fim, ws_finit, ws_out, ws_sums, ws_objq, ws_jac, ws_crit, ws_fixed, ws_status, ws_small, ws_gradc, ws_tsens;
    static constexpr size_t PIN_BYTES = 1 << 20;   // pinned host mirror of small evaluations
    void *pin = nullptr;
    // cross-candidate sums [objective_sum (1) | grad_sum (nvars) | fim_sum (fim_n^2)]: accumulated here by the kernels'
    // warp-shuffle + atomicAdd epilogues; ZERO AT REST -- the tail of the last kernel of an evaluation hands the sums
    // to their destination (all-reduced over the ranks with a communicator) and clears the block again
    double *acc = nullptr;
    unsigned *tail_counter = nullptr;     // last-block detection of the kernel that carries the tail (zero at rest)
    bool acc_dirty = false;               // an evaluation failed half-way: clear before the next one
    CommState *comm = nullptr;            // cb200_comm_init*: the multi-GPU data plane of this handle (comm.cu)
    long long last_B = 0;                 // candidates of the last evaluation (leading dimension of resident outputs)
    long long last_B_total = 0;           // comm_B_total of the last all-gather
    unsigned resident_valid = 0;          // want-bits whose results of the last evaluation sit in the workspaces
    long long gradc_n_last = 0;
    std::vector<SmallGraph> graphs;       // captured small evaluations
    bool graphs_enabled = true;
};

// comm.cu: what cb200_eval needs from the communicator
namespace cb200 {
// Opens the next epoch (every rank counts the same sequence): fills the tail's exchange / barrier fields and, for an
// all-gather on the P2P transport, the peer part of the shooting kernels' output description; *dk_local = where this
// rank's kernels store their own defect columns (inside the gather window).
int comm_begin(cb200_handle *h, bool reduce, bool gather, long long B, long long B_total, long long b0, TailArgs *tail,
               ShootOut *so, double **dk_local);
// NCCL transport: all-reduce the staged sums and deliver them; all-gather the local defect block
int comm_finish_nccl(cb200_handle *h, cudaStream_t st, double *stage, long long n_stage, const TailSeg *real, int n_real,
                     const double *dk_local, long long n_rows, long long B);
double *comm_red_stage(cb200_handle *h, long long n);
double *comm_gather_ptr(const cb200_handle *h);
int comm_nranks(const cb200_handle *h);
bool comm_is_p2p(const cb200_handle *h);
int comm_check(cb200_handle *h);
void comm_release(cb200_handle *h);
}  // namespace cb200
