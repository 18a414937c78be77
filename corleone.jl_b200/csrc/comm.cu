// Multi-GPU data plane behind the C ABI (include/corleone_b200.h, cb200_comm_*): one rank per GPU of one box.
//
// The reference distributes shooting intervals with Distributed.pmap (src/Corleone.jl:24, src/multiple_shooting.jl:
// 118-130): every worker's result travels back to the master over TCP.  Here a rank (process or thread, one GPU each)
// owns a contiguous slice of the candidate axis; per evaluation only the cross-candidate partial sums (objective,
// gradient, Fisher information) are all-reduced and the shooting defects all-gathered, over NVLink:
//
//   transport P2P  (default)  every rank maps every other rank's window (CUDA IPC between processes, peer access
//                             between threads of one process); the shooting kernels store the defects into all windows
//                             and the last block of the evaluation's last kernel runs a one-shot all-reduce (tail.cuh).
//                             No NCCL call per evaluation.
//   transport NCCL (fallback) ncclAllReduce / ncclAllGather on the handle's stream after the kernels
//                             (CB200_COMM_TRANSPORT=nccl, or when a peer window cannot be mapped).
//
// Bootstrap: an NCCL unique id the host distributes (cb200_comm_unique_id / cb200_comm_init), or a host-provided
// all-gather callback (cb200_comm_init_host -- e.g. MPI.Allgather or Distributed.jl on the Julia side); NCCL is
// dlopen'ed on first use, the library has no link-time dependency on it.
#include "handle.cuh"

#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <unistd.h>

#include <nccl.h>   // types only; the entry points are resolved with dlsym

using namespace cb200;

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi &nccl_api()
{
    static NcclApi api = [] {
        NcclApi a;
        // the copy the process already holds (torch brings its own) wins; then the system library
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            a.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) {
            if (const char *e = getenv("CB200_NCCL_LIB")) a.lib = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
            for (const char *n : names) {
                if (a.lib) break;
                a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            }
        }
        if (!a.lib) return a;
#define CB_SYM(field, name) a.field = (decltype(a.field))dlsym(a.lib, name)
        CB_SYM(GetUniqueId, "ncclGetUniqueId");
        CB_SYM(CommInitRank, "ncclCommInitRank");
        CB_SYM(CommDestroy, "ncclCommDestroy");
        CB_SYM(AllReduce, "ncclAllReduce");
        CB_SYM(AllGather, "ncclAllGather");
        CB_SYM(GetErrorString, "ncclGetErrorString");
#undef CB_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.AllGather && a.GetErrorString;
        return a;
    }();
    return api;
}

#define NCCL_TRY(expr)                                                                                          \
    do {                                                                                                        \
        ncclResult_t r_ = (expr);                                                                               \
        if (r_ != ncclSuccess)                                                                                  \
            return fail(CB200_ERR_COMM, "%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(r_), __FILE__, \
                        __LINE__);                                                                              \
    } while (0)

struct PeerRecord {   // what the ranks tell each other about one window
    long long pid;
    unsigned long long host;
    int device;
    int ok;
    void *ptr;
    cudaIpcMemHandle_t ipc;
};

unsigned long long host_id()
{
    char name[256] = {0};
    gethostname(name, sizeof name - 1);
    unsigned long long hsh = 1469598103934665603ull;
    for (const char *c = name; *c; c++) hsh = (hsh ^ (unsigned char)*c) * 1099511628211ull;
    // the boot id separates containers that share a hostname
    if (FILE *f = fopen("/proc/sys/kernel/random/boot_id", "r")) {
        char b[64] = {0};
        if (fgets(b, sizeof b, f))
            for (const char *c = b; *c; c++) hsh = (hsh ^ (unsigned char)*c) * 1099511628211ull;
        fclose(f);
    }
    return hsh;
}

// re-orders the NCCL all-gather result [rank][row][B] into the window layout [row][B_total]
__global__ void gather_reorder_kernel(const double *__restrict__ in, double *__restrict__ out, long long n_rows, long long B,
                                      int nranks)
{
    const long long Bt = B * nranks;
    const long long tot = n_rows * Bt;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
        const long long row = q / Bt, bg = q - row * Bt;
        const long long r = bg / B, b = bg - r * B;
        out[q] = in[(r * n_rows + row) * B + b];
    }
}

}  // namespace

struct CommState {
    int nranks = 1, rank = 0;
    int transport = CB200_TRANSPORT_NONE;
    cb200_allgather_fn fn = nullptr;
    void *fn_ctx = nullptr;
    ncclComm_t nccl = nullptr;
    // control window: [flags: 2 * nranks * FLAG_STRIDE u64 | slots: 2 * nranks * s_cap doubles]
    void *ctrl = nullptr;
    size_t ctrl_bytes = 0, slots_off = 0;
    int s_cap = 0;
    void *peer_ctrl[MAX_RANKS] = {};
    bool ctrl_ipc[MAX_RANKS] = {};
    // gather window: [2][gat_cap] doubles
    void *gat = nullptr;
    long long gat_cap = 0, gat_rows = 0, gat_B = 0;
    void *peer_gat[MAX_RANKS] = {};
    bool gat_ipc[MAX_RANKS] = {};
    CommPeers hp{};
    CommPeers *dp = nullptr;
    unsigned long long epoch = 0;
    unsigned long long *err_host = nullptr;   // mapped pinned word the tail writes on a time-out
    DevBuf red_stage, gat_stage, xchg;        // NCCL transport: staged sums, all-gather target; bootstrap buffer
    TailSeg real_seg[3] = {};                 // NCCL transport: where the reduced sums finally go
    int n_real = 0;
};

namespace {

// all-gather of `bytes` per rank through the bootstrap channel (host callback or NCCL); recv holds nranks * bytes
int bootstrap_allgather(cb200_handle *h, CommState *c, const void *send, size_t bytes, void *recv)
{
    if (c->fn) {
        const int rc = c->fn(c->fn_ctx, send, (int64_t)bytes, recv);
        if (rc != 0) return fail(CB200_ERR_COMM, "the host all-gather callback returned %d", rc);
        return CB200_OK;
    }
    if (!c->nccl) return fail(CB200_ERR_COMM, "communicator has no bootstrap channel");
    if (c->xchg.reserve(bytes * (size_t)(c->nranks + 1))) return fail(CB200_ERR_NOMEM, "device out of memory");
    char *d_send = (char *)c->xchg.p, *d_recv = d_send + bytes;
    CUDA_TRY(cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, h->stream));
    NCCL_TRY(nccl_api().AllGather(d_send, d_recv, bytes, ncclChar, c->nccl, h->stream));
    CUDA_TRY(cudaMemcpyAsync(recv, d_recv, bytes * (size_t)c->nranks, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return CB200_OK;
}

// Map every rank's window `local` into this process.  peer[p] = usable device pointer or nullptr.  Collective.
// Returns CB200_OK with *all_ok = whether EVERY rank mapped EVERY window (the ranks agree on the answer).
int map_windows(cb200_handle *h, CommState *c, void *local, void **peer, bool *is_ipc, bool *all_ok)
{
    const int n = c->nranks;
    PeerRecord mine{};
    mine.pid = (long long)getpid();
    mine.host = host_id();
    mine.device = h->device;
    mine.ptr = local;
    mine.ok = 1;
    if (local && cudaIpcGetMemHandle(&mine.ipc, local) != cudaSuccess) {
        cudaGetLastError();
        mine.ok = 0;   // still usable by the threads of this process
    }
    std::vector<PeerRecord> all((size_t)n);
    int rc = bootstrap_allgather(h, c, &mine, sizeof mine, all.data());
    if (rc) return rc;
    int ok = local ? 1 : 0;
    for (int p = 0; p < n && ok; p++) {
        peer[p] = nullptr;
        is_ipc[p] = false;
        const PeerRecord &r = all[(size_t)p];
        if (p == c->rank) {
            peer[p] = local;
            continue;
        }
        if (!r.ptr || r.host != mine.host) {
            ok = 0;
            break;
        }
        if (r.pid == mine.pid) {   // another thread of this process: plain (peer) access
            if (r.device != h->device) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, h->device, r.device) != cudaSuccess || !can) {
                    cudaGetLastError();
                    ok = 0;
                    break;
                }
                const cudaError_t e = cudaDeviceEnablePeerAccess(r.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                    ok = 0;
                    break;
                }
                cudaGetLastError();
            }
            peer[p] = r.ptr;
        } else {
            void *q = nullptr;
            if (!r.ok || cudaIpcOpenMemHandle(&q, r.ipc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            peer[p] = q;
            is_ipc[p] = true;
        }
    }
    // the ranks must agree: one failed mapping anywhere puts everybody on the fallback transport
    std::vector<int> oks((size_t)n, 0);
    rc = bootstrap_allgather(h, c, &ok, sizeof ok, oks.data());
    if (rc) return rc;
    bool every = true;
    for (int v : oks) every = every && (v != 0);
    if (!every) {
        for (int p = 0; p < n; p++) {
            if (is_ipc[p] && peer[p]) cudaIpcCloseMemHandle(peer[p]);
            peer[p] = nullptr;
            is_ipc[p] = false;
        }
    }
    *all_ok = every;
    return CB200_OK;
}

int comm_setup(cb200_handle *h, CommState *c, int64_t max_B_total)
{
    const int n = c->nranks;
    const int n2 = h->fim_n * h->fim_n;
    c->s_cap = 1 + h->L.nvars + n2;
    const size_t flag_bytes = sizeof(unsigned long long) * 2 * (size_t)n * FLAG_STRIDE;
    c->slots_off = (flag_bytes + 255) & ~size_t(255);
    c->ctrl_bytes = c->slots_off + sizeof(double) * 2 * (size_t)n * (size_t)c->s_cap;
    CUDA_TRY(cudaMalloc(&c->ctrl, c->ctrl_bytes));
    CUDA_TRY(cudaMemset(c->ctrl, 0, c->ctrl_bytes));
    const long long n_rows = (long long)h->defects.size();
    if (max_B_total > 0 && n_rows > 0) {
        c->gat_cap = n_rows * max_B_total;
        c->gat_rows = n_rows;
        c->gat_B = max_B_total;
        CUDA_TRY(cudaMalloc(&c->gat, sizeof(double) * 2 * (size_t)c->gat_cap));
        CUDA_TRY(cudaMemset(c->gat, 0, sizeof(double) * 2 * (size_t)c->gat_cap));
    }
    CUDA_TRY(cudaHostAlloc((void **)&c->err_host, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
    *c->err_host = 0;
    unsigned long long *err_dev = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer((void **)&err_dev, c->err_host, 0));
    CUDA_TRY(cudaDeviceSynchronize());   // the zeroed windows are in place before any peer can touch them

    const char *want = getenv("CB200_COMM_TRANSPORT");
    bool p2p = !(want && strcmp(want, "nccl") == 0);
    bool ok_ctrl = false, ok_gat = true;
    if (p2p) {
        int rc = map_windows(h, c, c->ctrl, c->peer_ctrl, c->ctrl_ipc, &ok_ctrl);
        if (rc) return rc;
        if (ok_ctrl && c->gat) {
            rc = map_windows(h, c, c->gat, c->peer_gat, c->gat_ipc, &ok_gat);
            if (rc) return rc;
        }
        p2p = ok_ctrl && ok_gat;
    }
    if (p2p) {
        c->transport = CB200_TRANSPORT_P2P;
    } else {
        if (!c->nccl)
            return fail(CB200_ERR_COMM, "the peer windows cannot be mapped (no P2P / IPC between the ranks) and the "
                                        "communicator has no NCCL transport to fall back to");
        if (want && strcmp(want, "p2p") == 0) return fail(CB200_ERR_COMM, "CB200_COMM_TRANSPORT=p2p, but the peer windows cannot be mapped");
        c->transport = CB200_TRANSPORT_NCCL;
    }
    CommPeers &P = c->hp;
    P.nranks = n;
    P.rank = c->rank;
    P.s_cap = c->s_cap;
    for (int p = 0; p < n; p++) {
        char *base = (char *)(c->transport == CB200_TRANSPORT_P2P ? c->peer_ctrl[p] : (p == c->rank ? c->ctrl : nullptr));
        P.flags[p] = (unsigned long long *)base;
        P.slots[p] = base ? (double *)(base + c->slots_off) : nullptr;
        P.gather[p] = (double *)(c->transport == CB200_TRANSPORT_P2P ? c->peer_gat[p] : (p == c->rank ? c->gat : nullptr));
    }
    P.gather_cap = c->gat_cap;
    P.err = err_dev;
    long long tmo_ms = 20000;
    if (const char *e = getenv("CB200_COMM_TIMEOUT_MS")) tmo_ms = atoll(e) > 0 ? atoll(e) : tmo_ms;
    P.timeout_ns = tmo_ms * 1000000ll;
    CUDA_TRY(cudaMalloc((void **)&c->dp, sizeof(CommPeers)));
    CUDA_TRY(cudaMemcpy(c->dp, &P, sizeof(CommPeers), cudaMemcpyHostToDevice));
    // nobody starts an epoch before every rank has its tables in place
    int one = 1;
    std::vector<int> ones((size_t)n);
    return bootstrap_allgather(h, c, &one, sizeof one, ones.data());
}

int comm_init_common(cb200_handle *h, int32_t nranks, int32_t rank, int64_t max_B_total, CommState *c)
{
    c->nranks = nranks;
    c->rank = rank;
    const int rc = comm_setup(h, c, max_B_total);
    if (rc) {
        h->comm = c;
        comm_release(h);
        return rc;
    }
    h->comm = c;
    return CB200_OK;
}

int check_init_args(cb200_handle *h, int32_t nranks, int32_t rank, int64_t max_B_total)
{
    if (!h) return fail(CB200_ERR_ARG, "null handle");
    if (nranks < 1 || nranks > MAX_RANKS) return fail(CB200_ERR_ARG, "nranks must be in 1..%d (one box)", MAX_RANKS);
    if (rank < 0 || rank >= nranks) return fail(CB200_ERR_ARG, "rank %d out of 0..%d", rank, nranks - 1);
    if (max_B_total < 0) return fail(CB200_ERR_ARG, "negative max_B_total");
    if (h->comm) return fail(CB200_ERR_ARG, "the handle already has a communicator (cb200_comm_destroy first)");
    return CB200_OK;
}

}  // namespace

extern "C" int cb200_comm_unique_id(void *id)
{
    if (!id) return fail(CB200_ERR_ARG, "null argument");
    static_assert(sizeof(ncclUniqueId) == CB200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    if (!nccl_api().ok) return fail(CB200_ERR_COMM, "libnccl.so.2 not found (set CB200_NCCL_LIB, or bootstrap with cb200_comm_init_host)");
    ncclUniqueId u;
    NCCL_TRY(nccl_api().GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return CB200_OK;
}

extern "C" int cb200_comm_init(cb200_handle *h, int32_t nranks, int32_t rank, const void *id, int64_t max_B_total)
{
    int rc = check_init_args(h, nranks, rank, max_B_total);
    if (rc) return rc;
    if (!id) return fail(CB200_ERR_ARG, "null unique id");
    if (!nccl_api().ok) return fail(CB200_ERR_COMM, "libnccl.so.2 not found (set CB200_NCCL_LIB, or bootstrap with cb200_comm_init_host)");
    std::lock_guard<std::mutex> lk(h->mu);
    CUDA_TRY(cudaSetDevice(h->device));
    CommState *c = new (std::nothrow) CommState();
    if (!c) return fail(CB200_ERR_NOMEM, "out of host memory");
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    const ncclResult_t r = nccl_api().CommInitRank(&c->nccl, nranks, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(CB200_ERR_COMM, "ncclCommInitRank failed: %s", nccl_api().GetErrorString(r));
    }
    return comm_init_common(h, nranks, rank, max_B_total, c);
}

extern "C" int cb200_comm_init_host(cb200_handle *h, int32_t nranks, int32_t rank, cb200_allgather_fn fn, void *ctx,
                                    int64_t max_B_total)
{
    int rc = check_init_args(h, nranks, rank, max_B_total);
    if (rc) return rc;
    if (!fn) return fail(CB200_ERR_ARG, "null all-gather callback");
    // no lock across the callbacks: the ranks may be threads of one process that call in here concurrently
    CUDA_TRY(cudaSetDevice(h->device));
    CommState *c = new (std::nothrow) CommState();
    if (!c) return fail(CB200_ERR_NOMEM, "out of host memory");
    c->fn = fn;
    c->fn_ctx = ctx;
    return comm_init_common(h, nranks, rank, max_B_total, c);
}

extern "C" int cb200_comm_destroy(cb200_handle *h)
{
    if (!h) return CB200_OK;
    std::lock_guard<std::mutex> lk(h->mu);
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    comm_release(h);
    return CB200_OK;
}

extern "C" int cb200_comm_query(const cb200_handle *h, cb200_comm_info *info)
{
    if (!h || !info) return fail(CB200_ERR_ARG, "null argument");
    memset(info, 0, sizeof *info);
    info->nranks = 1;
    if (!h->comm) return CB200_OK;
    const CommState *c = h->comm;
    info->nranks = c->nranks;
    info->rank = c->rank;
    info->transport = c->transport;
    info->epoch = (int64_t)c->epoch;
    info->max_B_total = c->gat_B;
    info->error = c->err_host ? (int64_t)*(volatile unsigned long long *)c->err_host : 0;
    return CB200_OK;
}

extern "C" int cb200_comm_gathered(const cb200_handle *h, double **ptr, int64_t *ld)
{
    if (!h || !ptr) return fail(CB200_ERR_ARG, "null argument");
    const CommState *c = h->comm;
    if (!c || !c->gat || c->epoch == 0) return fail(CB200_ERR_ARG, "no all-gathered defects on this handle");
    *ptr = (double *)c->gat + (long long)(c->epoch & 1ull) * c->gat_cap;
    if (ld) *ld = h->last_B_total;
    return CB200_OK;
}

namespace cb200 {

void comm_release(cb200_handle *h)
{
    CommState *c = h->comm;
    if (!c) return;
    for (int p = 0; p < MAX_RANKS; p++) {
        if (c->ctrl_ipc[p] && c->peer_ctrl[p]) cudaIpcCloseMemHandle(c->peer_ctrl[p]);
        if (c->gat_ipc[p] && c->peer_gat[p]) cudaIpcCloseMemHandle(c->peer_gat[p]);
    }
    if (c->nccl && nccl_api().ok) nccl_api().CommDestroy(c->nccl);
    if (c->dp) cudaFree(c->dp);
    if (c->ctrl) cudaFree(c->ctrl);
    if (c->gat) cudaFree(c->gat);
    if (c->err_host) cudaFreeHost(c->err_host);
    c->red_stage.release();
    c->gat_stage.release();
    c->xchg.release();
    cudaGetLastError();
    delete c;
    h->comm = nullptr;
}

int comm_check(cb200_handle *h)
{
    CommState *c = h->comm;
    if (!c || !c->err_host) return CB200_OK;
    const unsigned long long e = *(volatile unsigned long long *)c->err_host;
    if (e)
        return fail(CB200_ERR_COMM, "rank %d: peer %d did not arrive in epoch %llu within the time-out (a rank died, or the "
                                    "ranks do not issue the same sequence of collective evaluations)",
                    c->rank, (int)(e & 0xff) - 1, e >> 8);
    return CB200_OK;
}

bool comm_is_p2p(const cb200_handle *h) { return h->comm && h->comm->transport == CB200_TRANSPORT_P2P; }

// Open the next epoch: fills the tail arguments (which segments are all-reduced, completion barrier) and the
// all-gather part of the shooting kernels' output description.
int comm_begin(cb200_handle *h, bool reduce, bool gather, long long B, long long B_total, long long b0, TailArgs *tail,
               ShootOut *so, double **dk_local)
{
    CommState *c = h->comm;
    if (!c) return fail(CB200_ERR_ARG, "CB200_COMM_REDUCE / CB200_COMM_GATHER need a communicator (cb200_comm_init)");
    int rc = comm_check(h);
    if (rc) return rc;
    if (gather) {
        if (!c->gat) return fail(CB200_ERR_ARG, "the communicator was created without a gather window (max_B_total == 0) or the layer has no defects");
        if (B_total < 1 || b0 < 0 || b0 + B > B_total)
            return fail(CB200_ERR_ARG, "all-gather: candidates [%lld, %lld) do not fit comm_B_total = %lld", b0, b0 + B, B_total);
        if (c->gat_rows * B_total > c->gat_cap)
            return fail(CB200_ERR_ARG, "all-gather: comm_B_total = %lld exceeds the window (max_B_total = %lld)", B_total, c->gat_B);
        if (c->transport == CB200_TRANSPORT_NCCL && B * c->nranks != B_total)
            return fail(CB200_ERR_UNSUPPORTED, "the NCCL transport gathers equal shards only (B * nranks == comm_B_total)");
    }
    c->epoch++;
    const long long par = (long long)(c->epoch & 1ull);
    tail->peers = nullptr;
    tail->epoch = c->epoch;
    tail->exchange = 0;
    tail->barrier = 0;
    if (c->transport == CB200_TRANSPORT_P2P) {
        tail->peers = c->dp;
        tail->exchange = reduce ? 1 : 0;
        tail->barrier = 1;   // every collective evaluation closes its epoch: keeps the ranks within one epoch of each other
        // measurement knobs (tools/gpu_tail_probe.sh): what the tail costs without the exchange / without any peer traffic
        static const int probe = getenv("CB200_TAIL_PROBE") ? atoi(getenv("CB200_TAIL_PROBE")) : 0;
        if (probe == 1) tail->exchange = 0;                       // flags only
        if (probe == 2) { tail->exchange = 0; tail->barrier = 0; }   // local tail (results are NOT reduced: timing only)
        if (gather) {
            *dk_local = (double *)c->gat + par * c->gat_cap + b0;
            so->ldk = B_total;
            so->dk_peers = c->dp->gather;   // address of the table inside the device copy
            so->dk_peer_off = par * c->gat_cap + b0;
            so->n_dk_peers = c->nranks;
            so->dk_self = c->rank;
        }
    }
    h->last_B_total = gather ? B_total : 0;
    return CB200_OK;
}

// NCCL transport: the tail left the local sums in `stage`; all-reduce them and deliver; all-gather the local defect
// block [n_rows][B] into the window layout [n_rows][B_total].
int comm_finish_nccl(cb200_handle *h, cudaStream_t st, double *stage, long long n_stage, const TailSeg *real, int n_real,
                     const double *dk_local, long long n_rows, long long B)
{
    CommState *c = h->comm;
    if (n_stage > 0) {
        NCCL_TRY(nccl_api().AllReduce(stage, stage, (size_t)n_stage, ncclDouble, ncclSum, c->nccl, st));
        long long pos = 0;
        for (int s = 0; s < n_real; s++) {
            if (real[s].dst)
                CUDA_TRY(cudaMemcpyAsync(real[s].dst, stage + pos, sizeof(double) * (size_t)real[s].len, cudaMemcpyDeviceToDevice, st));
            pos += real[s].len;
        }
    }
    if (dk_local) {
        const size_t blk = (size_t)n_rows * (size_t)B;
        if (c->gat_stage.reserve(sizeof(double) * blk * (size_t)c->nranks)) return fail(CB200_ERR_NOMEM, "device out of memory");
        NCCL_TRY(nccl_api().AllGather(dk_local, c->gat_stage.p, blk, ncclDouble, c->nccl, st));
        double *out = (double *)c->gat + (long long)(c->epoch & 1ull) * c->gat_cap;
        gather_reorder_kernel<<<592, 256, 0, st>>>((const double *)c->gat_stage.p, out, n_rows, B, c->nranks);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return CB200_OK;
}

double *comm_red_stage(cb200_handle *h, long long n)
{
    CommState *c = h->comm;
    if (c->red_stage.reserve(sizeof(double) * (size_t)std::max<long long>(n, 1))) return nullptr;
    return (double *)c->red_stage.p;
}

double *comm_gather_ptr(const cb200_handle *h)
{
    const CommState *c = h->comm;
    return (double *)c->gat + (long long)(c->epoch & 1ull) * c->gat_cap;
}

int comm_nranks(const cb200_handle *h) { return h->comm ? h->comm->nranks : 1; }

}  // namespace cb200
