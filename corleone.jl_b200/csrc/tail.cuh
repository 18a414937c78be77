// Tail of an evaluation: cross-candidate sums -> destination, fused with the multi-GPU exchange.
//
// The reference combines per-interval results across workers with Distributed.pmap (src/Corleone.jl:24): every result
// is serialised over TCP to the master.  Here the only cross-GPU traffic of an evaluation is
//   * the all-reduce of the cross-candidate partial sums (objective, gradient, Fisher information), and
//   * the all-gather of the shooting defects,
// and both are done by the kernels that produce the data: the shooting kernels store the defects straight into every
// rank's window (P2P stores over NVLink, shoot.cuh / shoot_dense_mma.cuh), and the LAST block of the last kernel of the
// evaluation runs a one-shot all-reduce through the same windows -- push the local partial sums into "my" slot on every
// rank, release-store a flag per peer, acquire-wait for every peer's flag, add the slots in rank order (every rank gets
// bit-identical sums).  The flag exchange is at the same time the completion barrier of the all-gather: a peer sets its
// flag after all its stores of the epoch.  Slots, flags and gather buffers are double-buffered by epoch parity, so a
// rank may run one evaluation ahead of its slowest peer without overwriting data that is still being read.
#pragma once
#include "common.cuh"

namespace cb200 {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// run by every thread of ONE block after all contributions to t.acc are globally visible
__device__ __forceinline__ void tail_run(const TailArgs &t)
{
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nt = blockDim.x * blockDim.y * blockDim.z;
    const CommPeers *P = t.peers;
    if (!P || (!t.exchange && !t.barrier)) {   // single rank: sums -> destination, accumulators back to zero
        for (int s = 0; s < t.nseg; s++)
            for (int j = tid; j < t.seg[s].len; j += nt) {
                const double v = __ldcg(t.acc + t.seg[s].off + j);
                t.acc[t.seg[s].off + j] = 0.0;
                if (t.seg[s].dst) t.seg[s].dst[j] = v;
            }
        return;
    }
    const int n = P->nranks, me = P->rank, par = (int)(t.epoch & 1ull);
    if (t.exchange) {   // push: my partial sums into slot (par, me) of every rank's window
        int pos = 0;
        for (int s = 0; s < t.nseg; s++) {
            for (int j = tid; j < t.seg[s].len; j += nt) {
                const double v = __ldcg(t.acc + t.seg[s].off + j);
                t.acc[t.seg[s].off + j] = 0.0;
                for (int p = 0; p < n; p++) P->slots[p][((long long)(par * n + me)) * P->s_cap + pos + j] = v;
            }
            pos += t.seg[s].len;
        }
    } else {
        for (int s = 0; s < t.nseg; s++)
            for (int j = tid; j < t.seg[s].len; j += nt) {
                const double v = __ldcg(t.acc + t.seg[s].off + j);
                t.acc[t.seg[s].off + j] = 0.0;
                if (t.seg[s].dst) t.seg[s].dst[j] = v;
            }
    }
    __threadfence_system();
    __syncthreads();
    // signal: thread p tells rank p that everything this rank stores in epoch t.epoch (sums and gathered defects,
    // the latter by earlier kernels and blocks -- ordered before this point by the last-block protocol) is out
    __shared__ int timed_out;
    if (tid == 0) timed_out = 0;
    __syncthreads();
    if (tid < n) {
        __threadfence_system();
        st_release_sys(P->flags[tid] + (long long)(par * n + me) * FLAG_STRIDE, t.epoch);
        // wait: rank `tid` has signalled me
        const unsigned long long *f = P->flags[me] + (long long)(par * n + tid) * FLAG_STRIDE;
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(f) < t.epoch) {
            if ((++spins & 1023u) == 0 && (long long)(globaltimer_ns() - t0) > P->timeout_ns) {
                atomicExch(&timed_out, 1 + tid);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (timed_out) {   // a peer never arrived (crashed rank, ranks out of step): report instead of hanging the GPU
        if (tid == 0) {
            *(volatile unsigned long long *)P->err = (t.epoch << 8) | (unsigned long long)timed_out;
            __threadfence_system();
        }
        return;
    }
    if (t.exchange) {   // every rank adds the n slots in rank order: bit-identical results everywhere
        int pos = 0;
        for (int s = 0; s < t.nseg; s++) {
            for (int j = tid; j < t.seg[s].len; j += nt) {
                double sum = 0.0;
                for (int p = 0; p < n; p++) sum += ld_relaxed_sys(P->slots[me] + ((long long)(par * n + p)) * P->s_cap + pos + j);
                if (t.seg[s].dst) t.seg[s].dst[j] = sum;
            }
            pos += t.seg[s].len;
        }
    }
}

// Last-block detection: call at the very end of a kernel, by every thread that is still alive, with block-uniform
// control flow.  The block that increments the counter last owns the tail.
__device__ __forceinline__ void tail_maybe(const TailArgs &t)
{
    if (!t.counter) return;
    __shared__ int is_last;
    __threadfence();   // this thread's accumulator atomics / defect stores are visible before the count
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        const unsigned prev = atomicAdd(t.counter, 1u);
        is_last = (prev == total - 1);
        if (is_last) *t.counter = 0;   // zero at rest
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        tail_run(t);
    }
}

// stand-alone tail for pipelines whose last kernel cannot carry it (chunks on several streams)
__global__ void tail_kernel(const TailArgs t) { tail_run(t); }

}  // namespace cb200
