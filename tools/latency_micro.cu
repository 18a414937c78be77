// Microbenchmark: fixed cost of one small evaluation (BASELINE config 1) around its kernels, for the ways the results can
// travel.  A "kernel" here spins for a given time; everything else is what cb200_eval's small path does.
//   A  graph: H2D copy (2 KB) -> kernel -> small kernel -> D2H copy (8 KB); cudaStreamSynchronize      (the current path)
//   B  graph: kernel (reads pinned host memory itself) -> small kernel (writes the results to pinned host memory); sync
//   C  as B, one kernel only
//   D  as C, completion by a flag in pinned host memory the host polls instead of cudaStreamSynchronize
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_micro latency_micro.cu && ./latency_micro
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void spin(const double *in, double *out, int n_in, int n_out, long long ns, volatile unsigned long long *flag,
                     unsigned long long token)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_in; i += blockDim.x) acc += in[i];
    for (;;) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if ((long long)(t - t0) >= ns) break;
    }
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = acc + i;
    if (flag) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *flag = token;
    }
}

int main()
{
    const int n_in = 237, n_out = 1034;
    const long long kern_ns = 20000;
    double *h_in, *h_out, *d_in, *d_out;
    unsigned long long *h_flag;
    cudaHostAlloc(&h_in, 8 * n_in, cudaHostAllocDefault);
    cudaHostAlloc(&h_out, 8 * n_out, cudaHostAllocDefault);
    cudaHostAlloc(&h_flag, 8, cudaHostAllocDefault);
    cudaMalloc(&d_in, 8 * n_in);
    cudaMalloc(&d_out, 8 * n_out);
    for (int i = 0; i < n_in; i++) h_in[i] = i;
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    const char *names[4] = {"A copies + 2 kernels + sync", "B zero-copy, 2 kernels + sync", "C zero-copy, 1 kernel + sync",
                            "D zero-copy, 1 kernel + host flag"};
    for (int mode = 0; mode < 4; mode++) {
        cudaGraph_t g;
        cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (mode == 0) {
            cudaMemcpyAsync(d_in, h_in, 8 * n_in, cudaMemcpyHostToDevice, s);
            spin<<<4, 64, 0, s>>>(d_in, d_out, n_in, n_out, kern_ns, nullptr, 0);
            spin<<<1, 256, 0, s>>>(d_in, d_out, 1, 1, 0, nullptr, 0);
            cudaMemcpyAsync(h_out, d_out, 8 * n_out, cudaMemcpyDeviceToHost, s);
        } else if (mode == 1) {
            spin<<<4, 64, 0, s>>>(h_in, d_out, n_in, n_out, kern_ns, nullptr, 0);
            spin<<<1, 256, 0, s>>>(d_out, h_out, n_out, n_out, 0, nullptr, 0);
        } else {
            spin<<<1, 256, 0, s>>>(h_in, h_out, n_in, n_out, kern_ns, mode == 3 ? h_flag : nullptr, 1);
        }
        cudaStreamEndCapture(s, &g);
        cudaGraphInstantiate(&ge, g, 0);
        const int iters = 3000;
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            auto t0 = std::chrono::steady_clock::now();
            for (int it = 0; it < iters; it++) {
                if (mode == 3) {
                    *(volatile unsigned long long *)h_flag = 0;
                    cudaGraphLaunch(ge, s);
                    while (*(volatile unsigned long long *)h_flag == 0) {}
                } else {
                    cudaGraphLaunch(ge, s);
                    cudaStreamSynchronize(s);
                }
            }
            cudaStreamSynchronize(s);
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
            if (us < best) best = us;
        }
        printf("%-36s %6.2f us per call, %5.2f us around the %.0f us kernel\n", names[mode], best, best - kern_ns / 1000.0, kern_ns / 1000.0);
        cudaGraphExecDestroy(ge);
        cudaGraphDestroy(g);
    }
    return 0;
}
