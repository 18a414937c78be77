// Microbenchmark: FP64 FMA issue rate per SM sub-partition as a function of resident warps and per-thread ILP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_micro dfma_micro.cu && ./dfma_micro
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double *out, int iters, double m, double c)
{
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int ILP>
void run(int warps_per_sm)
{
    double *out;
    cudaMalloc(&out, 1 << 20);
    const int iters = 4096;
    k<ILP><<<148, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    k<ILP><<<148, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    double cyc;
    cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    const double per_smsp = warps_per_sm / 4.0;
    printf("warps/SMSP %.0f ILP %d: %.2f cycles per DFMA per warp, %.3f DFMA warp-instr/cycle/SMSP\n", per_smsp, ILP,
           cyc / ((double)iters * ILP), per_smsp * iters * ILP / cyc);
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 12, 16}) {
        run<1>(w);
        run<2>(w);
        run<4>(w);
        run<8>(w);
        run<16>(w);
    }
    return 0;
}
