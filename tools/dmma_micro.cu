// Microbenchmark: does FP64 mma.sync (DMMA m8n8k4) on sm_100a run on a pipe of its own, i.e. can it overlap DFMA?
// Three kernels: DMMA only, DFMA only, both interleaved.  Reports FMA/cycle/SM for each.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_micro dmma_micro.cu && ./dmma_micro
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int NM, int NF>
__global__ void k(double *out, int iters, double m, double c)
{
    double acc[NM > 0 ? 2 * NM : 1], f[NF > 0 ? NF : 1];
    for (int i = 0; i < 2 * NM; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < NF; i++) f[i] = threadIdx.x + i;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-6;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) dmma(acc[2 * i], acc[2 * i + 1], a, b);
#pragma unroll
        for (int i = 0; i < NF; i++) f[i] = fma(f[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 2 * NM; i++) s += acc[i];
    for (int i = 0; i < NF; i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int NM, int NF>
void run(int warps_per_sm)
{
    double *out;
    cudaMalloc(&out, 1 << 22);
    const int iters = 2048;
    for (int rep = 0; rep < 2; rep++) {
        k<NM, NF><<<148, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-9);
        cudaDeviceSynchronize();
    }
    double cyc;
    cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    const double fma_mma = (double)warps_per_sm * iters * NM * 256.0, fma_vec = (double)warps_per_sm * iters * NF * 32.0;
    printf("warps/SM %2d  DMMA x%d + DFMA x%2d per iter: %8.0f cycles, tensor %6.1f FMA/clk/SM, vector %6.1f FMA/clk/SM, total %6.1f\n",
           warps_per_sm, NM, NF, cyc, fma_mma / cyc, fma_vec / cyc, (fma_mma + fma_vec) / cyc);
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 16}) {
        run<4, 0>(w);
        run<8, 0>(w);
        run<0, 8>(w);
        run<0, 16>(w);
        run<4, 8>(w);
        run<4, 32>(w);
        run<8, 16>(w);
        run<2, 16>(w);
    }
    return 0;
}
