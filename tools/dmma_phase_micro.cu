// Microbenchmark: the practical FP64 ceiling of the dense kernel's instruction structure (shoot_dense_mma.cuh) without
// memory traffic.  Per right-hand-side call a warp issues 15 DMMA.8x8x4 (3 accumulator chains, 5 deep) and then ~22
// dependent DFMA (k = fma(c, y, d), stage combinations); the next call's A operand depends on the DFMA results.
//   MODE 0: as in the kernel (phases of different warps drift freely)
//   MODE 1: DMMA chains only (15 per call, 3 x 5 dependent)       MODE 2: DFMA part only
//   MODE 3: two tiles per warp (6 chains, 44 DFMA), half the warps
//   MODE 4: as 0, but the DFMAs of call n are interleaved with the DMMAs of call n+1 of a second tile (software pipelining)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_phase_micro dmma_phase_micro.cu && ./dmma_phase_micro
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct Tile {
    double y[5], s3[5], s4[5], yd[5], d[6], k[5];
};

__device__ __forceinline__ void tile_mma(Tile &T, const double *Wf)
{
#pragma unroll
    for (int s = 0; s < 6; s++) T.d[s] = 1e-3 * s;
#pragma unroll
    for (int kt = 0; kt < 5; kt++) {
        dmma(T.d[0], T.d[1], T.y[kt], Wf[kt * 3 + 0]);
        dmma(T.d[2], T.d[3], T.y[kt], Wf[kt * 3 + 1]);
        dmma(T.d[4], T.d[5], T.y[kt], Wf[kt * 3 + 2]);
    }
}
__device__ __forceinline__ void tile_fma(Tile &T, double c, double h)
{
    // 5 + 17 DFMA
#pragma unroll
    for (int s = 0; s < 5; s++) T.k[s] = fma(c, T.y[s], T.d[s]);
#pragma unroll
    for (int s = 0; s < 5; s++) {
        T.s3[s] = fma(h, T.k[s], T.s3[s]);
        T.s4[s] = fma(h, T.k[s], T.s4[s]);
        T.yd[s] = fma(h, T.k[s], T.yd[s]);
    }
    T.y[0] = fma(h, T.k[0], T.s3[0]);
    T.y[1] = fma(h, T.k[1], T.s3[1]);
#pragma unroll
    for (int s = 2; s < 5; s++) T.y[s] = T.s4[s] * 0.5 + T.s3[s] * 1e-3;   // 3 DMUL + 3 DFMA ~ extra
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double *out, int iters, double c, double h)
{
    double Wf[15];
    for (int i = 0; i < 15; i++) Wf[i] = 1e-3 * (threadIdx.x % 7 + i);
    Tile A, B;
    for (int s = 0; s < 5; s++) {
        A.y[s] = 1e-3 * threadIdx.x + s; A.s3[s] = A.s4[s] = A.yd[s] = 0.1 * s;
        B.y[s] = 2e-3 * threadIdx.x + s; B.s3[s] = B.s4[s] = B.yd[s] = 0.2 * s;
    }
    for (int s = 0; s < 6; s++) A.d[s] = B.d[s] = 0;
    long long t0 = clock64();
    if (MODE == 4) tile_mma(A, Wf);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { tile_mma(A, Wf); tile_fma(A, c, h); }
        if (MODE == 1) { tile_mma(A, Wf); A.y[0] = A.d[0]; A.y[1] = A.d[2]; A.y[2] = A.d[4]; A.y[3] = A.d[1]; A.y[4] = A.d[3]; }
        if (MODE == 2) { tile_fma(A, c, h); A.d[0] = A.y[0]; A.d[1] = A.y[1]; }
        if (MODE == 3) { tile_mma(A, Wf); tile_mma(B, Wf); tile_fma(A, c, h); tile_fma(B, c, h); }
        if (MODE == 4) { tile_mma(B, Wf); tile_fma(A, c, h); tile_mma(A, Wf); tile_fma(B, c, h); }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 5; i++) s += A.y[i] + A.yd[i] + B.y[i] + B.yd[i] + A.d[i] + B.d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}

template <int MODE>
void run(int warps, const char *what)
{
    double *out;
    cudaMalloc(&out, 1 << 22);
    const int iters = 4096;
    for (int rep = 0; rep < 2; rep++) {
        k<MODE><<<148, warps * 32>>>(out, iters, -0.3, 1e-3);
        cudaDeviceSynchronize();
    }
    double cyc;
    cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    const double tiles = (MODE >= 3) ? 2 : 1;
    const double mma = (MODE == 2) ? 0 : 15 * 256.0 * tiles, vec = (MODE == 1) ? 0 : 25 * 32.0 * tiles;
    printf("%-44s warps/SM %2d: %9.0f cycles  %6.1f FMA/clk/SM (tensor %5.1f + vector %5.1f)   %7.1f clk per call-tile per SMSP\n", what, warps, cyc,
           warps * iters * (mma + vec) / cyc, warps * iters * mma / cyc, warps * iters * vec / cyc, cyc / iters / tiles / (warps / 4.0));
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 12, 16}) {
        run<0>(w, "kernel structure (15 DMMA then 25 DFMA)");
        run<1>(w, "DMMA chains only");
        run<2>(w, "DFMA part only");
    }
    for (int w : {4, 8}) {
        run<3>(w, "two tiles per warp, phases back to back");
        run<4>(w, "two tiles per warp, software-pipelined");
    }
    return 0;
}
