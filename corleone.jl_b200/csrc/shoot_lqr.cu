// Instantiations of the shooting kernel for model Lqr (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_lqr(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_GA(Lqr, 0, 1)
        CB200_DISPATCH_GA(Lqr, 4, 4)  // 127 registers, no spills: 16 warps/SM (ptxas -v); best of the G x MINB sweep
        CB200_DISPATCH_GA(Lqr, 8, 1)
    default:
        return -1000;
    }
}
int forward_lqr(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Lqr, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
