// Instantiations of the shooting kernel for model Quadrotor7 (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_quadrotor7(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_GA(Quadrotor7, 0, 1)
        CB200_DISPATCH_GA(Quadrotor7, 1, 1)
    default:
        return -1000;
    }
}
int forward_quadrotor7(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Quadrotor7, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
