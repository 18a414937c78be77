// Instantiations of the shooting kernel for model Lotka2OedDu (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_lotka2_oed_du(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_GA(Lotka2OedDu, 0, 1)
        CB200_DISPATCH_GA(Lotka2OedDu, 1, 1)
    default:
        return -1000;
    }
}
int forward_lotka2_oed_du(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Lotka2OedDu, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
