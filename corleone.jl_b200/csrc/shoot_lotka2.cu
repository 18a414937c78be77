// Instantiations of the shooting kernel for model Lotka2 (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_lotka2(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_GA(Lotka2, 0, 1)
        CB200_DISPATCH_GA(Lotka2, 2, 1)
    default:
        return -1000;
    }
}
int forward_lotka2(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Lotka2, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
