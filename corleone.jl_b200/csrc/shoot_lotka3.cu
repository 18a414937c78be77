// Instantiations of the shooting kernel for model Lotka3 (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_lotka3(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_GA(Lotka3, 0, 1)
        CB200_DISPATCH_GA(Lotka3, 1, 1)
        CB200_DISPATCH_GA(Lotka3, 2, 1)
        CB200_DISPATCH_GA(Lotka3, 5, 1)
    default:
        return -1000;
    }
}
int forward_lotka3(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Lotka3, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
