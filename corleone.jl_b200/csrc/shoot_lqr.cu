// Instantiations of the shooting kernel for model Lqr (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_lqr(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_G(Lqr, 0)
        CB200_DISPATCH_G(Lqr, 1)
        CB200_DISPATCH_G(Lqr, 2)
        CB200_DISPATCH_G(Lqr, 4)
        CB200_DISPATCH_G(Lqr, 8)
    default:
        return -1000;
    }
}
}  // namespace cb200
