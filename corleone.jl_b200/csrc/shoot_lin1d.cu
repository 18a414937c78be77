// Instantiations of the shooting kernel for model Lin1d (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_lin1d(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_GA(Lin1d, 0, 1)
        CB200_DISPATCH_GA(Lin1d, 4, 1)
    default:
        return -1000;
    }
}
int forward_lin1d(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Lin1d, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
