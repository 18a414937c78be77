// Instantiations of the shooting kernel for model Dense20 (see shoot.cuh).
#define CB200_NEED_DENSE20_DATA 1
#include "models.cuh"
#include "shoot.cuh"
#include "shoot_dense_mma.cuh"
#include <cstdlib>
namespace cb200 {
int launch_dense20(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    // with sensitivities: the tensor-path kernel (shoot_dense_mma.cuh).  The thread-per-direction instantiation it
    // replaced (0.064 of the FP64 roof, 6 MB of code) is gone; without sensitivities the generic kernel integrates x
    if (G > 0) return launch_dense_mma(method, L, ps, ldb, B, o, L.R * G, L.max_M, s);
    switch (G) {
        CB200_DISPATCH_G(Dense20, 0, 1)
    default:
        return -1000;
    }
}
int dense20_set_data(const double *host_data, long long n)
{
    if (n != 420) return -1;
    return (int)cudaMemcpyToSymbol(c_dense20, host_data, sizeof(double) * 420);
}
int forward_dense20(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Dense20>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
