// Instantiations of the shooting kernels for the dense 24-state model (models.cuh DenseN, shoot_dense_mma.cuh).
#define CB200_NEED_DENSE20_DATA 1
#include "models.cuh"
#include "shoot.cuh"
#include "shoot_dense_mma.cuh"
#include <cstdlib>
namespace cb200 {
int launch_dense24(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    if (o.st_counter) return -1001;   // no kernel of this model carries the small-evaluation tail
    if (method == 2) return -1000;   // the adaptive mode of the dense family is built for the 20-state model only
    // with sensitivities: the tensor-path kernel (shoot_dense_mma.cuh); without, the generic kernel integrates x
    if (G > 0) return launch_dense_mma<24, 12>(method, L, ps, ldb, B, o, L.R * G, L.max_M, s);
    switch (G) {
        CB200_DISPATCH_G(DenseN<24>, 0, 1)
    default:
        return -1000;
    }
}
int dense24_set_data(const double *host_data, long long n)
{
    if (n != 24 * 24 + 24) return -1;
    return (int)cudaMemcpyToSymbol(c_dense, host_data, sizeof(double) * (24 * 24 + 24));
}
int forward_dense24(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<DenseN<24>>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
