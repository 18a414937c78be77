// Instantiations of the shooting kernels for the dense 20-state model (models.cuh DenseN, shoot_dense_mma.cuh).
#define CB200_NEED_DENSE20_DATA 1
#include "models.cuh"
#include "shoot.cuh"
#include "shoot_dense_mma.cuh"
#include <cstdlib>
namespace cb200 {
int launch_dense20(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    if (o.st_counter) return -1001;   // no kernel of this model carries the small-evaluation tail
    // adaptive mode (the reference's default `solve`; a fidelity path, not a throughput one): every unit picks its own
    // steps, so it runs on the generic kernel, one thread per (candidate, interval, direction)
    if (method == 2) {
        if (G == 1) return launch_one<DenseN<20>, 1, 2, 1>(L, ps, ldb, B, o, s);
        if (G == 0) return launch_one<DenseN<20>, 0, 2, 1>(L, ps, ldb, B, o, s);
        return -1000;
    }
    // with sensitivities: the tensor-path kernel (shoot_dense_mma.cuh); without, the generic kernel integrates x
    if (G > 0) return launch_dense_mma<20, 16>(method, L, ps, ldb, B, o, L.R * G, L.max_M, s);
    switch (G) {
        CB200_DISPATCH_G(DenseN<20>, 0, 1)
    default:
        return -1000;
    }
}
int dense20_set_data(const double *host_data, long long n)
{
    if (n != 20 * 20 + 20) return -1;
    return (int)cudaMemcpyToSymbol(c_dense, host_data, sizeof(double) * (20 * 20 + 20));
}
int forward_dense20(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<DenseN<20>, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
