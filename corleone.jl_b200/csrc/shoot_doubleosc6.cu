// Instantiations of the shooting kernel for model DoubleOsc6 (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
namespace cb200 {
int launch_doubleosc6(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    switch (G) {
        CB200_DISPATCH_GA(DoubleOsc6, 0, 1)
        CB200_DISPATCH_GA(DoubleOsc6, 1, 1)
    default:
        return -1000;
    }
}
int forward_doubleosc6(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<DoubleOsc6, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
