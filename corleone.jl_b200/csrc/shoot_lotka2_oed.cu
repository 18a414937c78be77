// Instantiations of the shooting kernel for model Lotka2Oed (see shoot.cuh).
#include "models.cuh"
#include "shoot.cuh"
#include <cstdlib>
namespace cb200 {
int launch_lotka2_oed(int G, int method, const DevLayer &L, const double *ps, long long ldb, long long B,
        const ShootOut &o, cudaStream_t s)
{
    if (G == 0 && method == 0) {   // tuning: CB200_OED_MINB=4 (125 registers, no spills, 16 warps/SM) / 6 (80 registers)
        static const int mb = getenv("CB200_OED_MINB") ? atoi(getenv("CB200_OED_MINB")) : 5;
        if (mb == 4) return launch_one<Lotka2Oed, 0, 0, 4>(L, ps, ldb, B, o, s);
        if (mb == 6) return launch_one<Lotka2Oed, 0, 0, 6>(L, ps, ldb, B, o, s);
    }
    switch (G) {
        CB200_DISPATCH_GA(Lotka2Oed, 0, 5)  // 96 registers: 20 warps/SM; G x MINB sweep in profiles/
        CB200_DISPATCH_GA(Lotka2Oed, 1, 1)
    default:
        return -1000;
    }
}
int forward_lotka2_oed(int method, const DevLayer &L, double *ps, long long ldb, long long B, const unsigned char *fixed,
        cudaStream_t s)
{
    return launch_forward<Lotka2Oed, true>(method, L, ps, ldb, B, fixed, s);
}
}  // namespace cb200
