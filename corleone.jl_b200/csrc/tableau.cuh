// Runge-Kutta tableaux used by the shooting kernels.
// Tsit5: Tsitouras 2011 5(4) pair, propagation weights only (fixed step; the reference runs
// OrdinaryDiffEqTsit5 with adaptive=false, dt=h -- SURVEY.md Appendix A).  RK4: classical.
#pragma once

namespace cb200 {
namespace ts5 {
constexpr double a21 = 0.161;
constexpr double a31 = -0.008480655492356989, a32 = 0.335480655492357;
constexpr double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
constexpr double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365,
                 a54 = -0.09249506636175525;
constexpr double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159,
                 a64 = -0.071584973281401, a65 = -0.028269050394068383;
constexpr double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996, a74 = 1.379008574103742,
                 a75 = -3.290069515436081, a76 = 2.324710524099774;
// embedded error estimator (btilde) of the 5(4) pair -- method 2, adaptive stepping
constexpr double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469, bt3 = 0.007880878010261995,
                 bt4 = -0.1447110071732629, bt5 = 0.5823571654525552, bt6 = -0.45808210592918697,
                 bt7 = 0.015151515151515152;
}  // namespace ts5

// Tsit5 coefficient slots: cf[0]=h a21; [1..2]=h a31,a32; [3..5]=h a4.; [6..9]=h a5.; [10..14]=h a6.; [15..20]=h a7.
__host__ __device__ inline void tsit5_coeffs(double h, double *cf)
{
    using namespace ts5;
    const double a[21] = {a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65,
                          a71, a72, a73, a74, a75, a76};
    for (int k = 0; k < 21; k++) cf[k] = h * a[k];
}
// RK4 coefficient slots: cf[0]=h/2, cf[1]=h, cf[2]=h/6, cf[3]=h/3
__host__ __device__ inline void rk4_coeffs(double h, double *cf)
{
    cf[0] = 0.5 * h;
    cf[1] = h;
    cf[2] = h / 6.0;
    cf[3] = h / 3.0;
}

}  // namespace cb200
