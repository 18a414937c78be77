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
}  // namespace ts5
}  // namespace cb200
