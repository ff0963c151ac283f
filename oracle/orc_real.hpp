// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Scalar type of the oracle.  The default build is double, like the reference.  Building with -DORC_QUAD evaluates
// the very same formulas in __float128 (libquadmath); comparing the two bounds the conditioning of the reference
// algorithm itself, i.e. how much of a GPU-vs-oracle difference is mere round-off of the FP64 reference
// (pin 8).  Constants such as M_PI stay the double values the reference uses.
#pragma once
#include <cmath>
#include <istream>
#ifdef ORC_QUAD
#include <quadmath.h>
namespace orc {
typedef __float128 real;
inline real m_sqrt(real x) { return sqrtq(x); }
inline real m_sin(real x) { return sinq(x); }
inline real m_cos(real x) { return cosq(x); }
inline real m_acos(real x) { return acosq(x); }
inline real m_asin(real x) { return asinq(x); }
inline real m_atan2(real y, real x) { return atan2q(y, x); }
inline real m_fabs(real x) { return fabsq(x); }
}  // namespace orc
#else
namespace orc {
typedef double real;
inline real m_sqrt(real x) { return std::sqrt(x); }
inline real m_sin(real x) { return std::sin(x); }
inline real m_cos(real x) { return std::cos(x); }
inline real m_acos(real x) { return std::acos(x); }
inline real m_asin(real x) { return std::asin(x); }
inline real m_atan2(real y, real x) { return std::atan2(y, x); }
inline real m_fabs(real x) { return std::fabs(x); }
}  // namespace orc
#endif
namespace orc {
// Text input always goes through double (the reference reads doubles: core.cpp:8-12, player.cpp:170-208).
struct RealIn { real& r; explicit RealIn(real& r_) : r(r_) {} };
inline std::istream& operator>>(std::istream& is, RealIn w) { double d; if (is >> d) w.r = d; return is; }
}  // namespace orc
