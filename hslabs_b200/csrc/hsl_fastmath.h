// Branch-free FP64 primitives for the frame kernels.
//
// The CUDA library versions of a/b, sqrt and atan2 expand to a MUFU seed + Newton steps followed by a range check
// that branches to (or CALLs) a slow path.  The frame kernels are latency bound (a handful of warps per SM
// sub-partition, each running long dependent FP64 chains), so what matters is that independent divisions / square
// roots / arctangents of one thread can be interleaved by the scheduler -- which the slow-path branches prevent.
// These versions are straight-line code: hardware seed (rcp.approx / rsqrt.approx, 2^-23 relative error), then
// Newton / Goldschmidt steps in FMA arithmetic.  Arguments on the evaluation path are ordinary magnitudes
// (lengths ~1e-3..1e2, no denormals, no overflow), which is all these routines are specified for.
// Accuracy (measured against the IEEE operations, tests/test_gpu_math.py): div, sqrt <= 1 ulp; atan2 <= 2 ulp.
#pragma once
#include <math.h>

#if defined(__CUDA_ARCH__)
#define HSL_DEVICE_MATH 1
#else
#define HSL_DEVICE_MATH 0
#endif

#ifndef HSL_HD
#if defined(__CUDACC__)
#define HSL_HD __host__ __device__ __forceinline__
#else
#define HSL_HD inline
#endif
#endif

HSL_HD double hsl_rcp_seed(double b) {
#if HSL_DEVICE_MATH
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  return r;
#else
  return (double)(1.0f / (float)b);  // host emulation: same order of accuracy as the hardware seed
#endif
}
HSL_HD double hsl_rsqrt_seed(double x) {
#if HSL_DEVICE_MATH
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return (double)(1.0f / sqrtf((float)x));
#endif
}
// 1/b
HSL_HD double hsl_rcp(double b) {
  double r = hsl_rcp_seed(b);
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  return fma(r, e, r);
}
// a/b: reciprocal, one residual correction of the quotient
HSL_HD double hsl_div(double a, double b) {
  const double r = hsl_rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}
// sqrt(x), x >= 0 (x == 0 -> 0; x < 0 -> NaN)
HSL_HD double hsl_sqrt(double x) {
  const double y = hsl_rsqrt_seed(x);
  double g = x * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  g = fma(d, h, g);
  return (x == 0.0) ? 0.0 : g;
}

// 1/sqrt(x), x > 0: seed + two Newton steps y <- y + y*(1 - x y^2)/2  (2^-23 -> 2^-45 -> 2^-89), ~1 ulp
HSL_HD double hsl_rsqrt(double x) {
  double y = hsl_rsqrt_seed(x);
  double e = fma(-(x * y), y, 1.0);
  y = fma(y * 0.5, e, y);
  e = fma(-(x * y), y, 1.0);
  return fma(y * 0.5, e, y);
}

// sin and cos of x in [0, pi] (the argument range of the gait generator's stepx / stepz, pergen.cpp:62-71).
// Quadrant reduction r = x - k*pi/2 with a two-part pi/2 (k in {0,1,2}; k*PIO2_HI is exact), then the Taylor
// series of sin and cos on |r| <= pi/4 (9 terms each, truncation < 1e-19; coefficients are exact reciprocals of
// factorials, so there is nothing to mis-tabulate).  Quadrant fix-up by selects: no branches.
HSL_HD void hsl_sincos_0_pi(double x, double* sn, double* cs) {
  const double PIO2_HI = 1.57079632580280303955e+00;  // pi/2 with the low 24 mantissa bits cleared
  const double PIO2_LO = 9.92093579680540425193e-10;  // pi/2 - PIO2_HI (residual 1.6e-26)
  const double kf = floor(fma(x, 6.36619772367581382433e-01, 0.5));  // round(x * 2/pi)
  double r = fma(-kf, PIO2_HI, x);
  r = fma(-kf, PIO2_LO, r);
  const double z = r * r;
  double ps = 1.0 / 355687428096000.0;          // 1/17!
  ps = fma(ps, z, -1.0 / 1307674368000.0);      // 1/15!
  ps = fma(ps, z, 1.0 / 6227020800.0);          // 1/13!
  ps = fma(ps, z, -1.0 / 39916800.0);           // 1/11!
  ps = fma(ps, z, 1.0 / 362880.0);              // 1/9!
  ps = fma(ps, z, -1.0 / 5040.0);               // 1/7!
  ps = fma(ps, z, 1.0 / 120.0);                 // 1/5!
  ps = fma(ps, z, -1.0 / 6.0);                  // 1/3!
  const double sr = fma(r * z, ps, r);
  double pc = 1.0 / 20922789888000.0;           // 1/16!
  pc = fma(pc, z, -1.0 / 87178291200.0);        // 1/14!
  pc = fma(pc, z, 1.0 / 479001600.0);           // 1/12!
  pc = fma(pc, z, -1.0 / 3628800.0);            // 1/10!
  pc = fma(pc, z, 1.0 / 40320.0);               // 1/8!
  pc = fma(pc, z, -1.0 / 720.0);                // 1/6!
  pc = fma(pc, z, 1.0 / 24.0);                  // 1/4!
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  // k = 0: (sr, cr)   k = 1: (cr, -sr)   k = 2: (-sr, -cr)
  const bool k1 = (kf == 1.0), k2 = (kf >= 2.0);
  *sn = k1 ? cr : (k2 ? -sr : sr);
  *cs = k1 ? -sr : (k2 ? -cr : cr);
}

// sin and cos for |x| <= pi (turn angles of curved gaits); larger arguments take the library routine.
HSL_HD void hsl_sincos_pm_pi(double x, double* sn, double* cs) {
  const double ax = fabs(x);
  if (ax <= 3.2) {
    double s, c;
    hsl_sincos_0_pi(ax, &s, &c);
    *sn = (x < 0) ? -s : s;
    *cs = c;
  } else {
    sincos(x, sn, cs);
  }
}

// Angle in (-pi, pi) of the unit vector (cd, sd) = (cos D, sin D):  D = 2 atan(sd / (1 + cd)).
// Used for the wrapped joint-angle difference over two frames (periodic.cpp:271-278), which is small, so the
// half-angle tangent t is small and the Taylor series of atan (coefficients +-1/(2k+1), exact) converges fast.
// *slow is set when |t| > 1/4 (|D| > 0.49 rad): the caller then takes the library atan2 for that value.
HSL_HD double hsl_small_angle(double sd, double cd, bool* slow) {
  const double t = hsl_div(sd, 1.0 + cd);
  const double z = t * t;
  *slow = !(z <= 0.0625);
  double p = 1.0 / 27.0;
  p = fma(p, z, -1.0 / 25.0);
  p = fma(p, z, 1.0 / 23.0);
  p = fma(p, z, -1.0 / 21.0);
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, -1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, -1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, -1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, -1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  const double a = fma(-t * z, p, t);  // atan(t) = t - t^3 (1/3 - z/5 + ...)
  return a + a;
}
HSL_HD double hsl_atan2(double y, double x) {  // self-test helper: the composition used by the kernels
  const double rn = hsl_rcp(hsl_sqrt(x * x + y * y));
  bool slow;
  const double d = hsl_small_angle(y * rn, x * rn, &slow);
  return slow ? atan2(y, x) : d;
}
