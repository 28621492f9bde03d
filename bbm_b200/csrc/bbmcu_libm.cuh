// Device restatements of the three float libm functions the reference's results depend on bit for bit
// (SURVEY.md fact 12): glibc 2.39's atan2f / atanf (fdlibm, sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c) and
// sinf / cosf (ARM optimized routines, sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf_data.c), written from
// the published algorithms.  Checked bit-for-bit against the host glibc on 5e7 random arguments each
// (tests/test_linearizer_loss_hostsim.py::test_glibc_trig_ports_match_host_libm repeats a 2e6 sample).
// Every product is rounded before it is added (the library is built with -fmad=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace bbmcu {

#ifndef BBMCU_D
#define BBMCU_D __host__ __device__ __forceinline__
#endif

BBMCU_D uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
BBMCU_D float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

// ---- glibc 2.39 atanf / atan2f (fdlibm) ---------------------------------------------------------
BBMCU_D float glibc_atanf(float x)
{
  const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
  const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
  const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f,
              a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f,
              a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
  int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff;
  int id;
  if(ix >= 0x4c000000) {
    if(ix > 0x7f800000) return x + x;
    return (hx > 0) ? hi3 + lo3 : -hi3 - lo3;
  }
  if(ix < 0x3ee00000) { if(ix < 0x31000000) return x; id = -1; }
  else {
    x = fabsf(x);
    if(ix < 0x3f980000) {
      if(ix < 0x3f300000) { id = 0; x = (2.0f*x - 1.0f) / (2.0f + x); }
      else                { id = 1; x = (x - 1.0f) / (x + 1.0f); }
    } else {
      if(ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f*x); }
      else                { id = 3; x = -1.0f / x; }
    }
  }
  float z = x*x, w = z*z;
  float s1 = z*(a0 + w*(a2 + w*(a4 + w*(a6 + w*(a8 + w*a10)))));
  float s2 = w*(a1 + w*(a3 + w*(a5 + w*(a7 + w*a9))));
  if(id < 0) return x - x*(s1 + s2);
  float hi = id == 0 ? hi0 : id == 1 ? hi1 : id == 2 ? hi2 : hi3;
  float lo = id == 0 ? lo0 : id == 1 ? lo1 : id == 2 ? lo2 : lo3;
  z = hi - ((x*(s1 + s2) - lo) - x);
  return (hx < 0) ? -z : z;
}

BBMCU_D float glibc_atan2f(float y, float x)
{
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff, hy = (int32_t)f2u(y), iy = hy & 0x7fffffff;
  if(ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if(hx == 0x3f800000) return glibc_atanf(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if(iy == 0) { switch(m) { case 0: case 1: return y; case 2: return pi + tiny; default: return -pi - tiny; } }
  if(ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if(ix == 0x7f800000) {
    if(iy == 0x7f800000) { switch(m) { case 0: return pi_o_4 + tiny; case 1: return -pi_o_4 - tiny; case 2: return 3.0f*pi_o_4 + tiny; default: return -3.0f*pi_o_4 - tiny; } }
    else { switch(m) { case 0: return 0.0f; case 1: return -0.0f; case 2: return pi + tiny; default: return -pi - tiny; } }
  }
  if(iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  int k = (iy - ix) >> 23;
  float z;
  if(k > 60) z = pi_o_2 + 0.5f*pi_lo;
  else if(hx < 0 && k < -60) z = 0.0f;
  else z = glibc_atanf(fabsf(y / x));
  switch(m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}

// ---- glibc 2.39 sinf / cosf (ARM optimized routines), |x| < 120 -----------------------------------
BBMCU_D float glibc_sincos_poly(double x, double x2, bool neg_cos, int n)
{
  const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  if((n & 1) == 0) {
    double x3 = x*x2, t = s2 + x2*s3, x7 = x3*x2, s = x + x3*s1;
    return (float)(s + x7*t);
  } else {
    double sg = neg_cos ? -1.0 : 1.0;
    double x4 = x2*x2, t2 = sg*c3 + x2*(sg*c4), t1 = sg*c0 + x2*(sg*c1), x6 = x4*x2, c = t1 + x4*(sg*c2);
    return (float)(c + x6*t2);
  }
}
template<bool COS> BBMCU_D float glibc_sincosf(float y)
{
  double x = y;
  uint32_t top = (f2u(y) >> 20) & 0x7ff;
  if(top < 0x3f4) {                                  // |y| < pi/4   (abstop12(0x1.921FB6p-1f) = 0x3f4)
    if(top < 0x398) return COS ? 1.0f : y;           // |y| < 2^-12
    return glibc_sincos_poly(x, x*x, false, COS ? 1 : 0);
  }
  if(top < 0x42f) {                                  // |y| < 120
    double r = x * 0x1.45F306DC9C883p+23;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - (double)n * 0x1.921FB54442D18p0;
    double sg = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return glibc_sincos_poly(x*sg, x*x, (n & 2) != 0, COS ? (n ^ 1) : n);
  }
  return COS ? cosf(y) : sinf(y);                    // never reached by the linearizers (angles in [-2pi, 2pi])
}
// sinf and cosf of the same argument with one quadrant reduction; each result is bit-identical to the separate calls.
// Round-to-nearest is symmetric under negation, so "polynomial with negated constants" (glibc's second table) and
// "odd polynomial of a negated argument" are exactly the negated plain polynomials: the signs are applied at the end.
BBMCU_D void glibc_sincosf_both(float y, float& sn, float& cs)
{
  const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  double x = y;
  const float ay = fabsf(y);
  if(!(ay < 120.0f)) { sn = sinf(y); cs = cosf(y); return; }          // (abstop12 < 0x42f in glibc; never taken by the samplers)
  // glibc skips the reduction below pi/4; the reduction then yields n = 0 and x - 0 * (pi/2) = x exactly, so one path
  // serves both.  Below 2^-12 it returns (y, 1) without the polynomials.
  double r = x * 0x1.45F306DC9C883p+23;
  const int n = ((int32_t)r + 0x800000) >> 24;
  x = x - (double)n * 0x1.921FB54442D18p0;
  double x2 = x*x;
  double x3 = x*x2, ts = s2 + x2*s3, x7 = x3*x2, ss = x + x3*s1;
  float A = (float)(ss + x7*ts);                                  // sin polynomial of the reduced argument
  double x4 = x2*x2, t2 = c3 + x2*c4, t1 = c0 + x2*c1, x6 = x4*x2, cc = t1 + x4*c2;
  float B = (float)(cc + x6*t2);                                  // cos polynomial
  const bool flip_a = ((n & 3) == 1) || ((n & 3) == 2);           // sign[n & 3] = {+, -, -, +}
  const bool flip_b = (n & 2) != 0;                               // the negated-constant table
  if(flip_a) A = -A;
  if(flip_b) B = -B;
  if(n & 1) { sn = B; cs = A; } else { sn = A; cs = B; }
  if(ay < 0x1p-12f) { sn = y; cs = 1.0f; }
}
BBMCU_D float glibc_sinf(float y) { return glibc_sincosf<false>(y); }
BBMCU_D float glibc_cosf(float y) { return glibc_sincosf<true>(y); }

} // namespace bbmcu
