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
// Precondition: |y| < 120 or NaN (every caller passes an angle in [-2 pi, 2 pi]); glibc's large-argument reduction is not
// restated.  The double polynomials are contracted to fused multiply-adds: the x86-64 libm the reference runs on selects
// its FMA build, and over 2e9 arguments in [-2 pi, 2 pi] the fused and the separately rounded evaluation never gave a
// different float anyway (the double rounding error is 2^-29 of a float ulp).
BBMCU_D void glibc_sincosf_both(float y, float& sn, float& cs)
{
  const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
  const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
  double x = y;
  const float ay = fabsf(y);
  // glibc skips the reduction below pi/4; the reduction then yields n = 0 and x - 0 * (pi/2) = x exactly, so one path
  // serves both.  Below 2^-12 it returns (y, 1) without the polynomials.
  double r = x * 0x1.45F306DC9C883p+23;
  const int n = ((int32_t)r + 0x800000) >> 24;
  x = fma(-(double)n, 0x1.921FB54442D18p0, x);
  double x2 = x*x;
  double x3 = x*x2, ts = fma(x2, s3, s2), x7 = x3*x2, ss = fma(x3, s1, x);
  float A = (float)fma(x7, ts, ss);                               // sin polynomial of the reduced argument
  double x4 = x2*x2, t2 = fma(x2, c4, c3), t1 = fma(x2, c1, c0), x6 = x4*x2, cc = fma(x4, c2, t1);
  float B = (float)fma(x6, t2, cc);                               // cos polynomial
  const bool flip_a = ((n & 3) == 1) || ((n & 3) == 2);           // sign[n & 3] = {+, -, -, +}
  const bool flip_b = (n & 2) != 0;                               // the negated-constant table
  if(flip_a) A = -A;
  if(flip_b) B = -B;
  if(n & 1) { sn = B; cs = A; } else { sn = A; cs = B; }
  if(ay < 0x1p-12f) { sn = y; cs = 1.0f; }
}
BBMCU_D float glibc_sinf(float y) { return glibc_sincosf<false>(y); }
BBMCU_D float glibc_cosf(float y) { return glibc_sincosf<true>(y); }


BBMCU_D uint64_t d2u(double f) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(f);
#else
  uint64_t u; memcpy(&u, &f, 8); return u;
#endif
}
BBMCU_D double u2d(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double f; memcpy(&f, &u, 8); return f;
#endif
}

// ---- glibc 2.39 expf / logf (ARM optimized routines: sysdeps/ieee754/flt-32/e_expf.c, e_logf.c with their
// exp2f_data / logf_data tables) and erff (fdlibm, sysdeps/ieee754/flt-32/s_erff.c), written from the published
// algorithms; the erff coefficients were cross-checked against the constants the host library carries.  expf follows
// the x86-64 multiarch FMA build that the dynamic loader selects on every FMA-capable host (one fused
// multiply-subtract in the reduction, fused polynomial steps).  Each agrees with the host glibc on 2e8 random
// arguments without a single differing bit (tests/test_linearizer_loss_hostsim.py repeats a sample).
// The Beckmann visible-normal sampler ends in erfinv near +-1, which amplifies a last-bit difference of any of these
// by up to 1e4: with them it reproduces the reference's sampled directions instead of 4e-5 of them being off.
// (tables: one copy in device memory and one for the host-compiled tests; a function-local array would live in
// local memory and be rebuilt on every call)
#define BBMCU_EXP2F_T { \
0x3ff0000000000000, 0x3fefd9b0d3158574, 0x3fefb5586cf9890f, 0x3fef9301d0125b51, 0x3fef72b83c7d517b, 0x3fef54873168b9aa, 0x3fef387a6e756238, 0x3fef1e9df51fdee1, \
0x3fef06fe0a31b715, 0x3feef1a7373aa9cb, 0x3feedea64c123422, 0x3feece086061892d, 0x3feebfdad5362a27, 0x3feeb42b569d4f82, 0x3feeab07dd485429, 0x3feea47eb03a5585, \
0x3feea09e667f3bcd, 0x3fee9f75e8ec5f74, 0x3feea11473eb0187, 0x3feea589994cce13, 0x3feeace5422aa0db, 0x3feeb737b0cdc5e5, 0x3feec49182a3f090, 0x3feed503b23e255d, \
0x3feee89f995ad3ad, 0x3feeff76f2fb5e47, 0x3fef199bdd85529c, 0x3fef3720dcef9069, 0x3fef5818dcfba487, 0x3fef7c97337b9b5f, 0x3fefa4afa2a490da, 0x3fefd0765b6e4540 }
#ifdef __CUDACC__
static __device__ const uint64_t d_exp2f_T[32] = BBMCU_EXP2F_T;
#endif
static const uint64_t h_exp2f_T[32] = BBMCU_EXP2F_T;
BBMCU_D uint64_t exp2f_T(int i)
{
#ifdef __CUDA_ARCH__
  return d_exp2f_T[i];
#else
  return h_exp2f_T[i];
#endif
}
BBMCU_D float glibc_expf(float x)
{
  const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5/32/32/32, C1 = 0x1.ebfce50fac4f3p-3/32/32, C2 = 0x1.62e42ff0c52d6p-1/32;
  double xd = (double)x;
  uint32_t abstop = (f2u(x) >> 20) & 0x7ff;
  if(abstop >= 0x42b) {                       /* |x| >= 88 or NaN/Inf */
    if(f2u(x) == f2u(u2f(0xff800000u))) return 0.0f;
    if(abstop >= 0x7f8) return x + x;
    if(x > 0x1.62e42ep6f) return u2f(0x7f800000u);   /* overflow */
    if(x < -0x1.9fe368p6f) return 0.0f;      /* underflow */
  }
  double z = InvLn2N * xd;
  double kd = z + SHIFT;
  uint64_t ki = d2u(kd);
  kd -= SHIFT;
  double r = fma(InvLn2N, xd, -kd);
  uint64_t t = exp2f_T((int)(ki % 32));
  t += ki << (52 - 5);
  double s = u2d(t);
  z = fma(C0, r, C1);
  double r2 = r * r;
  double y = fma(C2, r, 1.0);
  y = fma(z, r2, y);
  y = y * s;
  return (float)y;
}
#define BBMCU_LOGF_T { \
 {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2}, {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3}, \
 {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3}, {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4}, \
 {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0}, {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5}, {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4}, \
 {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3}, {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2}, {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2} }
#ifdef __CUDACC__
static __device__ const double d_logf_T[16][2] = BBMCU_LOGF_T;
#endif
static const double h_logf_T[16][2] = BBMCU_LOGF_T;
BBMCU_D void logf_T(int i, double& invc, double& logc)
{
#ifdef __CUDA_ARCH__
  invc = d_logf_T[i][0]; logc = d_logf_T[i][1];
#else
  invc = h_logf_T[i][0]; logc = h_logf_T[i][1];
#endif
}
BBMCU_D float glibc_logf(float x)
{
  const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2, Ln2 = 0x1.62e42fefa39efp-1;
  uint32_t ix = f2u(x);
  if(ix == 0x3f800000) return 0.0f;
  if(ix - 0x00800000 >= 0x7f800000 - 0x00800000) {
    if(ix * 2 == 0) return u2f(0xff800000u);
    if(ix == 0x7f800000) return x;
    if((ix & 0x80000000) || ix * 2 >= 0xff000000) return u2f(0x7fc00000u);          // __math_invalidf: NaN
    ix = f2u(x * 0x1p23f); ix -= 23 << 23;
  }
  uint32_t tmp = ix - 0x3f330000;
  int i = (tmp >> (23 - 4)) % 16;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000);
  double invc, logc; logf_T(i, invc, logc);
  double z = (double)u2f(iz);
  double r = z * invc - 1;
  double y0 = logc + (double)k * Ln2;
  double r2 = r * r;
  double y = A1 * r + A2;
  y = A0 * r2 + y;
  y = y * r2 + (y0 + r);
  return (float)y;
}
BBMCU_D float glibc_erff(float x)
{
  const float erx = 8.4506291151e-01f, efx = 1.2837916613e-01f, efx8 = 1.0270333290e+00f;
  uint32_t hx = f2u(x), ix = hx & 0x7fffffff;
  if(ix >= 0x7f800000) { int i = ((uint32_t)hx >> 31) << 1; return (float)(1 - i) + 1.0f / x; }
  if(ix < 0x3f580000) {
    if(ix < 0x31800000) { if(ix < 0x04000000) return 0.0625f * (16.0f * x + u2f(0x400375d4) * x); return x + u2f(0x3e0375d4) * x; }
    float z = x * x;
    float r = u2f(0x3e0375d4) + z * (-u2f(0x3ea66beb) + z * (-u2f(0x3ce9528f) + z * (-u2f(0x3bbd1489) + z * u2f(0xb7c756b1))));
    float s = 1.0f + z * (u2f(0x3ecbbbce) + z * (u2f(0x3d852a63) + z * (u2f(0x3ba68116) + z * (u2f(0x390aee49) + z * u2f(0xb684e21a)))));
    float y = r / s;
    return x + x * y;
  }
  if(ix < 0x3fa00000) {
    float s = fabsf(x) - 1.0f;
    float P = -u2f(0x3b1acdc6) + s * (u2f(0x3ed46805) + s * (-u2f(0x3ebe9208) + s * (u2f(0x3ea2fe54) + s * (-u2f(0x3de31cc2) + s * (u2f(0x3d1151b3) + s * u2f(0xbb0df9c0))))));
    float Q = 1.0f + s * (u2f(0x3dd9f331) + s * (u2f(0x3f0a5785) + s * (u2f(0x3d931ae7) + s * (u2f(0x3e013307) + s * (u2f(0x3c5f6e13) + s * u2f(0x3c445aa3))))));
    return ((int32_t)hx >= 0) ? erx + P / Q : -erx - P / Q;
  }
  if(ix >= 0x40c00000) return ((int32_t)hx >= 0) ? 1.0f - 1e-30f : 1e-30f - 1.0f;
  float ax = fabsf(x);
  float s = 1.0f / (ax * ax);
  float R, S;
  if(ix < 0x4036DB6E) {
    R = -u2f(0x3c21a093) + s * (-u2f(0x3f31a0b7) + s * (-u2f(0x4128f022) + s * (-u2f(0x42798057) + s * (-u2f(0x4322658c) + s * (-u2f(0x43389ae7) + s * (-u2f(0x42a2932b) + s * u2f(0xc11d077e)))))));
    S = 1.0f + s * (u2f(0x419d35ce) + s * (u2f(0x4309a863) + s * (u2f(0x43d9486f) + s * (u2f(0x442158c9) + s * (u2f(0x43d6810b) + s * (u2f(0x42d9451f) + s * (u2f(0x40d23f7c) + s * u2f(0xbd777f97))))))));
  } else {
    R = -u2f(0x3c21a092) + s * (-u2f(0x3f4c9dd4) + s * (-u2f(0x418e104b) + s * (-u2f(0x4320a2ea) + s * (-u2f(0x441f6441) + s * (-u2f(0x4480230b) + s * u2f(0xc3f1c275))))));
    S = 1.0f + s * (u2f(0x41f2b459) + s * (u2f(0x43a2e571) + s * (u2f(0x44c01759) + s * (u2f(0x4547fdbb) + s * (u2f(0x451f90ce) + s * (u2f(0x43ed43a7) + s * u2f(0xc1b38712)))))));
  }
  float z = u2f(f2u(ax) & 0xfffff000);
  float r = glibc_expf(-z * z - 0.5625f) * glibc_expf((z - ax) * (z + ax) + R / S);
  return ((int32_t)hx >= 0) ? 1.0f - r / ax : r / ax - 1.0f;
}

// ---- glibc 2.39 tanf (sysdeps/ieee754/flt-32/s_tanf.c, k_tanf.c) and erfcf (fdlibm, s_erff.c) ---------------------------
// Written from the published algorithms; both agree with the host glibc on EVERY float (erfcf: all 2^32 arguments;
// tanf: every |x| <= 100) without a single differing bit (tools/libm_sweep.cpp; tests/test_linearizer_loss_hostsim.py
// repeats a sample).  LowSmooth's sampler goes through tan and atan next to a cancelling E - 2, and the He family's
// sampling CDF is built from erfc: a last-bit difference in either moved 5e-5 (LowSmooth) and 5e-6 (He: a xi within one
// ulp of a CDF entry picks the neighbouring bin) of the sampled directions beyond the 1e-5 contract.
BBMCU_D float glibc_kernel_tanf(float x, float y, int iy)
{
  const float pio4 = 7.8539812565e-01f, pio4lo = 3.7748947079e-08f;
  const float T0 = u2f(0x3eaaaaab), T1 = u2f(0x3e088889), T2 = u2f(0x3d5d0dd1), T3 = u2f(0x3cb327a4), T4 = u2f(0x3c11371f), T5 = u2f(0x3b6b6916),
              T6 = u2f(0x3abede48), T7 = u2f(0x3a1a26c8), T8 = u2f(0x398137b9), T9 = u2f(0x38a3f445), T10 = u2f(0x3895c07a), T11 = u2f(0xb79bae5f), T12 = u2f(0x37d95384);
  int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff;
  if(ix < 0x39000000) {                       /* |x| < 2**-13 */
    if((int)x == 0) {
      if((ix | (iy + 1)) == 0) return 1.0f / fabsf(x);
      else if(iy == 1) return x;
      else return -1.0f / x;
    }
  }
  if(ix >= 0x3f2ca140) {                      /* |x| >= 0.6744 */
    if(hx < 0) { x = -x; y = -y; }
    float z = pio4 - x, w = pio4lo - y;
    x = z + w; y = 0.0f;
    if(fabsf(x) < 0x1p-13f) return (float)(1 - ((hx >> 30) & 2)) * (float)iy * (1.0f - 2.0f * (float)iy * x);
  }
  float z = x*x, w = z*z;
  float r = T1 + w*(T3 + w*(T5 + w*(T7 + w*(T9 + w*T11))));
  float v = z*(T2 + w*(T4 + w*(T6 + w*(T8 + w*(T10 + w*T12)))));
  float s = z*x;
  r = y + z*(s*(r + v) + y);
  r += T0*s;
  w = x + r;
  if(ix >= 0x3f2ca140) {
    v = (float)iy;
    return (float)(1 - ((hx >> 30) & 2)) * (v - 2.0f*(x - (w*w/(w + v) - r)));
  }
  if(iy == 1) return w;
  /* compute -1.0/(x+r) accurately */
  float a, t;
  z = u2f(f2u(w) & 0xfffff000u);
  v = r - (z - x);
  t = a = -1.0f / w;
  t = u2f(f2u(t) & 0xfffff000u);
  s = 1.0f + t*z;
  return t + a*(s + t*v);
}
// glibc 2.39 tanf (sysdeps/ieee754/flt-32/s_tanf.c): |x| <= pi/4 goes straight to the kernel; otherwise the double-precision
// quadrant reduction of sinf/cosf (n = round(x * 2/pi), x - n * pi/2 in double), the reduced argument split into a float
// head and tail, and the fdlibm kernel with iy = +1 (n even) or -1 (n odd).  |x| < 120 (every caller passes an angle of a
// few turns); the large-argument reduction is not restated.
BBMCU_D float glibc_tanf(float x)
{
  int32_t ix = (int32_t)f2u(x) & 0x7fffffff;
  if(ix <= 0x3f490fda) return glibc_kernel_tanf(x, 0.0f, 1);
  if(ix >= 0x7f800000) return x - x;
  double xd = (double)x;
  double r = xd * 0x1.45F306DC9C883p+23;
  int n = ((int32_t)r + 0x800000) >> 24;
  xd = xd - (double)n * 0x1.921FB54442D18p0;
  float y0 = (float)xd;
  float y1 = (float)(xd - (double)y0);
  return glibc_kernel_tanf(y0, y1, 1 - ((n & 1) << 1));
}
BBMCU_D float glibc_erfcf(float x)
{
  const float erx = 8.4506291151e-01f;
  int32_t hx = (int32_t)f2u(x), ix = hx & 0x7fffffff;
  if(ix >= 0x7f800000) return (float)(((uint32_t)hx >> 31) << 1) + 1.0f / x;
  if(ix < 0x3f580000) {                       /* |x| < 0.84375 */
    if(ix < 0x23800000) return 1.0f - x;
    float z = x*x;
    float r = u2f(0x3e0375d4) + z * (-u2f(0x3ea66beb) + z * (-u2f(0x3ce9528f) + z * (-u2f(0x3bbd1489) + z * u2f(0xb7c756b1))));
    float s = 1.0f + z * (u2f(0x3ecbbbce) + z * (u2f(0x3d852a63) + z * (u2f(0x3ba68116) + z * (u2f(0x390aee49) + z * u2f(0xb684e21a)))));
    float y = r / s;
    if(hx < 0x3e800000) return 1.0f - (x + x*y);
    r = x*y; r += (x - 0.5f);
    return 0.5f - r;
  }
  if(ix < 0x3fa00000) {                       /* 0.84375 <= |x| < 1.25 */
    float s = fabsf(x) - 1.0f;
    float P = -u2f(0x3b1acdc6) + s * (u2f(0x3ed46805) + s * (-u2f(0x3ebe9208) + s * (u2f(0x3ea2fe54) + s * (-u2f(0x3de31cc2) + s * (u2f(0x3d1151b3) + s * u2f(0xbb0df9c0))))));
    float Q = 1.0f + s * (u2f(0x3dd9f331) + s * (u2f(0x3f0a5785) + s * (u2f(0x3d931ae7) + s * (u2f(0x3e013307) + s * (u2f(0x3c5f6e13) + s * u2f(0x3c445aa3))))));
    if(hx >= 0) { float z = 1.0f - erx; return z - P / Q; }
    float z = erx + P / Q; return 1.0f + z;
  }
  if(ix < 0x41e00000) {                       /* |x| < 28 */
    float ax = fabsf(x);
    float s = 1.0f / (ax * ax);
    float R, S;
    if(ix < 0x4036DB6D) {
      R = -u2f(0x3c21a093) + s * (-u2f(0x3f31a0b7) + s * (-u2f(0x4128f022) + s * (-u2f(0x42798057) + s * (-u2f(0x4322658c) + s * (-u2f(0x43389ae7) + s * (-u2f(0x42a2932b) + s * u2f(0xc11d077e)))))));
      S = 1.0f + s * (u2f(0x419d35ce) + s * (u2f(0x4309a863) + s * (u2f(0x43d9486f) + s * (u2f(0x442158c9) + s * (u2f(0x43d6810b) + s * (u2f(0x42d9451f) + s * (u2f(0x40d23f7c) + s * u2f(0xbd777f97))))))));
    } else {
      if(hx < 0 && ix >= 0x40c00000) return 2.0f - 1e-30f;
      R = -u2f(0x3c21a092) + s * (-u2f(0x3f4c9dd4) + s * (-u2f(0x418e104b) + s * (-u2f(0x4320a2ea) + s * (-u2f(0x441f6441) + s * (-u2f(0x4480230b) + s * u2f(0xc3f1c275))))));
      S = 1.0f + s * (u2f(0x41f2b459) + s * (u2f(0x43a2e571) + s * (u2f(0x44c01759) + s * (u2f(0x4547fdbb) + s * (u2f(0x451f90ce) + s * (u2f(0x43ed43a7) + s * u2f(0xc1b38712)))))));
    }
    float z = u2f(f2u(ax) & 0xffffe000u);
    float r = glibc_expf(-z * z - 0.5625f) * glibc_expf((z - ax) * (z + ax) + R / S);
    if(hx > 0) return r / ax;
    return 2.0f - r / ax;
  }
  if(hx > 0) return 1e-30f * 1e-30f;
  return 2.0f - 1e-30f;
}

} // namespace bbmcu
