// Linearizers: MERL half/difference grid (direction pair <-> bin index) and the regular spherical
// grid, fused into the kernels.
//
// Restates include/linearizer/merl_linearizer.h:34-37,50-83,94-123 and
// include/linearizer/spherical_linearizer.h:57,71-111 operation by operation in FP32/FP64 exactly
// as the native backbone evaluates them.
//
// Bit-exact bin indices need the host libm's float atan2f / sinf / cosf results.
// The reference is judged on glibc 2.39 (x86-64), so this file carries device restatements of the
// algorithms glibc 2.39 uses for those three functions:
//   * atan2f / atanf : the fdlibm single-precision routines (sysdeps/ieee754/flt-32/e_atan2f.c,
//     s_atanf.c) - pure float arithmetic with the published 11-term polynomial and hi/lo tables;
//   * sinf / cosf    : the ARM optimized-routines versions glibc adopted in 2.28
//     (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf_data.c) - double-precision quadrant
//     reduction and polynomials.
// Both were checked here bit-for-bit against the host glibc on 5e7 random arguments each
// (tests/test_libm_port.py repeats that through the host-compiled copy of this header).
// The library is compiled with -fmad=false, so none of the float expressions below is contracted.
#pragma once
#include "bbmcu_math.cuh"

namespace bbmcu {

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
BBMCU_D float glibc_sinf(float y) { return glibc_sincosf<false>(y); }
BBMCU_D float glibc_cosf(float y) { return glibc_sincosf<true>(y); }

// spherical::phi / theta with the host libm's atan2f (core/spherical.h:26-46)
BBMCU_D float lin_phi(f3 v) { float r = glibc_atan2f(v.y, v.x); return r < 0.0f ? r + kTwoPi : r; }
BBMCU_D float lin_theta(f3 v) { return sph_theta(v); }

BBMCU_D f3 rot_z(float c, float s, f3 v)   // rotationZ(cossin) * v: rows (c,-s,0), (s,c,0), (0,0,1)  (transform.h:74-81, mat.h:107-116)
{ return make_f3((0.0f + c*v.x) + (-s)*v.y + 0.0f*v.z, (0.0f + s*v.x) + c*v.y + 0.0f*v.z, (0.0f + 0.0f*v.x) + 0.0f*v.y + 1.0f*v.z); }
BBMCU_D f3 rot_y(float c, float s, f3 v)   // rotationY(cossin) * v: rows (c,0,s), (0,1,0), (-s,0,c)  (transform.h:47-54)
{ return make_f3((0.0f + c*v.x) + 0.0f*v.y + s*v.z, (0.0f + 0.0f*v.x) + 1.0f*v.y + 0.0f*v.z, (0.0f + (-s)*v.x) + 0.0f*v.y + c*v.z); }

constexpr uint32_t kMerlBins = 90u*90u*180u;

// merl_linearizer::operator()(in, out) -> bin index, kMerlBins when masked (merl_linearizer.h:94-123).
// NaN directions (antipodal grazing pairs) give an out-of-range value, reported as 0xFFFFFFFF.
BBMCU_D uint32_t merl_index(f3 in, f3 out)
{
  if(!((in.z >= 0.0f) && (out.z >= 0.0f))) return kMerlBins;
  f3 h = halfway(in, out);
  float ph = lin_phi(h), th = lin_theta(h);
  f3 t = rot_z(glibc_cosf(-ph), glibc_sinf(-ph), in);
  f3 d = rot_y(glibc_cosf(-th), glibc_sinf(-th), t);
  float pd = lin_phi(d), td = lin_theta(d);
  if(dot(in, out) > (float)(1.0 - (double)kEps)) pd = 0.0f;     // 1 - Epsilon() is exactly representable
  if(pd >= kPi) pd = pd - kPi;
  float iDp = floorf((pd / kPi + kEps) * 180.0f);
  float iDt = floorf((td / kHalfPi + kEps) * 90.0f);
  float iHt = floorf(m_safe_sqrt(th / kHalfPi + kEps) * 90.0f);
  // phi_h index: floor(safe_sqrt(ph / pi + eps) * 1) clamped to [0, 0]
  iDp = clampf(iDp, 0.0f, 179.0f);        // std::clamp: NaN stays NaN
  iDt = clampf(iDt, 0.0f, 89.0f);
  iHt = clampf(iHt, 0.0f, 89.0f);
  float idx = (iHt*90.0f + iDt)*180.0f + iDp;
  if(!(idx == idx)) return 0xFFFFFFFFu;
  return (uint32_t)idx;
}

// merl_linearizer::operator()(idx) -> (in, out) (merl_linearizer.h:50-83, with phi/theta of the
// half vector re-derived from the vector as core/vec_transform.h:117-123 does)
BBMCU_D void merl_dirs(uint32_t idx, f3& in, f3& out)
{
  if(!(idx < kMerlBins)) { in = make_f3(0, 0, 0); out = make_f3(0, 0, 0); return; }
  uint32_t iDp = idx % 180u, iDt = (idx / 180u) % 90u, iHt = idx / 16200u;
  float qh = (float)iHt / 90.0f;
  // pow(float2, 2.0) * (0.5 * Sphere()): the product runs in double and is rounded once
  float thh = (float)(((double)qh * (double)qh) * (double)(0.5f*kPi));
  float phd = ((float)iDp / 180.0f) * (0.5f*kTwoPi);
  float thd = ((float)iDt / 90.0f) * (0.5f*kPi);
  // spherical::convert(phi = 0, theta): cos(0) = 1, sin(0) = 0
  float sth = glibc_sinf(thh), cth = glibc_cosf(thh);
  f3 half = make_f3(1.0f*sth, 0.0f*sth, cth);
  float std_ = glibc_sinf(thd), ctd = glibc_cosf(thd);
  f3 diff = make_f3(glibc_cosf(phd)*std_, glibc_sinf(phd)*std_, ctd);
  float ph = lin_phi(half), th = lin_theta(half);
  float cz = glibc_cosf(ph), sz = glibc_sinf(ph), cy = glibc_cosf(th), sy = glibc_sinf(th);
  in = rot_z(cz, sz, rot_y(cy, sy, diff));
  out = rot_z(cz, sz, rot_y(cy, sy, make_f3(-diff.x, -diff.y, diff.z)));
  in.z = fmaxf(in.z, 0.0f);
  out.z = fmaxf(out.z, 0.0f);
}

// spherical_linearizer (spherical_linearizer.h:37-111)
struct SphericalGrid
{
  uint32_t n_in_phi, n_in_theta, n_out_phi, n_out_theta;
  float start_in_phi, start_in_theta, start_out_phi, start_out_theta;
  float size_in_phi, size_in_theta, size_out_phi, size_out_theta;     // end - start, formed in float
  BBMCU_HD uint64_t size() const { return (uint64_t)n_in_phi * n_in_theta * n_out_phi * n_out_theta; }
};

BBMCU_D f3 snap_eps(f3 v) { return make_f3(fabsf(v.x) < kEps ? 0.0f : v.x, fabsf(v.y) < kEps ? 0.0f : v.y, fabsf(v.z) < kEps ? 0.0f : v.z); }

BBMCU_D void spherical_dirs(const SphericalGrid& g, uint64_t idx, f3& in, f3& out)
{
  if(!(idx < g.size())) { in = make_f3(0, 0, 0); out = make_f3(0, 0, 0); return; }
  uint64_t t = idx;
  uint32_t ot = (uint32_t)(t % g.n_out_theta); t /= g.n_out_theta;
  uint32_t op = (uint32_t)(t % g.n_out_phi);   t /= g.n_out_phi;
  uint32_t it = (uint32_t)(t % g.n_in_theta);  t /= g.n_in_theta;
  uint32_t ip = (uint32_t)t;
  // theta includes both end points (/(n-1), at least 1); phi does not (/n)
  float s_it = (float)(g.n_in_theta > 1 ? g.n_in_theta - 1 : 1), s_ot = (float)(g.n_out_theta > 1 ? g.n_out_theta - 1 : 1);
  float in_phi = (float)ip * g.size_in_phi / (float)g.n_in_phi + g.start_in_phi;
  float in_theta = (float)it * g.size_in_theta / s_it + g.start_in_theta;
  float out_phi = (float)op * g.size_out_phi / (float)g.n_out_phi + g.start_out_phi;
  float out_theta = (float)ot * g.size_out_theta / s_ot + g.start_out_theta;
  float st = glibc_sinf(in_theta);
  in = snap_eps(make_f3(glibc_cosf(in_phi)*st, glibc_sinf(in_phi)*st, glibc_cosf(in_theta)));
  st = glibc_sinf(out_theta);
  out = snap_eps(make_f3(glibc_cosf(out_phi)*st, glibc_sinf(out_phi)*st, glibc_cosf(out_theta)));
}

} // namespace bbmcu
