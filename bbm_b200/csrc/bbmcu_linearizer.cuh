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
#include "bbmcu_libm.cuh"

namespace bbmcu {

// spherical::phi / theta with the host libm's atan2f (core/spherical.h:26-46)
BBMCU_D float lin_phi(f3 v) { float r = glibc_atan2f(v.y, v.x); return r < 0.0f ? r + kTwoPi : r; }
BBMCU_D float lin_theta(f3 v) { return sph_theta(v); }

// products and sums rounded one by one whatever the translation unit's -fmad setting: these run inside the loss kernels
// (built -fmad=true) as well and must return the bits of the -fmad=false build there too
BBMCU_D float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
BBMCU_D float add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
BBMCU_D float dot3_rn(float a0, float a1, float a2, f3 v) { return add_rn(add_rn(add_rn(0.0f, mul_rn(a0, v.x)), mul_rn(a1, v.y)), mul_rn(a2, v.z)); }   // ((0 + a0 x) + a1 y) + a2 z  (horizontal.h:87-91)
BBMCU_D f3 rot_z(float c, float s, f3 v)   // rotationZ(cossin) * v: rows (c,-s,0), (s,c,0), (0,0,1)  (transform.h:74-81, mat.h:107-116)
{ return make_f3(dot3_rn(c, -s, 0.0f, v), dot3_rn(s, c, 0.0f, v), dot3_rn(0.0f, 0.0f, 1.0f, v)); }
BBMCU_D f3 rot_y(float c, float s, f3 v)   // rotationY(cossin) * v: rows (c,0,s), (0,1,0), (-s,0,c)  (transform.h:47-54)
{ return make_f3(dot3_rn(c, 0.0f, s, v), dot3_rn(0.0f, 1.0f, 0.0f, v), dot3_rn(-s, 0.0f, c, v)); }

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
// half vector re-derived from the vector as core/vec_transform.h:117-123 does).
//
// The computation is separable: everything transcendental depends on ONE of the three bin coordinates
//   iHt -> the two rotations  (cos/sin of theta(half), cos/sin of phi(half))        merl_lin_half
//   iDt -> sin/cos of theta_d                                                       merl_lin_dtheta
//   iDp -> cos/sin of phi_d                                                         merl_lin_dphi
// and the rest (merl_dirs_assemble) is ~40 float operations.  merl_dirs() evaluates the three pieces per bin;
// merl_dirs_tab() reads them from a 900-float table built once per device with the very same functions, so both
// return identical bits - this is what lets eval and loss kernels generate the grid's directions in registers
// instead of reading 24 B per sample ("fused linearizer").
struct MerlLinHalf { float cy, sy, cz, sz; };
BBMCU_D MerlLinHalf merl_lin_half(uint32_t iHt)
{
  float qh = (float)iHt / 90.0f;
  // pow(float2, 2.0) * (0.5 * Sphere()): the product runs in double and is rounded once
  float thh = (float)(((double)qh * (double)qh) * (double)(0.5f*kPi));
  // spherical::convert(phi = 0, theta): cos(0) = 1, sin(0) = 0
  float sth = glibc_sinf(thh), cth = glibc_cosf(thh);
  f3 half = make_f3(1.0f*sth, 0.0f*sth, cth);
  float ph = lin_phi(half), th = lin_theta(half);
  MerlLinHalf r; r.cz = glibc_cosf(ph); r.sz = glibc_sinf(ph); r.cy = glibc_cosf(th); r.sy = glibc_sinf(th);
  return r;
}
BBMCU_D void merl_lin_dtheta(uint32_t iDt, float& s, float& c) { float thd = ((float)iDt / 90.0f) * (0.5f*kPi); s = glibc_sinf(thd); c = glibc_cosf(thd); }
BBMCU_D void merl_lin_dphi(uint32_t iDp, float& c, float& s) { float phd = ((float)iDp / 180.0f) * (0.5f*kTwoPi); c = glibc_cosf(phd); s = glibc_sinf(phd); }
BBMCU_D void merl_dirs_assemble(const MerlLinHalf& h, float std_, float ctd, float cpd, float spd, f3& in, f3& out)
{
  f3 diff = make_f3(mul_rn(cpd, std_), mul_rn(spd, std_), ctd);
  in = rot_y(h.cy, h.sy, diff);
  out = rot_y(h.cy, h.sy, make_f3(-diff.x, -diff.y, diff.z));
  if(h.cz == 1.0f && h.sz == 0.0f)
  {
    // phi(half) is exactly 0 for every bin of this grid (half = (sin, 0, cos)), so rotationZ is the identity matrix; its
    // row sums ((0 + 1 x) + (-0) y) + 0 z leave every finite x unchanged except that -0 becomes +0: x + 0 does the same
    in = make_f3(add_rn(in.x, 0.0f), add_rn(in.y, 0.0f), add_rn(in.z, 0.0f));
    out = make_f3(add_rn(out.x, 0.0f), add_rn(out.y, 0.0f), add_rn(out.z, 0.0f));
  }
  else { in = rot_z(h.cz, h.sz, in); out = rot_z(h.cz, h.sz, out); }
  in.z = fmaxf(in.z, 0.0f);
  out.z = fmaxf(out.z, 0.0f);
}
BBMCU_D void merl_dirs(uint32_t idx, f3& in, f3& out)
{
  if(!(idx < kMerlBins)) { in = make_f3(0, 0, 0); out = make_f3(0, 0, 0); return; }
  uint32_t iDp = idx % 180u, iDt = (idx / 180u) % 90u, iHt = idx / 16200u;
  MerlLinHalf h = merl_lin_half(iHt);
  float std_, ctd, cpd, spd;
  merl_lin_dtheta(iDt, std_, ctd);
  merl_lin_dphi(iDp, cpd, spd);
  merl_dirs_assemble(h, std_, ctd, cpd, spd, in, out);
}

// table layout: [0, 360) = (cy, sy, cz, sz) per iHt; [360, 540) = (sin, cos) theta_d per iDt; [540, 900) = (cos, sin) phi_d per iDp
constexpr int kMerlLinTabFloats = 900;
BBMCU_D void merl_lin_tab_fill(float* tab, int j)           // entry j of 90 + 90 + 180 = 360 table rows
{
  if(j < 90) { MerlLinHalf h = merl_lin_half((uint32_t)j); tab[4*j] = h.cy; tab[4*j + 1] = h.sy; tab[4*j + 2] = h.cz; tab[4*j + 3] = h.sz; }
  else if(j < 180) { float s, c; merl_lin_dtheta((uint32_t)(j - 90), s, c); tab[360 + 2*(j - 90)] = s; tab[360 + 2*(j - 90) + 1] = c; }
  else if(j < 360) { float c, s; merl_lin_dphi((uint32_t)(j - 180), c, s); tab[540 + 2*(j - 180)] = c; tab[540 + 2*(j - 180) + 1] = s; }
}
BBMCU_D void merl_dirs_tab(const float* tab, uint32_t idx, f3& in, f3& out)
{
  if(!(idx < kMerlBins)) { in = make_f3(0, 0, 0); out = make_f3(0, 0, 0); return; }
  uint32_t iDp = idx % 180u, iDt = (idx / 180u) % 90u, iHt = idx / 16200u;
  MerlLinHalf h; h.cy = tab[4*iHt]; h.sy = tab[4*iHt + 1]; h.cz = tab[4*iHt + 2]; h.sz = tab[4*iHt + 3];
  merl_dirs_assemble(h, tab[360 + 2*iDt], tab[360 + 2*iDt + 1], tab[540 + 2*iDp], tab[540 + 2*iDp + 1], in, out);
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

// spherical_linearizer(idx) is separable like the MERL one: (cos, sin) of each phi sample and (sin, cos) of each theta
// sample, then two products and the epsilon snap per direction.  spherical_dirs() evaluates the four pieces per index;
// spherical_dirs_tab() reads them from a table of 2 (n_in_phi + n_in_theta + n_out_phi + n_out_theta) floats built once
// per loss object with the same functions - identical bits, no transcendental per sample.
BBMCU_D void sph_lin_split(const SphericalGrid& g, uint64_t idx, uint32_t& ip, uint32_t& it, uint32_t& op, uint32_t& ot)
{
  uint64_t t = idx;
  ot = (uint32_t)(t % g.n_out_theta); t /= g.n_out_theta;
  op = (uint32_t)(t % g.n_out_phi);   t /= g.n_out_phi;
  it = (uint32_t)(t % g.n_in_theta);  t /= g.n_in_theta;
  ip = (uint32_t)t;
}
// entry j of the table: [in phi | in theta | out phi | out theta], two floats each: (cos, sin) for phi, (sin, cos) for theta
BBMCU_D void sph_lin_entry(const SphericalGrid& g, uint32_t j, float& a, float& b)
{
  // theta includes both end points (/(n-1), at least 1); phi does not (/n)
  if(j < g.n_in_phi) { float ph = (float)j * g.size_in_phi / (float)g.n_in_phi + g.start_in_phi; a = glibc_cosf(ph); b = glibc_sinf(ph); return; }
  j -= g.n_in_phi;
  if(j < g.n_in_theta) { float th = (float)j * g.size_in_theta / (float)(g.n_in_theta > 1 ? g.n_in_theta - 1 : 1) + g.start_in_theta; a = glibc_sinf(th); b = glibc_cosf(th); return; }
  j -= g.n_in_theta;
  if(j < g.n_out_phi) { float ph = (float)j * g.size_out_phi / (float)g.n_out_phi + g.start_out_phi; a = glibc_cosf(ph); b = glibc_sinf(ph); return; }
  j -= g.n_out_phi;
  { float th = (float)j * g.size_out_theta / (float)(g.n_out_theta > 1 ? g.n_out_theta - 1 : 1) + g.start_out_theta; a = glibc_sinf(th); b = glibc_cosf(th); }
}
BBMCU_HD uint64_t sph_lin_entries(const SphericalGrid& g) { return (uint64_t)g.n_in_phi + g.n_in_theta + g.n_out_phi + g.n_out_theta; }
BBMCU_D f3 sph_lin_assemble(float cp, float sp, float st, float ct) { return snap_eps(make_f3(mul_rn(cp, st), mul_rn(sp, st), ct)); }

BBMCU_D void spherical_dirs(const SphericalGrid& g, uint64_t idx, f3& in, f3& out)
{
  if(!(idx < g.size())) { in = make_f3(0, 0, 0); out = make_f3(0, 0, 0); return; }
  uint32_t ip, it, op, ot;
  sph_lin_split(g, idx, ip, it, op, ot);
  float cp, sp, st, ct;
  sph_lin_entry(g, ip, cp, sp); sph_lin_entry(g, g.n_in_phi + it, st, ct);
  in = sph_lin_assemble(cp, sp, st, ct);
  sph_lin_entry(g, g.n_in_phi + g.n_in_theta + op, cp, sp); sph_lin_entry(g, g.n_in_phi + g.n_in_theta + g.n_out_phi + ot, st, ct);
  out = sph_lin_assemble(cp, sp, st, ct);
}
BBMCU_D void spherical_dirs_tab(const SphericalGrid& g, const float* tab, uint64_t idx, f3& in, f3& out)
{
  if(!(idx < g.size())) { in = make_f3(0, 0, 0); out = make_f3(0, 0, 0); return; }
  uint32_t ip, it, op, ot;
  sph_lin_split(g, idx, ip, it, op, ot);
  const float* t1 = tab + 2*(size_t)g.n_in_phi; const float* t2 = t1 + 2*(size_t)g.n_in_theta; const float* t3 = t2 + 2*(size_t)g.n_out_phi;
#ifdef __CUDA_ARCH__
  in = sph_lin_assemble(__ldg(tab + 2*ip), __ldg(tab + 2*ip + 1), __ldg(t1 + 2*it), __ldg(t1 + 2*it + 1));
  out = sph_lin_assemble(__ldg(t2 + 2*op), __ldg(t2 + 2*op + 1), __ldg(t3 + 2*ot), __ldg(t3 + 2*ot + 1));
#else
  in = sph_lin_assemble(tab[2*ip], tab[2*ip + 1], t1[2*it], t1[2*it + 1]);
  out = sph_lin_assemble(t2[2*op], t2[2*op + 1], t3[2*ot], t3[2*ot + 1]);
#endif
}

} // namespace bbmcu
