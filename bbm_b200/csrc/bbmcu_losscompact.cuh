// Compact batched loss(+gradient) pass for Aggregate(Lambertian, microfacet lobe) - the shape of BASELINE configs[2]
// ("Ngan-style fit of Cook-Torrance") and of the Cook-Torrance entries of the reference's fits/*.fit.
//
// The generic tile kernel (bbmcu_losskernel.cuh) pushes dual numbers through the model templates once per (sample,
// parameter set): 186 instructions per sample and set for the Cook-Torrance aggregate, of which the model's mathematics
// is about a third.  Here the work is split by what it depends on (include/bsdfmodel/cooktorrance.h:29-34,
// include/ndf/beckmann.h:60-75, include/maskingshadowing/vgroove.h:41-46, include/bbm/fresnel_cook.h:47-55,
// include/loss/cosine_weighted_l2.h:25-34, cosine_weighted_log.h:32-43):
//   * per SAMPLE, once per tile (directions only): the Beckmann exponent's numerator T = tan^2(theta_h), the V-groove term
//     with every parameter-free factor folded in (G / (N z_i z_o cos^4 theta_h)), the Fresnel cosine, the metric's weights
//     and the measured value already multiplied by the cosine (or its logarithm): 9 floats - so a thread keeps 8
//     samples in registers instead of 4 and the warp reduction is paid half as often per sample;
//   * per PARAMETER SET, once per block (parameters only): 1/alpha^2, the normalisation of D, the coefficients of
//     d log D / d alpha, eta^2 - 1, albedo / pi - computed by one thread per set while the block stages its sets in shared
//     memory;
//   * per (sample, set): one exp2, the Fresnel term with its closed-form derivative, three channel errors, nine
//     accumulations.
// Factors common to a whole gradient column (2, 1/pi) are applied once per block partial.
// Samples below the horizon (neither lobe contributes: the term does not depend on the parameters) are summed once per
// tile into a constant instead of being evaluated per set.
// Results agree with the generic kernel to float rounding (tests/test_loss_compact_hostsim.py on the host,
// tests/test_gpu_round2.py::test_compact_loss_equals_generic_tile_kernel on the device); BBMCU_LOSS_NO_COMPACT=1 keeps
// the generic kernel.
#pragma once
#include "bbmcu_lossop.cuh"
#ifdef __CUDACC__
#include "bbmcu_losskernel.cuh"
#endif

namespace bbmcu {

BBMCU_D float c_ex2(float x)
{
#ifdef __CUDA_ARCH__
  float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return exp2f(x);
#endif
}

BBMCU_D float c_lg2(float x)
{
#ifdef __CUDA_ARCH__
  float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return log2f(x);
#endif
}

// ---- Fresnel policies: NS floats per parameter set, one float per sample, value + derivative -------------------------
// fresnel::cook (fresnel_cook.h:47-55): g = sqrt(eta^2 + c^2 - 1), p = (g - c)/(g + c), q = (c(g + c) - 1)/(c(g - c) + 1),
// F = p^2 (1 + q^2) / 2;  dp/dg = 2c/(g + c)^2, dq/dg = 2c(1 - c^2)/(c(g - c) + 1)^2, dg/deta = eta/g
struct CFCook
{
  static constexpr int NS = 2, NP = 1;
  BBMCU_D static void set(const float* a, float* d) { d[0] = a[0]*a[0] - 1.0f; d[1] = a[0]; }
  BBMCU_D static float inv(float c) { return c; }
  template<bool WG> BBMCU_D static void eval(const float* d, float c, float& Fv, float (&dF)[NP])
  {
    const float g2r = fmaf(c, c, d[0]);
    const float g2 = fmaxf(g2r, 1e-30f);                       // total internal reflection (eta < 1): g = 0, F = 1
    const float rs = q_rsqrt(g2), g = g2*rs;
    const float A = g - c, B = g + c;
    const float n2 = fmaf(c, B, -1.0f), d2 = fmaf(c, A, 1.0f);
    const float R = q_rcp(B*d2);
    const float iB = d2*R, id2 = B*R;
    const float p = A*iB, q = n2*id2;
    const float t = fmaf(q, q, 1.0f);
    const float pt = p*t;
    const float Fraw = 0.5f*p*pt;
    Fv = fmaxf(Fraw, 0.0f);
    if(WG)
    {
      const float c2 = c + c;
      const float dp = c2*(iB*iB), dq = (c2*fmaf(-c, c, 1.0f))*(id2*id2);
      const float dFdg = fmaf(pt, dp, (p*p)*(q*dq));
      const float ers = (g2r > 0.0f) ? d[1]*rs : 0.0f;
      dF[0] = dFdg*ers;
    }
  }
};
// fresnel::schlick of a reflectance at normal incidence (fresnel_schlick.h:48-51): R0 + (1 - R0) (1 - c)^5
struct CFSchlick
{
  static constexpr int NS = 1, NP = 1;
  BBMCU_D static void set(const float* a, float* d) { d[0] = a[0]; }
  BBMCU_D static float inv(float c) { return (float)schlick_w(c); }
  template<bool WG> BBMCU_D static void eval(const float* d, float w, float& Fv, float (&dF)[NP])
  {
    Fv = fmaf(-d[0], w, d[0]) + w;
    if(WG) dF[0] = 1.0f - w;
  }
};

// ---- D G policies: NS floats per set, NI floats per sample ------------------------------------------------------------
// isotropic Beckmann (beckmann.h:60-75) times the V-groove term (vgroove.h:41-46), which has no parameter:
//   D G / (N z_i z_o) = exp(-T / alpha^2) / (alpha^2 [pi]) * [G / (N z_i z_o cos^4 theta_h)],   T = tan^2 theta_h
//   d log D / d alpha = 2 T / alpha^3 - 2 / alpha
template<bool NORMALIZE> struct CDGBeckmannVGroove
{
  static constexpr int NS = 4, NP = 1, NI = 2;
  BBMCU_D static void set(const float* a, float* d)
  {
    const float al = a[0], ia2 = 1.0f/(al*al);
    d[0] = -1.44269504088896340736f*ia2; d[1] = NORMALIZE ? kInvPi*ia2 : ia2; d[2] = -2.0f/al; d[3] = 2.0f*ia2/al;
  }
  // caller guarantees in.z > 0, out.z > 0, inh > 0, outh > 0;  k = 1 / (N z_i z_o)
  BBMCU_D static void inv(f3 in, f3 out, f3 h, float inh, float outh, float k, float* I)
  {
    const float c2 = h.z*h.z;
    I[0] = (h.x*h.x + h.y*h.y)/c2;
    const float gi = q_div(2.0f*h.z*in.z, inh), go = q_div(2.0f*h.z*out.z, outh);
    const float G = fminf(1.0f, fminf(gi, go));
    I[1] = fminf(G*k/(c2*c2), 3.0e38f);
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& DG, float (&dDG)[NP])
  {
    DG = (c_ex2(I[0]*d[0])*I[1])*d[1];
    if(WG) dDG[0] = DG*fmaf(I[0], d[3], d[2]);
  }
};

// isotropic GGX (ggx.h:60-75) times the uncorrelated Smith term of its own G1 (ggx.h:184-187):
//   D = alpha^2 / (pi (q + alpha^2 (1 - q))^2),  q = h_x^2 + h_y^2;   G1(v) = 2 / (1 + s_v),  s_v = sqrt(1 + alpha^2 tan^2 theta_v)
//   D G / (N z_i z_o) = alpha^2 [4 / (pi N z_i z_o)] / (den^2 (1 + s_i)(1 + s_o)),   den = q + alpha^2 (1 - q)
//   d log (D G) / d alpha = 2/alpha - 4 alpha (1 - q) / den - alpha (t_i / (s_i (1 + s_i)) + t_o / (s_o (1 + s_o)))
// three reciprocals serve the value and the derivative: 1/den, 1/(s_i (1 + s_i)), 1/(s_o (1 + s_o))
struct CDGGgxSmith
{
  static constexpr int NS = 4, NP = 1, NI = 4;
  BBMCU_D static void set(const float* a, float* d) { const float al = a[0]; d[0] = al*al; d[1] = al; d[2] = 2.0f/al; d[3] = 4.0f*al; }
  // caller guarantees in.z > 0, out.z > 0, inh > 0, outh > 0;  k = 1 / (N z_i z_o)
  BBMCU_D static void inv(f3 in, f3 out, f3 h, float, float, float k, float* I)
  {
    I[0] = h.x*h.x + h.y*h.y;
    I[1] = fminf(sinTheta2(in)/(in.z*in.z), 1e30f);
    I[2] = fminf(sinTheta2(out)/(out.z*out.z), 1e30f);
    I[3] = fminf(4.0f*kInvPi*k, 3.0e38f);
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& DG, float (&dDG)[NP])
  {
    const float z2 = 1.0f - I[0];
    const float den = fmaf(d[0], z2, I[0]);
    const float si = q_sqrt(fmaf(d[0], I[1], 1.0f)), so = q_sqrt(fmaf(d[0], I[2], 1.0f));
    const float r1 = q_rcp(den), r2 = q_rcp(fmaf(si, si, si)), r3 = q_rcp(fmaf(so, so, so));
    DG = ((d[0]*I[3])*(r1*r1))*((r2*si)*(r3*so));
    if(WG) dDG[0] = DG*(fmaf(-d[3]*z2, r1, d[2]) - d[1]*fmaf(I[1], r2, I[2]*r3));
  }
};

// Low et al. 2012 microfacet distribution (ndf/low.h:40-60), unnormalised, times the V-groove term:
//   D = (1 + B x)^-C,  x = 1 - h_z;   d D / d B = -C D x / (1 + B x),   d D / d C = -D ln(1 + B x)
// The reference calls the double pow; C is of order one (0.02 .. 2.7 in fits/low_lowmicrofacet_E2.fit), so the 2^-22 absolute
// error of lg2.approx moves D by C ln 2 2^-22 < 1e-6 relative; x = 1 - h_z is exact in float for h_z >= 1/2.
struct CDGLowVGroove
{
  static constexpr int NS = 2, NP = 2, NI = 2;
  BBMCU_D static void set(const float* a, float* d) { d[0] = a[0]; d[1] = a[1]; }
  BBMCU_D static void inv(f3 in, f3 out, f3 h, float inh, float outh, float k, float* I)
  {
    I[0] = 1.0f - h.z;
    const float gi = q_div(2.0f*h.z*in.z, inh), go = q_div(2.0f*h.z*out.z, outh);
    I[1] = fminf(fminf(1.0f, fminf(gi, go))*k, 3.0e38f);
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& DG, float (&dDG)[NP])
  {
    const float b = fmaf(d[0], I[0], 1.0f);
    const float L = c_lg2(b);
    DG = c_ex2(-d[1]*L)*I[1];
    if(WG) { dDG[0] = -(d[1]*DG)*(I[0]*q_rcp(b)); dDG[1] = -DG*(L*0.69314718055994530942f); }
  }
};

// ---- lobe policies: a scaled lobe  scale_rgb * u(in, out; theta)  with a gray u -------------------------------------
//   NS floats per parameter set (set), NI >= 2 floats per sample (inv; zero = a sample where the lobe vanishes), NP fit
//   parameters after the scale, in attribute order;  eval<WG>(d, I, u, du[NP])
// microfacet<NDF, G, F, N> (microfacet.h:74-102) from a D G policy and a Fresnel policy
template<class M, class DG, class CF> struct LPMicrofacet
{
  static constexpr int NS = DG::NS + CF::NS, NP = DG::NP + CF::NP, NI = DG::NI + 1;
  static_assert(M::SCALE == 0 && M::NA == 3 + NP, "leading scale, every other attribute a fit parameter");
  BBMCU_D static void set(const float* a, float* d) { DG::set(a + (M::OFF_NDF - 3), d); CF::set(a + (M::OFF_F - 3), d + DG::NS); }
  BBMCU_D static void zero(float* I)
  {
#pragma unroll
    for(int i=0; i < DG::NI; ++i) I[i] = 0.0f;
    I[DG::NI] = CF::inv(0.5f);
  }
  BBMCU_D static void inv(f3 in, f3 out, float* I)
  {
    zero(I);
    if(!((in.z > 0.0f) && (out.z > 0.0f))) return;
    const f3 h = halfway(in, out);
    const float inh = dot(in, h), outh = dot(out, h);
    if(!((inh > 0.0f) && (outh > 0.0f))) return;
    DG::inv(in, out, h, inh, outh, q_rcp((float)M::norm() * (in.z*out.z)), I);
    I[DG::NI] = CF::inv(0.5f*(inh + outh));
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& u, float (&du)[NP])
  {
    float DGv, dDG[DG::NP], Fv, dF[CF::NP];
    DG::template eval<WG>(d, I, DGv, dDG);
    CF::template eval<WG>(d + DG::NS, I[DG::NI], Fv, dF);
    u = DGv*Fv;
    if(WG)
    {
#pragma unroll
      for(int j=0; j < DG::NP; ++j) du[j] = dDG[j]*Fv;
#pragma unroll
      for(int j=0; j < CF::NP; ++j) du[DG::NP + j] = DGv*dF[j];
    }
  }
};

// isotropic Ashikhmin-Shirley lobe (ashikhminshirley.h:60-80; the NganAshikhminShirley / LowAshikhminShirley fits):
//   u = F(h.in) (n + 1) / (8 pi) h_z^n / (h.in max(z_i, z_o));  attributes after the scale: Fresnel, n
//   d u / d n = u (1 / (n + 1) + ln h_z)
// per sample: log2 h_z (from the accurate logf: the exponent n log2 h_z reaches -126 before the power vanishes), 1 / (h.in max(z)),
// the Fresnel cosine
template<class M, class CF> struct LPAshikhminShirleyIso
{
  static constexpr int NS = CF::NS + 3, NP = CF::NP + 1, NI = 3;
  static_assert(M::SCALE == 0 && M::NA == 3 + NP && M::OFF_N == 3 + CF::NP, "scale, Fresnel, exponent");
  BBMCU_D static void set(const float* a, float* d) { CF::set(a, d); const float n = a[CF::NP]; d[CF::NS] = n; d[CF::NS + 1] = (n + 1.0f)/M::kEightPi; d[CF::NS + 2] = 1.0f/(n + 1.0f); }
  BBMCU_D static void zero(float* I) { I[0] = 0.0f; I[1] = 0.0f; I[2] = CF::inv(0.5f); }
  BBMCU_D static void inv(f3 in, f3 out, float* I)
  {
    zero(I);
    if(!((in.z > 0.0f) && (out.z > 0.0f))) return;
    const f3 h = halfway(in, out);
    const float hin = dot(h, in);
    I[0] = fmaxf(logf(h.z)*1.44269504088896340736f, -1e4f);
    I[1] = fminf(1.0f/(hin*fmaxf(in.z, out.z)), 3.0e38f);
    I[2] = CF::inv(hin);
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& u, float (&du)[NP])
  {
    float Fv, dF[CF::NP];
    CF::template eval<WG>(d, I[2], Fv, dF);
    const float lobe = (d[CF::NS + 1]*c_ex2(d[CF::NS]*I[0]))*I[1];
    u = Fv*lobe;
    if(WG)
    {
#pragma unroll
      for(int j=0; j < CF::NP; ++j) du[j] = dF[j]*lobe;
      du[CF::NP] = u*fmaf(I[0], 0.69314718055994530942f, d[CF::NS + 2]);
    }
  }
};

// Phong / NganBlinnPhong (phong.h:43-60): u = (n + 2) / (2 pi) cos_alpha^n, cos_alpha = max(reflect_z(in).out, 0); d u / d n = u (1/(n + 2) + ln cos_alpha)
template<class M> struct LPPhong
{
  static constexpr int NS = 3, NP = 1, NI = 2;
  static_assert(M::SCALE == 0 && M::NA == 4, "scale, exponent");
  BBMCU_D static void set(const float* a, float* d) { const float n = a[0]; d[0] = n; d[1] = (n + 2.0f)*(0.5f*kInvPi); d[2] = 1.0f/(n + 2.0f); }
  BBMCU_D static void zero(float* I) { I[0] = -1e4f; I[1] = 0.0f; }
  // the lobe lives on in.z >= 0, out.z >= 0 - the caller's domain
  BBMCU_D static void inv(f3 in, f3 out, float* I)
  {
    const float ca = fmaxf(dot(reflect_z(in), out), 0.0f);
    I[0] = (ca > 0.0f) ? fmaxf(logf(ca)*1.44269504088896340736f, -1e4f) : -1e4f;
    I[1] = 0.0f;
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& u, float (&du)[NP])
  {
    u = d[1]*c_ex2(d[0]*I[0]);
    if(WG) du[0] = u*fmaf(I[0], 0.69314718055994530942f, d[2]);
  }
};

// isotropic Lafortune lobe with Ngan's normalisation (lafortune.h:50-70, ngan.h:54-129):
//   u = max(d, 0)^n nrm,  d = Cxy (x_i x_o + y_i y_o) + Cz z_i z_o,  nrm = (n + 2) / (2 pi max(Cz^2, Cxy^2)^(n/2))
// nrm and its three parameter derivatives are per-set constants; d depends on both, so its logarithm (the accurate logf: n is
// large for sharp lobes) is taken per sample and set
template<class M> struct LPNganLafortune
{
  static constexpr int NS = 7, NP = 3, NI = 2;
  static_assert(M::SCALE == 0 && M::NA == 6, "scale, Cxy, Cz, n");
  BBMCU_D static void set(const float* a, float* d)
  {
    const float cxy = a[0], cz = a[1], n = a[2];
    const float cz2 = cz*cz, cxy2 = cxy*cxy, m2 = fmaxf(cz2, cxy2);
    const float nrm = (n + 2.0f)*(0.5f*kInvPi)/powf(m2, n*0.5f);
    d[0] = cxy; d[1] = cz; d[2] = n; d[3] = nrm;
    // m_max(cz2, cxy2) of the dual-number path keeps the first argument on a tie
    d[4] = (cxy2 > cz2) ? -nrm*n*cxy/m2 : 0.0f;
    d[5] = (cxy2 > cz2) ? 0.0f : -nrm*n*cz/m2;
    d[6] = nrm*(1.0f/(n + 2.0f) - 0.5f*logf(m2));
  }
  BBMCU_D static void zero(float* I) { I[0] = 0.0f; I[1] = 0.0f; }
  BBMCU_D static void inv(f3 in, f3 out, float* I)
  {
    zero(I);
    if(!((in.z > 0.0f) && (out.z > 0.0f))) return;
    I[0] = in.x*out.x + in.y*out.y; I[1] = in.z*out.z;
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& u, float (&du)[NP])
  {
    const float dd = fmaf(d[0], I[0], d[1]*I[1]);
    const bool on = dd > 0.0f;
    const float ds = on ? dd : 1.0f;
    const float ln = logf(ds);
    const float pw = on ? c_ex2(d[2]*(ln*1.44269504088896340736f)) : 0.0f;
    u = pw*d[3];
    if(WG)
    {
      const float g = (d[2]*u)*q_rcp(ds);                        // d u / d d
      du[0] = fmaf(g, I[0], pw*d[4]);
      du[1] = fmaf(g, I[1], pw*d[5]);
      du[2] = pw*fmaf(d[3], ln, d[6]);
    }
  }
};

// isotropic Ward lobes (ward.h:45-60, wardduer.h:55-70; the NganWard / NganWardDuer fits): with H = in + out (not normalised)
//   u = exp(-T / alpha^2) / (4 pi alpha^2 w),  T = (H_x^2 + H_y^2) / H_z^2,  w = sqrt(z_i z_o) (Ward) or z_i z_o (Duer)
//   d u / d alpha = u (2 T / alpha^3 - 2 / alpha)
// At the horizon (z = 0) 1 / w is infinite and the reference returns Inf or NaN: nothing is clamped here, the
// loss of a grid that touches the horizon is non-finite in both.
template<class M, int VARIANT> struct LPWardIso
{
  static constexpr int NS = 4, NP = 1, NI = 2;
  static_assert(M::SCALE == 0 && M::NA == 4 && VARIANT <= 1, "scale, roughness");
  BBMCU_D static void set(const float* a, float* d)
  {
    const float al = a[0], ia2 = 1.0f/(al*al);
    d[0] = -1.44269504088896340736f*ia2; d[1] = ia2; d[2] = -2.0f/al; d[3] = 2.0f*ia2/al;
  }
  BBMCU_D static void zero(float* I) { I[0] = 0.0f; I[1] = 0.0f; }
  // the lobe lives on in.z >= 0, out.z >= 0 - the caller's domain
  BBMCU_D static void inv(f3 in, f3 out, float* I)
  {
    const f3 H = in + out;
    I[0] = (H.x*H.x + H.y*H.y)/(H.z*H.z);
    I[1] = 1.0f/(M::kFourPi * (VARIANT == 0 ? sqrtf(in.z*out.z) : in.z*out.z));
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& u, float (&du)[NP])
  {
    u = (c_ex2(I[0]*d[0])*I[1])*d[1];
    if(WG) du[0] = u*fmaf(I[0], d[3], d[2]);
  }
};

// Low et al. 2012 smooth-surface lobe (lowsmooth.h:50-70): u = S Q,  S = (1 + B Dp^2)^-C,  Q = fresnel::cook(eta, cos_D);
// attributes after the scale A: B, C, eta.  Same exponent range as CDGLowVGroove (C of order one).
template<class M> struct LPLowSmooth
{
  static constexpr int NS = 2 + CFCook::NS, NP = 3, NI = 2;
  static_assert(M::SCALE == 0 && M::NA == 6, "scale, B, C, eta");
  BBMCU_D static void set(const float* a, float* d) { d[0] = a[0]; d[1] = a[1]; CFCook::set(a + 2, d + 2); }
  BBMCU_D static void zero(float* I) { I[0] = 0.0f; I[1] = 0.5f; }
  // the lobe lives on in.z >= 0, out.z >= 0 - the caller's domain
  BBMCU_D static void inv(f3 in, f3 out, float* I)
  {
    const float sx = in.x + out.x, sy = in.y + out.y, dx = in.x - out.x, dy = in.y - out.y;
    I[0] = sx*sx + sy*sy;
    I[1] = (float)safe_sqrt_d(1.0 - 0.25*(double)(dx*dx + dy*dy));
  }
  template<bool WG> BBMCU_D static void eval(const float* d, const float* I, float& u, float (&du)[NP])
  {
    const float b = fmaf(d[0], I[0], 1.0f);
    const float L = c_lg2(b);
    const float S = c_ex2(-d[1]*L);
    float Q, dQ[1];
    CFCook::template eval<WG>(d + 2, I[1], Q, dQ);
    u = S*Q;
    if(WG) { du[0] = -(d[1]*u)*(I[0]*q_rcp(b)); du[1] = -u*(L*0.69314718055994530942f); du[2] = S*dQ[0]; }
  }
};

// ---- Aggregate(Lambertian, M) - or M alone (LAMB = false) - with the lobe M described by a lobe policy ----------------
// fit parameters in attribute order: [albedo rgb,] scale rgb, the lobe's parameters
template<class M, class LP, int SPT = 8, bool LAMB = true> struct CompactPair
{
  using Model = M;
  static constexpr int kSPT = SPT, kThreads = 1024/SPT;        // samples per thread x threads = one tile of 1024 samples
  static constexpr int L0 = LAMB ? 3 : 0;                      // attribute floats (and gradient columns) of the diffuse lobe
  static constexpr int NRAW = L0 + M::NA;                      // attribute floats of a parameter set
  static constexpr int P = L0 + 3 + LP::NP, C = 1 + P;
  static constexpr int OFF_LP = 6, NSET = (6 + LP::NS + 3) & ~3;
  static_assert(NRAW == L0 + 3 + LP::NP && M::SCALE == 0, "every attribute is a fit parameter; leading scale");
  struct Sample { float I[LP::NI]; float cm, w, wc, r[3]; };

  BBMCU_D static void set(const float* raw, float* d)
  {
#pragma unroll
    for(int i=0; i < NSET; ++i) d[i] = 0.0f;
    if(LAMB) { d[0] = raw[0]*kInvPi; d[1] = raw[1]*kInvPi; d[2] = raw[2]*kInvPi; }
    d[3] = raw[L0]; d[4] = raw[L0 + 1]; d[5] = raw[L0 + 2];
    LP::set(raw + L0 + 3, d + OFF_LP);
  }
  // factor of column j of the block partial: the per-sample sums leave out 2 (d e / d v) and 1/pi (d v / d albedo)
  BBMCU_HD static double col_scale(int j) { return j == 0 ? 1.0 : ((LAMB && j <= 3) ? 2.0/kPiD : 2.0); }

  // direction-only part of a sample.  state: 0 regular, 1 padding (past the end of the shard), 2 below the horizon - both
  // lobes are zero there (lambertian.h:38-44, microfacet.h:74-81), so the term does not depend on the parameters: the
  // sample is neutral in the per-set loop (cos_i and the measured value read as 0, the lobe at its zero) and keeps the
  // weight in `w` and the true cos_i in `wc` for below_const()
  BBMCU_D static Sample make_geom(int metric, f3 in, f3 out, bool live, int& state)
  {
    Sample s;
    LP::zero(s.I);
    s.cm = 0.0f; s.w = 0.0f; s.wc = 0.0f; s.r[0] = s.r[1] = s.r[2] = 0.0f;
    state = 1;
    if(!live) return s;
    const float cm = fmaxf(in.z, 0.0f), w = metric_weight(metric, in, out);
    if(!((in.z >= 0.0f) && (out.z >= 0.0f)))
    {
      s.w = fminf(w, 3.0e38f); s.wc = cm; state = 2;             // cm = 0 and r = 0 make every term of the per-set loop vanish
      return s;
    }
    state = 0;
    LP::inv(in, out, s.I);
    s.cm = cm; s.w = w; s.wc = w*cm;
    return s;
  }
  // the measured value of one material at the sample (regular samples only; the others keep zeros)
  template<bool LOG> BBMCU_D static void set_ref(Sample& s, const Spec<float>& ref, bool regular)
  {
    if(LOG) { s.r[0] = regular ? logf(1.0f + ref.r*s.cm) : 0.0f; s.r[1] = regular ? logf(1.0f + ref.g*s.cm) : 0.0f; s.r[2] = regular ? logf(1.0f + ref.b*s.cm) : 0.0f; }
    else    { s.r[0] = regular ? ref.r*s.cm : 0.0f; s.r[1] = regular ? ref.g*s.cm : 0.0f; s.r[2] = regular ? ref.b*s.cm : 0.0f; }
  }
  BBMCU_D static float below_const(int metric, const Sample& s, const Spec<float>& ref) { return loss_term_g(metric, s.wc, s.w, Spec<float>(0.0f), ref, nullptr); }

  template<bool WG, bool LOG> BBMCU_D static void accumulate(const float* d, const Sample& s, float* acc)
  {
    float u, du[LP::NP];
    LP::template eval<WG>(d + OFF_LP, s.I, u, du);
    const float v0 = fmaf(d[3], u, d[0]), v1 = fmaf(d[4], u, d[1]), v2 = fmaf(d[5], u, d[2]);
    float t0, t1, t2, k0, k1, k2;
    if(LOG)
    {
      const float a0 = fmaf(v0, s.cm, 1.0f), a1 = fmaf(v1, s.cm, 1.0f), a2 = fmaf(v2, s.cm, 1.0f);
      t0 = logf(a0) - s.r[0]; t1 = logf(a1) - s.r[1]; t2 = logf(a2) - s.r[2];
      if(WG) { k0 = s.wc*q_rcp(a0); k1 = s.wc*q_rcp(a1); k2 = s.wc*q_rcp(a2); }
    }
    else
    {
      t0 = fmaf(v0, s.cm, -s.r[0]); t1 = fmaf(v1, s.cm, -s.r[1]); t2 = fmaf(v2, s.cm, -s.r[2]);
      if(WG) { k0 = k1 = k2 = s.wc; }
    }
    acc[0] = fmaf(fmaf(t2, t2, fmaf(t1, t1, t0*t0)), s.w, acc[0]);
    if(WG)
    {
      const float e0 = t0*k0, e1 = t1*k1, e2 = t2*k2;        // (d e / d v) / 2
      if(LAMB) { acc[1] += e0; acc[2] += e1; acc[3] += e2; }
      acc[L0 + 1] = fmaf(e0, u, acc[L0 + 1]); acc[L0 + 2] = fmaf(e1, u, acc[L0 + 2]); acc[L0 + 3] = fmaf(e2, u, acc[L0 + 3]);
      const float sdv = fmaf(e2, d[5], fmaf(e1, d[4], e0*d[3]));
#pragma unroll
      for(int j=0; j < LP::NP; ++j) acc[L0 + 4 + j] = fmaf(sdv, du[j], acc[L0 + 4 + j]);
    }
  }
};

// which specular lobes have a compact kernel: CompactOf<M>::type for Aggregate(Lambertian, M), CompactSingleOf<M>::type for M alone
template<class M> struct CompactOf { static constexpr bool value = false; };
template<class CP> struct WithoutLambertian;
template<class M, class LP, int SPT> struct WithoutLambertian<CompactPair<M, LP, SPT, true>> { using type = CompactPair<M, LP, SPT, false>; };
template<class M, class = void> struct CompactSingleOf { static constexpr bool value = false; };
template<class M> struct CompactSingleOf<M, typename std::enable_if<CompactOf<M>::value>::type> { static constexpr bool value = true; using type = typename WithoutLambertian<typename CompactOf<M>::type>::type; };
template<bool NRM, int NORM> struct CompactOf<Microfacet<NdfBeckmann<false, NRM>, GVGroove, FresnelCookIor, NORM, true>>
{ using M = Microfacet<NdfBeckmann<false, NRM>, GVGroove, FresnelCookIor, NORM, true>; static constexpr bool value = true; using type = CompactPair<M, LPMicrofacet<M, CDGBeckmannVGroove<NRM>, CFCook>>; };
template<bool NRM, int NORM> struct CompactOf<Microfacet<NdfBeckmann<false, NRM>, GVGroove, FresnelSchlickR0, NORM, true>>
{ using M = Microfacet<NdfBeckmann<false, NRM>, GVGroove, FresnelSchlickR0, NORM, true>; static constexpr bool value = true; using type = CompactPair<M, LPMicrofacet<M, CDGBeckmannVGroove<NRM>, CFSchlick>>; };
template<int NORM> struct CompactOf<Microfacet<NdfLow, GVGroove, FresnelCookIor, NORM, true>>
{ using M = Microfacet<NdfLow, GVGroove, FresnelCookIor, NORM, true>; static constexpr bool value = true; using type = CompactPair<M, LPMicrofacet<M, CDGLowVGroove, CFCook>>; };
template<int NORM> struct CompactOf<Microfacet<NdfGGX<false>, GUncorrelated, FresnelCookIor, NORM, true>>
{ using M = Microfacet<NdfGGX<false>, GUncorrelated, FresnelCookIor, NORM, true>; static constexpr bool value = true; using type = CompactPair<M, LPMicrofacet<M, CDGGgxSmith, CFCook>, 4>; };
template<> struct CompactOf<AshikhminShirley<FresnelSchlickR0, false, true>>
{ using M = AshikhminShirley<FresnelSchlickR0, false, true>; static constexpr bool value = true; using type = CompactPair<M, LPAshikhminShirleyIso<M, CFSchlick>>; };
template<> struct CompactOf<AshikhminShirley<FresnelCookIor, false, true>>
{ using M = AshikhminShirley<FresnelCookIor, false, true>; static constexpr bool value = true; using type = CompactPair<M, LPAshikhminShirleyIso<M, CFCook>>; };
template<> struct CompactOf<Ward<false, 0>> { static constexpr bool value = true; using type = CompactPair<Ward<false, 0>, LPWardIso<Ward<false, 0>, 0>>; };
template<> struct CompactOf<Ward<false, 1>> { static constexpr bool value = true; using type = CompactPair<Ward<false, 1>, LPWardIso<Ward<false, 1>, 1>>; };
template<> struct CompactOf<LowSmooth> { static constexpr bool value = true; using type = CompactPair<LowSmooth, LPLowSmooth<LowSmooth>>; };
template<> struct CompactOf<Phong> { static constexpr bool value = true; using type = CompactPair<Phong, LPPhong<Phong>>; };
template<> struct CompactOf<Lafortune<false, true>> { static constexpr bool value = true; using type = CompactPair<Lafortune<false, true>, LPNganLafortune<Lafortune<false, true>>>; };

constexpr int kCThreads = 128, kCSPT = 8;     // the default shape (CompactPair<..., SPT = 8>)

#ifdef __CUDACC__

// Block (x, y, z): sample tiles x, x + gridDim.x, ...; parameter sets [y kpb, (y+1) kpb) of each material; materials
// [z mpb, (z+1) mpb).  The direction-only part of a tile's samples is computed ONCE and serves every material of the block
// (a multi-material launch reads 12 B per sample and material and nothing else: with one parameter set per material
// this is a DRAM stream); per material the three measured planes are loaded and folded into the samples, then the
// parameter sets of that material run.  Block partials per (material, set, tile), written in groups of kTileKChunk rows.
template<class CL, bool WG, bool LOG>
__global__ void __launch_bounds__(CL::kThreads, 512/CL::kThreads) k_loss_tile_compact(const LossArgs a, int K, int k_per_block, int n_tiles, int m_per_block)
{
  constexpr int C = WG ? CL::C : 1;
  constexpr int NW = CL::kThreads/32, kCSPT = CL::kSPT, kCThreads = CL::kThreads;
  static_assert(kCThreads*kCSPT == kTileSamples, "same tiles (and block partial rows) as the generic tile kernel");
  extern __shared__ __align__(16) float s_set[];                // (m1 - m0) x (k1 - k0) x NSET
  __shared__ float s_red[kTileKChunk][NW][C];
  __shared__ float s_lin[kMerlLinTabFloats];
  const int m0 = blockIdx.z * m_per_block, m1 = min(a.n_materials, m0 + m_per_block);
  const int k0 = blockIdx.y * k_per_block, k1 = min(K, k0 + k_per_block), nk = k1 - k0;
  for(int idx = threadIdx.x; idx < (m1 - m0)*nk; idx += blockDim.x)
  {
    const int mm = idx / nk, k = idx - mm*nk;
    const size_t set = (size_t)(m0 + mm)*K + (size_t)(k0 + k);
    float raw[CL::NRAW];
#pragma unroll
    for(int j=0; j < CL::NRAW; ++j) raw[j] = loss_attr(a, set*a.attr_stride + j);
    float d[CL::NSET];
    CL::set(raw, d);
#pragma unroll
    for(int j=0; j < CL::NSET; ++j) s_set[(size_t)idx*CL::NSET + j] = d[j];
  }
  loss_stage_lin(a, s_lin);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for(int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    typename CL::Sample smp[kCSPT];
    unsigned regular = 0u, below = 0u;
#pragma unroll
    for(int s=0; s < kCSPT; ++s)
    {
      const size_t i = (size_t)tile*kTileSamples + (size_t)s*kCThreads + threadIdx.x;
      const bool live = i < a.n;
      f3 in, out;
      loss_dirs(a, s_lin, live ? i : 0, in, out);
      int state;
      smp[s] = CL::make_geom(a.metric, in, out, live, state);
      regular |= (state == 0 ? 1u : 0u) << s; below |= (state == 2 ? 1u : 0u) << s;
    }
    // rows of s_red waiting to be written: (material, set) pairs seq0, seq0 + 1, ... of this block's m-major sequence
    int pending = 0, seq0 = 0;
    const float4* sp = reinterpret_cast<const float4*>(s_set);
    auto flush = [&]() {
      __syncthreads();
      for(int t = threadIdx.x; t < pending*(1 + a.P); t += blockDim.x)
      {
        const int q = t / (1 + a.P), j = t - q*(1 + a.P);
        const int seq = seq0 + q, mm = seq / nk, kk = seq - mm*nk;
        const size_t row = (size_t)(m0 + mm)*K + (size_t)(k0 + kk);
        double v = 0.0;
        if(j < C)
        {
#pragma unroll
          for(int w=0; w < NW; ++w) v += (double)s_red[q][w][j];
          v *= CL::col_scale(j);
        }
        a.partial[(row*(1 + a.P) + j)*n_tiles + tile] = v;
      }
      __syncthreads();
      seq0 += pending; pending = 0;
    };
    for(int mat = m0; mat < m1; ++mat)
    {
      const float* refp = a.ref + (size_t)mat * a.ref_stride;
      float e_const = 0.0f;
      // the three measured planes of this material, loaded straight into the registers they end up in
#pragma unroll
      for(int s=0; s < kCSPT; ++s)
      {
        const size_t i = (size_t)tile*kTileSamples + (size_t)s*kCThreads + threadIdx.x;
        const size_t ii = i < a.n ? i : 0;
        smp[s].r[0] = __ldg(refp + ii); smp[s].r[1] = __ldg(refp + a.n + ii); smp[s].r[2] = __ldg(refp + 2*a.n + ii);
      }
      if(below)
      {
#pragma unroll
        for(int s=0; s < kCSPT; ++s) if((below >> s) & 1u) e_const += CL::below_const(a.metric, smp[s], Spec<float>(smp[s].r[0], smp[s].r[1], smp[s].r[2]));
      }
#pragma unroll
      for(int s=0; s < kCSPT; ++s) CL::template set_ref<LOG>(smp[s], Spec<float>(smp[s].r[0], smp[s].r[1], smp[s].r[2]), (regular >> s) & 1u);
      // parameter sets of this material in chunks of up to kTileKChunk rows of s_red; the rows are written out when the
      // next chunk would not fit (nk >= kTileKChunk: after every chunk; a few sets per material: several materials share a flush)
      for(int kc=0; kc < nk; kc += kTileKChunk)
      {
        const int nkk = min(kTileKChunk, nk - kc);
        float (*red_rows)[NW][C] = s_red + pending;
        for(int kk=0; kk < nkk; ++kk)
        {
          float d[CL::NSET];
#pragma unroll
          for(int j=0; j < CL::NSET/4; ++j) { const float4 q = sp[j]; d[4*j] = q.x; d[4*j + 1] = q.y; d[4*j + 2] = q.z; d[4*j + 3] = q.w; }
          sp += CL::NSET/4;                                     // the block's sets lie in the order they are visited
          float acc[CL::C];
          acc[0] = e_const;
#pragma unroll
          for(int j=1; j < CL::C; ++j) acc[j] = 0.0f;
#pragma unroll
          for(int s=0; s < kCSPT; ++s) CL::template accumulate<WG, LOG>(d, smp[s], acc);
          float red[C];
#pragma unroll
          for(int j=0; j < C; ++j) red[j] = acc[j];
          int ridx; float rval; bool rwriter;
          warp_reduce_multi<C>(red, lane, ridx, rval, rwriter);
          if(rwriter && ridx < C) red_rows[kk][warp][ridx] = rval;
        }
        pending += nkk;
        if(pending + min(kTileKChunk, nk) > kTileKChunk) flush();
      }
    }
    if(pending) flush();
  }
}

// launch shape: tiles x k-splits x material groups.  A block costs its tile's direction-only work once, per material the
// reference planes, per (material, set) one evaluation; blocks are equal and `slots` of them are resident at a time, so
// the launch takes ceil(blocks / slots) block durations: the (k-split, material-split) pair that minimises the product,
// then the finest pair within 2 % of it (weights: instructions per sample from profiles/r02_s43_ncu_loss_tile_compact_ct.txt)
struct CompactShape { unsigned tiles, ksplit, msplit; int kpb, mpb; };
inline CompactShape loss_compact_shape(size_t n, size_t K, size_t M, int nset, size_t slots)
{
  static thread_local struct { size_t n, K, M, slots; int nset; CompactShape s; bool valid; } memo = {0, 0, 0, 0, 0, {}, false};
  if(memo.valid && memo.n == n && memo.K == K && memo.M == M && memo.slots == slots && memo.nset == nset) return memo.s;
  CompactShape r;
  r.tiles = (unsigned)std::max<size_t>(1, (n + kTileSamples - 1) / kTileSamples);
  const size_t max_sets = std::max<size_t>(1, (size_t)(24*1024) / ((size_t)nset*sizeof(float)));      // (material, set) pairs a block can stage
  const double w_set = 80.0, w_mat = 30.0, w_tile = 330.0;
  auto cost_of = [&](size_t ks, size_t ms, bool& valid) {
    const size_t kpb = (K + ks - 1) / ks, mpb = (M + ms - 1) / ms;
    valid = ((K + kpb - 1) / kpb == ks) && ((M + mpb - 1) / mpb == ms) && kpb*mpb <= max_sets;
    const size_t blocks = (size_t)r.tiles * ks * ms, rounds = (blocks + slots - 1) / slots;
    return (double)rounds * (w_tile + (double)mpb * (w_mat + (double)kpb * w_set));
  };
  double best = 0.0; size_t bks = K, bms = M;
  for(size_t ms = 1; ms <= M; ++ms)
    for(size_t ks = 1; ks <= K && ks <= 256; ++ks) { bool v; const double c = cost_of(ks, ms, v); if(v && (best == 0.0 || c < best)) best = c; }
  size_t fine = 0;
  for(size_t ms = 1; ms <= M; ++ms)
    for(size_t ks = 1; ks <= K && ks <= 256; ++ks) { bool v; const double c = cost_of(ks, ms, v); if(v && c <= best*1.02 && ks*ms >= fine) { fine = ks*ms; bks = ks; bms = ms; } }
  if(best == 0.0) { bms = M; bks = (K + max_sets - 1) / max_sets; }   // more sets than 256 ranges can stage (K > 256 x 512): as many ranges as it takes, one material per block
  r.kpb = (int)((K + bks - 1) / bks); r.mpb = (int)((M + bms - 1) / bms);
  r.ksplit = (unsigned)((K + r.kpb - 1) / r.kpb); r.msplit = (unsigned)((M + r.mpb - 1) / r.mpb);
  memo = {n, K, M, slots, nset, r, true};
  return r;
}

template<class CL> static void launch_loss_compact_static(cudaStream_t s, const LossArgs& a, unsigned K)
{
  const bool log_metric = a.metric > METRIC_BIERON_L2;
  const int wg = a.want_grad ? 1 : 0, lg = log_metric ? 1 : 0;
  static int per_sm[2][2] = {{0, 0}, {0, 0}};
  if(per_sm[wg][lg] == 0)
  {
    int v = 0;
    cudaError_t e = wg ? (lg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_loss_tile_compact<CL, true, true>, CL::kThreads, 24*1024)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_loss_tile_compact<CL, true, false>, CL::kThreads, 24*1024))
                       : (lg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_loss_tile_compact<CL, false, true>, CL::kThreads, 24*1024)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_loss_tile_compact<CL, false, false>, CL::kThreads, 24*1024));
    per_sm[wg][lg] = (e == cudaSuccess && v > 0) ? v : 4;
  }
  const size_t slots = (size_t)a.sm_count * (size_t)per_sm[wg][lg];
  const CompactShape sh = loss_compact_shape(a.n, K, (size_t)a.n_materials, CL::NSET, slots);
  const size_t smem = (size_t)sh.kpb*sh.mpb*CL::NSET*sizeof(float);
  // blocks along the tile axis: all tiles unless that makes more than 8 rounds of resident blocks - then each block walks
  // over several tiles and the staging of its parameter sets is paid once
  const size_t other = (size_t)sh.ksplit * sh.msplit, target = slots * 8;
  unsigned gx = sh.tiles;
  if((size_t)sh.tiles * other > target) gx = (unsigned)std::max<size_t>(1, (target + other - 1) / other);
  if(gx > sh.tiles) gx = sh.tiles;
  const dim3 grid(gx, sh.ksplit, sh.msplit);
  if(wg) { if(lg) k_loss_tile_compact<CL, true, true><<<grid, CL::kThreads, smem, s>>>(a, (int)K, sh.kpb, (int)sh.tiles, sh.mpb);
           else   k_loss_tile_compact<CL, true, false><<<grid, CL::kThreads, smem, s>>>(a, (int)K, sh.kpb, (int)sh.tiles, sh.mpb); }
  else   { if(lg) k_loss_tile_compact<CL, false, true><<<grid, CL::kThreads, smem, s>>>(a, (int)K, sh.kpb, (int)sh.tiles, sh.mpb);
           else   k_loss_tile_compact<CL, false, false><<<grid, CL::kThreads, smem, s>>>(a, (int)K, sh.kpb, (int)sh.tiles, sh.mpb); }
}

// Aggregate(Lambertian, model) - or the model alone - through the compact kernel; false if `model` has none (bbmcu_loss_pair_compact.cu)
bool launch_loss_pair_compact(int model, cudaStream_t, const LossArgs&, unsigned K);
bool launch_loss_single_compact(int model, cudaStream_t, const LossArgs&, unsigned K);
#endif

} // namespace bbmcu
