// Microfacet building blocks: Fresnel terms, normal distributions (D, G1, sample, pdf),
// joint masking-shadowing terms and the microfacet combinator.
//
// Behaviour follows (restated, not copied):
//   include/bbm/fresnel_cook.h:41-56, fresnel_schlick.h:42-52, fresnel_complex.h:29-51,
//   include/core/ior.h:58-72, include/bsdfmodel/bagher.h:38-51
//   include/ndf/beckmann.h:49-201, ggx.h:50-189, phong.h:40-131, low.h:44-132,
//   include/ndf/studentt.h:41-197, sgd.h:48-193
//   include/maskingshadowing/vgroove.h:30-47, uncorrelated.h:30-42, heightcorrelated.h:30-54,
//   include/maskingshadowing/vanginneken.h:37-74
//   include/bsdfmodel/microfacet.h:74-199, scaledmodel.h:50-67
//
// Every eval-side function is a template over the scalar type T (float, or Dual<N> for the
// analytic parameter gradient); directions are always plain float.  sample / pdf are float only.
#pragma once
#include "bbmcu_math.cuh"

namespace bbmcu {

// =============================================================================================
// Fresnel
// =============================================================================================
// ior::convert(ior <- reflectance): (1 + sqrt(r)) / (1 - sqrt(r))      (core/ior.h:66-72)
template<class T> BBMCU_D T ior_from_reflectance(const T& r) { T t = m_safe_sqrt(r); return (1.0f + t) / (1.0f - t); }
// ior::convert(reflectance <- ior): ((n-1)/(n+1))^2                      (core/ior.h:58-64)
template<class T> BBMCU_D T reflectance_from_ior(const T& n) { T t = (n - 1.0f) / (n + 1.0f); return t*t; }

// Cook-Torrance Fresnel of an index of refraction (fresnel_cook.h:47-55)
template<class T> BBMCU_D T fresnel_cook(const T& eta, float c)
{
  T g = m_safe_sqrt(eta*eta + c*c - 1.0f);
  T a = (g - c) / (g + c);
  T b = (c*(g + c) - 1.0f) / (c*(g - c) + 1.0f);
  return m_max(0.5f * (a*a) * (1.0f + b*b), 0.0f);
}
// quick variant for eval: g keeps the IEEE square root (g - c cancels for eta -> 1), the two quotients are final values.
// reflectance() keeps the IEEE template above: it feeds the lobe-selection weights of Aggregate sampling, where a
// last-bit change of a weight moves the renormalised xi and with it the sampled direction.
BBMCU_D float fresnel_cook_q(const float& eta, float c)
{
  float g = m_safe_sqrt(eta*eta + c*c - 1.0f);
  float a = q_div(g - c, g + c);
  float b = q_div(c*(g + c) - 1.0f, c*(g - c) + 1.0f);
  return m_max(0.5f * (a*a) * (1.0f + b*b), 0.0f);
}
// Schlick of a reflectance at normal incidence: R0 + (1-R0) * pow(1-c, 5.0)  (fresnel_schlick.h:48-51).
// pow(float, 5.0) is a double, so the product and the sum run in double and are rounded once.
BBMCU_D double schlick_w(float c) { double w = (double)(1.0f - c); double w2 = w*w; return w2*w2*w; }
BBMCU_D float fresnel_schlick(float R0, float c) { return (float)((double)R0 + (double)(1.0f - R0) * schlick_w(c)); }
template<int N> BBMCU_D Dual<N> fresnel_schlick(const Dual<N>& R0, float c) { float w = (float)schlick_w(c); Dual<N> r = R0 + (1.0f - R0) * w; r.v = fresnel_schlick(R0.v, c); return r; }

// Conductor Fresnel, Shirley 1985 in real arithmetic (fresnel_complex.h:31-50)
template<class T> BBMCU_D T fresnel_complex(const T& n, const T& k, float c)
{
  float c2 = c*c, s2 = 1.0f - c2;
  T n2 = n*n, k2 = k*k;
  T temp = n2 - k2 - s2;
  T a2b2 = m_safe_sqrt(temp*temp + 4.0f*n2*k2);
  T a = m_safe_sqrt(0.5f * (a2b2 + temp));
  T a2c = 2.0f*a*c;
  T Rs = (a2b2 - a2c + c2) / (a2b2 + a2c + c2);
  T Rp = Rs * (c2*a2b2 - (a2c - s2)*s2) / (c2*a2b2 + (a2c + s2)*s2);
  return 0.5f*(Rs + Rp);
}

// Fresnel policies: NA attribute floats; eval returns T (scalar) or Spec<T> (spectral).
struct FresnelCookIor      { static constexpr int NA = 1; template<class T> BBMCU_D static T eval(const T* a, float c) { return fresnel_cook(a[0], c); }
                             template<class T> BBMCU_D static T evalq(const T* a, float c) { if constexpr (std::is_same<T, float>::value) return fresnel_cook_q(a[0], c); else return fresnel_cook(a[0], c); } };
struct FresnelSchlickR0    { static constexpr int NA = 1; template<class T> BBMCU_D static T eval(const T* a, float c) { return fresnel_schlick(a[0], c); }  template<class T> BBMCU_D static T evalq(const T* a, float c) { return eval<T>(a, c); } };
struct FresnelComplexScalar{ static constexpr int NA = 2; template<class T> BBMCU_D static T eval(const T* a, float c) { return fresnel_complex(a[0], a[1], c); }  template<class T> BBMCU_D static T evalq(const T* a, float c) { return eval<T>(a, c); } };
struct FresnelComplexRGB   { static constexpr int NA = 6; template<class T> BBMCU_D static Spec<T> eval(const T* a, float c) {
    return Spec<T>(fresnel_complex(a[0], a[3], c), fresnel_complex(a[1], a[4], c), fresnel_complex(a[2], a[5], c)); }
                             template<class T> BBMCU_D static Spec<T> evalq(const T* a, float c) { return eval<T>(a, c); } };
struct FresnelSchlickRGB   { static constexpr int NA = 3; template<class T> BBMCU_D static Spec<T> eval(const T* a, float c) {
    return Spec<T>(fresnel_schlick(a[0], c), fresnel_schlick(a[1], c), fresnel_schlick(a[2], c)); }
                             template<class T> BBMCU_D static Spec<T> evalq(const T* a, float c) { return eval<T>(a, c); } };
// fresnel::bagher: schlick(F0) - F1*cos   (bagher.h:47-50); attribute = [F0 rgb, F1 rgb]
struct FresnelBagher       { static constexpr int NA = 6; template<class T> BBMCU_D static Spec<T> eval(const T* a, float c) {
    return Spec<T>(fresnel_schlick(a[0], c) - a[3]*c, fresnel_schlick(a[1], c) - a[4]*c, fresnel_schlick(a[2], c) - a[5]*c); }
                             template<class T> BBMCU_D static Spec<T> evalq(const T* a, float c) { return eval<T>(a, c); } };

// scalar-or-spectrum helpers
template<class T> BBMCU_D Spec<T> to_spec(const T& a) { return Spec<T>(a); }
template<class T> BBMCU_D Spec<T> to_spec(const Spec<T>& a) { return a; }

// the rational fit shared by the Beckmann and Phong G1 (beckmann.h:196, phong.h:127), evaluated in
// double by the reference (double literals) and rounded to float.
BBMCU_D float smith_rational(float a)
{
  if(!(a < 1.6f)) return 1.0f;
  return q_div(3.535f*a + 2.181f*a*a, 1.0f + 2.276f*a + 2.577f*a*a);       // positive terms only: no cancellation
}
template<int N> BBMCU_D Dual<N> smith_rational(const Dual<N>& a)
{
  if(!(a.v < 1.6f)) return Dual<N>(1.0f);
  float x = a.v;
  float num = 3.535f*x + 2.181f*x*x, den = 1.0f + 2.276f*x + 2.577f*x*x;
  float dnum = 3.535f + 2.0f*2.181f*x, dden = 2.276f + 2.0f*2.577f*x;
  return chain(a, smith_rational(x), (dnum*den - num*dden)/(den*den));
}

// =============================================================================================
// Normal distributions.  Interface:
//   NA                      attribute floats
//   D<T>(h, a)              eval (h.z > 0 checked here)           -> T or Spec<T>
//   G1<T>(v, m, a)          monodirectional shadowing               -> T or Spec<T>
//   sample(view, xi, a)     microfacet normal (xi range already checked by the caller too)
//   pdf(view, m, a)
// =============================================================================================
template<bool ANISO> struct Alpha2 { template<class T> BBMCU_D static void get(const T* a, T& ax, T& ay) { ax = a[0]; ay = ANISO ? a[1] : a[0]; } };

BBMCU_D bool xi_valid(f2 xi) { return (xi.x >= 0.0f) && (xi.y >= 0.0f) && (xi.x <= 1.0f) && (xi.y <= 1.0f); }

// ---- Beckmann (beckmann.h) -------------------------------------------------------------------
template<bool ANISO, bool NORMALIZE, bool SAMPLE_VISIBLE = true>
struct NdfBeckmann
{
  static constexpr int NA = ANISO ? 2 : 1;
  template<class T> BBMCU_D static T D(f3 h, const T* a)
  {
    if(!(h.z > 0.0f)) return T(0.0f);
    T ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    float c2 = h.z*h.z;
    T d;
    if constexpr (std::is_same<T, float>::value)
    {
      // the exponent keeps the reference's IEEE operations (exp amplifies its relative error by |exponent|, 10^2..10^4
      // for fitted roughness 0.003-0.02); the quotient outside is a final value.  The three IEEE divisions go through
      // the unguarded fast path when roughness (uniform) and cos^2 (almost always) are in range: same bits, no slow-path
      // regions.  |h.x|, |h.y| <= 1 over alpha > 1e-14 stay below 1e14, their squares below 2e28, the exponent below 2e35.
      float e;
      if((ax > 1e-14f) && (ax < 1e30f) && (ay > 1e-14f) && (ay < 1e30f) && (c2 > 1e-7f))
      {
        const float sx = ieee_div_raw(h.x, ax), sy = ieee_div_raw(h.y, ay);
        e = ieee_div_raw(-(sx*sx + sy*sy), c2);
      }
      else { const float sx = h.x/ax, sy = h.y/ay; e = -(sx*sx + sy*sy) / c2; }
      d = q_div(m_exp(e), ax*ay*c2*c2);
    }
    else
    {
      T sx = h.x/ax, sy = h.y/ay;
      d = m_exp(-(sx*sx + sy*sy) / c2) / (ax*ay*c2*c2);
    }
    if(NORMALIZE) d = d * kInvPi;
    return d;
  }
  template<class T> BBMCU_D static T G1(f3 v, f3 m, const T* a)
  {
    if(!((v.z > 0.0f) && (dot(v, m) > 0.0f))) return T(0.0f);
    T aa;
    if(ANISO) { T sx = v.x*a[0], sy = v.y*a[1]; aa = m_rsqrt((sx*sx + sy*sy) / (v.z*v.z)); }
    else if constexpr (std::is_same<T, float>::value) aa = q_rcp(a[0] * q_tanTheta(v));
    else aa = m_rcp(a[0] * tanTheta(v));
    return smith_rational(aa);
  }
  BBMCU_D static float pdf(f3 view, f3 m, const float* a)
  {
    if(!(m.z > 0.0f)) return 0.0f;
    float p = D<float>(m, a);
    if(SAMPLE_VISIBLE) p *= q_div(G1<float>(view, m, a) * fabsf(dot(view, m)), view.z);
    else p *= m.z;
    return (p > 0.0f) ? p : 0.0f;
  }
  BBMCU_D static f3 sample(f3 view, f2 xi, const float* a)
  {
    if(!xi_valid(xi)) return make_f3(0, 0, 0);
    float ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    if(SAMPLE_VISIBLE)
    {
      // Jakob 2014 (beckmann.h:87-115); the double/float mix follows the reference's promotions.  erff / expf / logf
      // are the host libm's (bbmcu_libm.cuh): the Newton iteration ends in erfinv near +-1, where one differing bit of
      // any of them moves the sampled slope by up to 1e4 ulp.
      f3 vs = normalize(make_f3(view.x*ax, view.y*ay, view.z));
      float tanT = tanTheta(vs);
      float maxval = glibc_erff(1.0f / tanT);
      float x0 = clampf(xi.x, 1e-5f, (float)(1.0 - 10e-6));
      float x1 = clampf(xi.y, 1e-5f, (float)(1.0 - 10e-6));
      float x = maxval - (maxval + 1.0f) * glibc_erff(sqrtf(-glibc_logf(x0)));
      float gauss = kInvSqrtPi * tanT * glibc_expf(-(vs.z*vs.z));
      x0 = (float)((double)x0 * (1.0 + (double)maxval + (double)gauss));
#pragma unroll
      for(int i=0; i < 3; ++i)
      {
        float slope = (float)erfinv_ref(x);
        float g = kInvSqrtPi * tanT * glibc_expf(-slope*slope);
        float value = (float)(1.0 + (double)x + (double)g - (double)x0);
        float deriv = (float)(1.0 - (double)(slope*tanT));
        x -= value / deriv;
      }
      float s0 = 0.0f, s1 = 0.0f;
      if(x > -1.0f && x < 1.0f) { s0 = (float)erfinv_ref(x); s1 = (float)erfinv_ref((float)(2.0*(double)x1 - 1.0)); }
      f2 cs = cossinPhi(vs);
      float ux = (cs.x*s0 + (-cs.y)*s1) * ax;      // rotation2d(cos,sin) * slope, then unstretch
      float uy = (cs.y*s0 + cs.x*s1) * ay;
      return normalize(make_f3(-ux, -uy, 1.0f));
    }
    else
    {
      // Walter 2007 with anisotropic extension (beckmann.h:117-135)
      float cp = cosf(kTwoPi * xi.x), sp = sinf(kTwoPi * xi.x);
      float nrm;
      if(ANISO) { cp *= ax; sp *= ay; nrm = cp*cp + sp*sp; float r = 1.0f/sqrtf(nrm); cp *= r; sp *= r; }
      else nrm = ax*ax;
      float cosT = (float)(1.0 / sqrt(1.0 - (double)(nrm*logf(xi.y))));
      float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
      return make_f3(cp*sinT, sp*sinT, cosT);
    }
  }
};

// ---- GGX (ggx.h) -----------------------------------------------------------------------------
template<bool ANISO>
struct NdfGGX
{
  static constexpr int NA = ANISO ? 2 : 1;
  static constexpr bool kQuickHalfway = true;      // D is algebraic in h: a 2-ulp half vector moves it by < 1e-6
  template<class T> BBMCU_D static T D(f3 h, const T* a)
  {
    if(!(h.z > 0.0f)) return T(0.0f);
    T ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    if constexpr (std::is_same<T, float>::value)
    {
      float sx = h.x * q_rcp(ax), sy = ANISO ? h.y * q_rcp(ay) : h.y * q_rcp(ax);
      float t = (sx*sx + sy*sy) + h.z*h.z;
      return q_rcp(kPi * (ax*ay) * (t*t));
    }
    else
    {
      T sx = h.x/ax, sy = h.y/ay;
      T t = (sx*sx + sy*sy) + h.z*h.z;
      return m_rcp(kPi * (ax*ay) * (t*t));
    }
  }
  template<class T> BBMCU_D static T G1(f3 v, f3 m, const T* a)
  {
    if(!((v.z > 0.0f) && (dot(v, m) > 0.0f))) return T(0.0f);
    T ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    T r2 = ax*ay;
    if constexpr (std::is_same<T, float>::value) return g1_tail(r2 * q_div(sinTheta2(v), v.z*v.z));
    else return g1_tail(r2 * tanTheta2(v));
  }
  // 2.0 / (1.0 + sqrt(1.0 + x)): the reference evaluates this in double (ggx.h:184-187); a final value, so quick ops
  BBMCU_D static float g1_tail(float x) { return 2.0f * q_rcp(1.0f + q_sqrt(1.0f + x)); }
  template<int N> BBMCU_D static Dual<N> g1_tail(const Dual<N>& x) { float s = sqrtf(1.0f + x.v); float den = 1.0f + s; return chain(x, 2.0f/den, -1.0f/(den*den*s)); }

  BBMCU_D static float pdf(f3 view, f3 m, const float* a)
  {
    if(!(m.z > 0.0f)) return 0.0f;
    return pdf_visible(D<float>(m, a), G1<float>(view, m, a), fabsf(dot(view, m)), view.z);
  }
  // the visible-normal pdf from an already evaluated D(m) and G1(view, m) (the fused eval + pdf path has both)
  static constexpr bool kHasPdfVisible = true;
  BBMCU_D static float pdf_visible(float Dm, float G1v, float abs_vm, float vz) { float p = Dm * q_div(G1v * abs_vm, vz); return (p > 0.0f) ? p : 0.0f; }
  // Heitz 2017 visible-normal sampling (ggx.h:86-107).  Every step up to the normal n reproduces the reference's
  // roundings (no FMA, IEEE sqrt / divide, its float-double mix): the two differences in it - (xi1 - a)/(1 - a) for
  // grazing views and 1 - P1^2 - P2^2 for xi0 -> 1 - cancel, so a last-bit change upstream would move the sampled
  // direction by far more than 1e-5.  Spots where a float operation provably returns the reference's double result
  // rounded to float use the float operation:
  //   * cross(vs, z) and cross(T1, vs) with their exact zeros removed;
  //   * a = (float)(1.0 / (1.0 + (double)vs.z)): 1 + vs.z = hi + lo exactly (Fast2Sum), one Newton step on the IEEE
  //     reciprocal of hi with fused residuals gives the correctly rounded quotient;
  //   * (double)(xi1 / a) * pi and 1.0 * r * sin(phi): products of two floats are exact in double, so the float
  //     product is the same single rounding;
  //   * cos / sin(phi): the host libm's values (bbmcu_libm.cuh), one shared quadrant reduction.
  // Only the final normalisation of the stretched normal is a quick op (nothing cancels after it).
  BBMCU_D static f3 sample(f3 view, f2 xi, const float* a) { return xi_valid(xi) ? sample_unchecked(view, xi, a) : make_f3(0, 0, 0); }
  static constexpr bool kHasSampleUnchecked = true;
  // xi already validated by the caller
  BBMCU_D static f3 sample_unchecked(f3 view, f2 xi, const float* a)
  {
    float ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    // (*_nr / *_raw: IEEE results through the math library's own fast path, see bbmcu_math.cuh.  Ranges: the stretched
    // view is range-tested once in normalize_nr; vs.x^2 + vs.y^2 is in (1.19e-7, 1] inside the branch; hi is in [1, 2],
    // a in [0.5, 1]; xi0 in [0, 1] needs the lower test only)
    // Branch-free: both cases of the reference's two select() calls run through ONE float-pair evaluation (a warp holds
    // both cases almost always, so a branch would execute both sides anyway), and the unguarded *_raw results of
    // out-of-range operands are selected away.
    f3 vs = normalize_nr(make_f3(view.x*ax, view.y*ay, view.z));
    const bool tilted = vs.z < 0.99999988079071044921875f;
    const float rr = ieee_rcp_raw(ieee_sqrt_raw(vs.y*vs.y + vs.x*vs.x));          // garbage for vs = +z: not selected
    const f3 T1 = make_f3(tilted ? vs.y*rr : 1.0f, tilted ? (-vs.x)*rr : 0.0f, 0.0f);
    const f3 T2 = make_f3(T1.y*vs.z, -(T1.x*vs.z), T1.x*vs.y - T1.y*vs.x);
    const float hi = 1.0f + vs.z, lo = vs.z - (hi - 1.0f);
    const float r0 = ieee_rcp_raw(hi);
    const float aa = fmaf(r0, fmaf(-lo, r0, fmaf(-hi, r0, 1.0f)), r0);
    // sqrt(xi0); below 1e-30 the root (< 1e-15) is taken as 0: it moves the normal by less than 1e-15
    const float r = sel_gt(xi.x, 1e-30f, ieee_sqrt_raw(xi.x));
    const bool lower = xi.y < aa;
    // phi = (float)((double)(xi1 / a) * pi)  or  (float)((1.0 + (double)(xi1 - a) / (1.0 - (double)a)) * pi):
    // quotient, 1 + q and the product are carried as hi + lo pairs (~2^-46) and rounded once.  In the first case the pair
    // is (xi1 / a, 0) and the last line returns the correctly rounded float product, i.e. the reference's value exactly.
    // The second differs from the double evaluation only where the double path itself rounds twice across a float tie
    // (1.6e-5 of a grazing-biased test set, ~1e-7 of uniform inputs; one ulp of phi each).
    const float num = lower ? xi.y : xi.y - aa, den = lower ? aa : 1.0f - aa, base = lower ? 0.0f : 1.0f;
    const float qh = ieee_div_raw(num, den);
    const float ql = lower ? 0.0f : q_div(fmaf(-qh, den, num), den);
    const float sh = base + qh, sl = ((base - sh) + qh) + ql;
    const float ph = sh * kPi;
    const float phi = ph + fmaf(sl, kPi, fmaf(sh, kPi, -ph));
    float cp, sp; glibc_sincosf_both(phi, sp, cp);
    const float P1 = r*cp;
    // P2 = (float)(1.0 * r * sp)  or  (float)((double)vs.z * r * sp): the same product as a pair, rounded once
    const float zz = lower ? 1.0f : vs.z;
    const float hi2 = zz * r, lo2 = fmaf(zz, r, -hi2);
    const float t2 = hi2 * sp;
    const float P2 = t2 + fmaf(lo2, sp, fmaf(hi2, sp, -t2));
    // (float)sqrt(max(1.0 - (double)(P1*P1) - (double)(P2*P2), 0)): both squares are float products, the differences are
    // exact in double; here 1 - p1 - p2 is carried as a float pair (Fast2Sum twice: 1 >= p1, and 1 - p1 >= p2/2), the root
    // is the IEEE float root of the high part plus one Newton correction with the exact residual, rounded once (differs
    // from the double evaluation only within 2^-22 ulp of a rounding tie: ~1e-7 of samples, by one ulp of w).  Below
    // 1e-30 the root is taken as 0.
    const float p1 = P1*P1, p2 = P2*P2;
    const float t = 1.0f - p1, e1 = (1.0f - t) - p1;
    const float u = t - p2, e2 = (t - u) - p2;
    const float lo_s = e1 + e2;
    const float s_hi = u + lo_s, s_lo = lo_s - (s_hi - u);
    const float q = ieee_sqrt_raw(s_hi);
    const float w = sel_gt(s_hi, 1e-30f, fmaf(fmaf(-q, q, s_hi) + s_lo, 0.5f * q_rcp(q), q));
    // n = P1 T1 + P2 T2 + w vs feeds only the final normalisation: fused multiply-adds (1e-7 of |n| = 1, nothing cancels after)
    const f3 n = make_f3(fmaf(T1.x, P1, fmaf(T2.x, P2, vs.x*w)), fmaf(T1.y, P1, fmaf(T2.y, P2, vs.y*w)), fmaf(T2.z, P2, vs.z*w));
    return q_normalize(make_f3(n.x*ax, n.y*ay, fmaxf(0.0f, n.z)));
  }
};

// ---- Phong NDF (ndf/phong.h) -----------------------------------------------------------------
struct NdfPhong
{
  static constexpr int NA = 1;
  template<class T> BBMCU_D static T D(f3 h, const T* a)
  {
    if(!(h.z > 0.0f)) return T(0.0f);
    T nrm = (a[0] + 2.0f) / kTwoPi;
    return m_pow(h.z, a[0]) * nrm;
  }
  template<class T> BBMCU_D static T G1(f3 v, f3 m, const T* a)
  {
    if(!((v.z > 0.0f) && (dot(v, m) > 0.0f))) return T(0.0f);
    T aa = m_sqrt(0.5f*a[0] + 1.0f) / tanTheta(v);
    return smith_rational(aa);
  }
  BBMCU_D static float pdf(f3, f3 m, const float* a)
  {
    if(!(m.z > 0.0f)) return 0.0f;
    return D<float>(m, a) * fabsf(m.z);
  }
  BBMCU_D static f3 sample(f3, f2 xi, const float* a)
  {
    if(!xi_valid(xi)) return make_f3(0, 0, 0);
    float cosT = (float)pow((double)xi.x, 1.0 / (double)(a[0] + 2.0f));
    float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
    float ph = xi.y * kTwoPi;
    return make_f3(cosf(ph)*sinT, sinf(ph)*sinT, cosT);
  }
};

// ---- Low NDF (ndf/low.h): (1 + B (1 - h.z))^-C, unnormalised; G1 = 1 ---------------------------
struct NdfLow
{
  static constexpr int NA = 2;
  // the reference evaluates the base in double (1.0 literals) and calls the double pow
  BBMCU_D static float D_(f3 h, float B, float C) { return (float)pow(1.0 + (double)B*(1.0 - (double)h.z), -(double)C); }
  template<class T> BBMCU_D static T D(f3 h, const T* a);
  template<class T> BBMCU_D static T G1(f3, f3, const T*) { return T(1.0f); }
  BBMCU_D static float pdf(f3, f3 m, const float* a)
  {
    if(!(m.z > 0.0f)) return 0.0f;
    float B = a[0], C = a[1];
    float nrm = (fabsf(C - 1.0f) < kEps) ? (float)(1.0f / log(1.0 + (double)B))
                                         : (float)(((double)C - 1.0) / (1.0 - pow(1.0 + (double)B, 1.0 - (double)C)));
    float p = D_(m, B, C) * B * ((0.5f*kInvPi) * nrm);
    return (p > 0.0f) ? p : 0.0f;
  }
  BBMCU_D static f3 sample(f3, f2 xi, const float* a)
  {
    if(!xi_valid(xi)) return make_f3(0, 0, 0);
    float B = a[0], C = a[1];
    float term = (fabsf(C - 1.0f) < kEps)
               ? (float)exp((double)xi.x * log(1.0 + (double)B))
               : (float)pow(1.0 + (double)xi.x * (pow(1.0 + (double)B, 1.0 - (double)C) - 1.0), -1.0 / ((double)C - 1.0));
    float cosT = (float)((1.0 + (double)B - (double)term) / (double)B);
    float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
    float ph = xi.y * kTwoPi;
    return make_f3(cosf(ph)*sinT, sinf(ph)*sinT, cosT);
  }
};
template<> BBMCU_D float NdfLow::D<float>(f3 h, const float* a) { return (h.z > 0.0f) ? D_(h, a[0], a[1]) : 0.0f; }
template<class T> BBMCU_D T NdfLow::D(f3 h, const T* a)
{
  if(!(h.z > 0.0f)) return T(0.0f);
  T base = 1.0f + a[0]*(1.0f - h.z);
  T r = m_pow(base, -a[1]);
  r.v = D_(h, a[0].v, a[1].v);
  return r;
}

// ---- Student-t NDF (ndf/studentt.h) -----------------------------------------------------------
template<bool ANISO>
struct NdfStudentT
{
  static constexpr int NA = ANISO ? 3 : 2;     // roughness[1|2], gamma
  template<class T> BBMCU_D static T D(f3 h, const T* a)
  {
    if(!(h.z > 0.0f)) return T(0.0f);
    T ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    const T& gamma = a[NA-1];
    float c2 = h.z*h.z;
    T nrm = kPi * (ax*ay) * (c2*c2);
    T sx = h.x/ax, sy = h.y/ay;
    T den = m_pow(1.0f + (sx*sx + sy*sy) / ((gamma - 1.0f)*c2), gamma);
    return m_rcp(nrm * den);
  }
  BBMCU_D static float ratfit(float x, float n0, float n1, float n2, float n3, float d0, float d1, float d2, float d3)
  {
    // the reference evaluates these rational fits with double literals and rounds num/den to float
    double x1 = x, x2 = (double)(x*x), x3 = (double)((x*x)*x);
    float num = (float)(n0 + n1*x1 + n2*x2 + n3*x3), den = (float)(d0 + d1*x1 + d2*x2 + d3*x3);
    return num / den;
  }
  template<int N> BBMCU_D static Dual<N> ratfit(const Dual<N>& x, float n0, float n1, float n2, float n3, float d0, float d1, float d2, float d3)
  {
    float v = x.v, num = n0 + v*(n1 + v*(n2 + v*n3)), den = d0 + v*(d1 + v*(d2 + v*d3));
    float dn = n1 + v*(2.0f*n2 + v*3.0f*n3), dd = d1 + v*(2.0f*d2 + v*3.0f*d3);
    return chain(x, ratfit(v, n0, n1, n2, n3, d0, d1, d2, d3), (dn*den - num*dd)/(den*den));
  }
  // The factors of G1 that depend on gamma alone (two rational fits, a powf, two tgammaf, a square root): with the
  // parameters uniform over a launch the element-wise eval kernels form them once per thread (k_foreach4, Op::group_pre)
  // instead of twice per evaluation.  G1<float> goes through the same two functions, so both routes give the same bits.
  static constexpr bool kHasPre = true;
  struct Pre { float F22, F23, S1s, tgr, sq; };
  BBMCU_D static Pre pre(const float* a)
  {
    const float gamma = a[NA-1];
    Pre q;
    q.F22 = ratfit(gamma, 14.402f, -27.145f, 20.574f, -2.745f, -30.612f, 86.567f, -84.341f, 29.938f);
    q.F23 = ratfit(gamma, -129.404f, 324.987f, -299.305f, 93.268f, -92.609f, 256.006f, -245.663f, 86.064f);
    q.S1s = m_pow(gamma - 1.0f, gamma) / (2.0f*gamma - 3.0f);
    q.tgr = m_tgamma(gamma - 0.5f) / m_tgamma(gamma) * kInvSqrtPi;
    q.sq = m_sqrt(gamma - 1.0f);
    return q;
  }
  BBMCU_D static float G1_core(f3 v, const float* a, const Pre& q)     // after the two early returns of G1
  {
    float ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    const float gamma = a[NA-1];
    float sx = v.x*ax, sy = v.y*ay;
    float z = v.z * m_rsqrt(sx*sx + sy*sy);
    float S1 = m_pow((gamma - 1.0f) + z*z, 1.5f - gamma) / z;
    float F21 = ratfit(z, 0.0f, 1.066f, 2.655f, 4.892f, 1.038f, 2.969f, 4.305f, 4.418f);
    float F24 = ratfit(z, 6.537f, 6.074f, -0.623f, 5.223f, 6.538f, 6.103f, -3.218f, 6.347f);
    float S2 = F21 * (q.F22 + q.F23*F24);
    float lambda = q.tgr * (q.S1s*S1 + q.sq*S2) - 0.5f;
    return 1.0f / (1.0f + lambda);
  }
  BBMCU_D static float G1_pre(f3 v, f3 m, const float* a, const Pre& q)
  {
    if(!((v.z > 0.0f) && (dot(v, m) > 0.0f))) return 0.0f;
    if(!(v.z < 0.99999988079071044921875f)) return 1.0f;
    return G1_core(v, a, q);
  }
  template<class T> BBMCU_D static T G1(f3 v, f3 m, const T* a)
  {
    if(!((v.z > 0.0f) && (dot(v, m) > 0.0f))) return T(0.0f);
    if(!(v.z < 0.99999988079071044921875f)) return T(1.0f);
    if constexpr (std::is_same<T, float>::value) return G1_core(v, a, pre(a));
    else
    {
    T ax, ay; Alpha2<ANISO>::get(a, ax, ay);
    const T& gamma = a[NA-1];
    T sx = v.x*ax, sy = v.y*ay;
    T z = v.z * m_rsqrt(sx*sx + sy*sy);
    T S1 = m_pow((gamma - 1.0f) + z*z, 1.5f - gamma) / z;
    T F21 = ratfit(z, 0.0f, 1.066f, 2.655f, 4.892f, 1.038f, 2.969f, 4.305f, 4.418f);
    T F22 = ratfit(gamma, 14.402f, -27.145f, 20.574f, -2.745f, -30.612f, 86.567f, -84.341f, 29.938f);
    T F23 = ratfit(gamma, -129.404f, 324.987f, -299.305f, 93.268f, -92.609f, 256.006f, -245.663f, 86.064f);
    T F24 = ratfit(z, 6.537f, 6.074f, -0.623f, 5.223f, 6.538f, 6.103f, -3.218f, 6.347f);
    T S2 = F21 * (F22 + F23*F24);
    T S1s = m_pow(gamma - 1.0f, gamma) / (2.0f*gamma - 3.0f);
    T lambda = m_tgamma(gamma - 0.5f) / m_tgamma(gamma) * kInvSqrtPi * (S1s*S1 + m_sqrt(gamma - 1.0f)*S2) - 0.5f;
    return 1.0f / (1.0f + lambda);
    }
  }
  BBMCU_D static float pdf(f3, f3 m, const float* a)
  {
    if(!(m.z > 0.0f)) return 0.0f;
    float p = D<float>(m, a) * m.z;
    return (p > 0.0f) ? p : 0.0f;
  }
  BBMCU_D static f3 sample(f3, f2 xi, const float* a)
  {
    if(!xi_valid(xi)) return make_f3(0, 0, 0);
    float gamma = a[NA-1];
    float cp = cosf(kTwoPi * xi.x), sp = sinf(kTwoPi * xi.x);
    float nrm;
    if(ANISO) {
      float qx = cp/a[0], qy = sp/a[1];
      nrm = 1.0f / (qx*qx + qy*qy);
      cp *= a[0]; sp *= a[1];
      float r = 1.0f/sqrtf(cp*cp + sp*sp); cp *= r; sp *= r;
    }
    else nrm = a[0]*a[0];
    float tan2 = (float)((pow((double)xi.y, 1.0 / (1.0 - (double)gamma)) - 1.0) * (double)(gamma - 1.0f) * (double)nrm);
    float cosT = (float)(1.0 / sqrt(1.0 + (double)tan2));
    float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
    return make_f3(cp*sinT, sp*sinT, cosT);
  }
};

// ---- Shifted Gamma Distribution (ndf/sgd.h), spectral.  Attribute block in reflection order:
//      K[3] Lambda[3] c[3] theta0[3] k[3] (Dependent) then alpha[3] p[3] ----------------------------
struct NdfSGD
{
  static constexpr int NA = 21;
  template<class T> BBMCU_D static T D1(float tan2, float c4pi, const T& alpha, const T& p, const T& K)
  {
    T temp = alpha + tan2/alpha;
    T den = m_pow(temp, p);
    T P22 = (den > kEps) ? m_exp(-temp)/den : T(0.0f);
    return P22 / c4pi * K;
  }
  template<class T> BBMCU_D static Spec<T> D(f3 h, const T* a)
  {
    if(!(h.z > 0.0f)) return Spec<T>(T(0.0f));
    float tan2 = tanTheta2(h);
    float c2 = h.z*h.z;
    float c4pi = kPi * (float)((double)c2*(double)c2);   // Pi * pow(cos, 4.0)
    return Spec<T>(D1(tan2, c4pi, a[15], a[18], a[0]), D1(tan2, c4pi, a[16], a[19], a[1]), D1(tan2, c4pi, a[17], a[20], a[2]));
  }
  // 1 + Lambda (1 - exp(c (theta - theta0)^k)): the 1 - exp(.) cancels, so exp/pow are evaluated in double
  // and rounded to float (what glibc's correctly rounded expf/powf return) before the float arithmetic
  BBMCU_D static float G11(float theta, float Lambda, float c, float theta0, float k)
  {
    if(!(theta > theta0)) return 1.0f;
    float pw = (float)pow((double)(theta - theta0), (double)k);
    float ex = (float)exp((double)(c*pw));
    return 1.0f + Lambda*(1.0f - ex);
  }
  template<int N> BBMCU_D static Dual<N> G11(float theta, const Dual<N>& Lambda, const Dual<N>& c, const Dual<N>& theta0, const Dual<N>& k)
  {
    if(!(theta > theta0.v)) return Dual<N>(1.0f);
    Dual<N> r = 1.0f + Lambda*(1.0f - m_exp(c*m_pow(theta - theta0, k)));
    r.v = G11(theta, Lambda.v, c.v, theta0.v, k.v);
    return r;
  }
  template<class T> BBMCU_D static Spec<T> G1(f3 v, f3 m, const T* a)
  {
    if(!((v.z > 0.0f) && (dot(v, m) > 0.0f))) return Spec<T>(T(0.0f));
    float theta = sph_theta(v);
    return Spec<T>(G11(theta, a[3], a[6], a[9], a[12]), G11(theta, a[4], a[7], a[10], a[13]), G11(theta, a[5], a[8], a[11], a[14]));
  }
  BBMCU_D static float avg_alpha(const float* a) { return ((a[15] + a[16]) + a[17]) / 3.0f; }
  BBMCU_D static float pdf(f3 view, f3 m, const float* a) { float al = avg_alpha(a); return NdfGGX<false>::pdf(view, m, &al); }
  BBMCU_D static f3 sample(f3 view, f2 xi, const float* a) { float al = avg_alpha(a); return NdfGGX<false>::sample(view, xi, &al); }
};

// =============================================================================================
// Joint masking-shadowing
// =============================================================================================
BBMCU_D bool g_mask(f3 in, f3 out, f3 m) { return (dot(in, m) > 0.0f) && (dot(out, m) > 0.0f); }

struct GVGroove {
  template<class NDF, class T> BBMCU_D static auto eval(f3 in, f3 out, f3 m, const T* a) -> decltype(NDF::template G1<T>(in, m, a))
  {
    using R = decltype(NDF::template G1<T>(in, m, a));
    if(!g_mask(in, out, m)) return R(T(0.0f));
    // min(1.0, min(2.0*m.z*in.z/(in.m), 2.0*m.z*out.z/(out.m))): double in the reference (vgroove.h:41-46), a final
    // value here (products and one quotient)
    float gi = q_div(2.0f*m.z*in.z, dot(in, m));
    float go = q_div(2.0f*m.z*out.z, dot(out, m));
    return R(T(fminf(1.0f, fminf(gi, go))));
  }
};
struct GUncorrelated {
  template<class NDF, class T> BBMCU_D static auto eval(f3 in, f3 out, f3 m, const T* a) -> decltype(NDF::template G1<T>(in, m, a))
  {
    using R = decltype(NDF::template G1<T>(in, m, a));
    if(!g_mask(in, out, m)) return R(T(0.0f));
    return NDF::template G1<T>(in, m, a) * NDF::template G1<T>(out, m, a);
  }
};
struct GHeightCorrelated {
  template<class NDF, class T> BBMCU_D static T eval(f3 in, f3 out, f3 m, const T* a)
  {
    if(!g_mask(in, out, m)) return T(0.0f);
    T gi = NDF::template G1<T>(in, m, a), go = NDF::template G1<T>(out, m, a);
    T gio = gi*go;
    T den = gi + go - gio;
    return (den > kEps) ? gio/den : T(0.0f);
  }
};
struct GVanGinneken {
  template<class NDF, class T> BBMCU_D static T eval(f3 in, f3 out, f3 m, const T* a)
  {
    if(!g_mask(in, out, m)) return T(0.0f);
    float phi = fabsf(sph_phi(in) - sph_phi(out));                 // un-wrapped, as the reference (vanginneken.h:50)
    float lambda = (float)(4.41*(double)phi / (4.41*(double)phi + 1.0));
    T gi = NDF::template G1<T>(in, m, a), go = NDF::template G1<T>(out, m, a);
    T gio = gi*go;
    T den = m_max(gi, go) + lambda*(m_min(gi, go) - gio);
    return (den > kEps) ? gio/den : T(0.0f);
  }
};

// =============================================================================================
// microfacet<NDF, G, F, NormalizationFactor>, optionally scaled by a leading RGB attribute.
// Attribute block: [scale rgb (if SCALED)] [NDF attributes] [Fresnel attributes]
// NORM: 0 Unnormalized (1.0), 1 Walter (4.0), 2 Cook (pi as double)   (microfacet.h:31-36)
// =============================================================================================
// half vector for eval / pdf: IEEE normalisation unless the NDF declares itself insensitive (kQuickHalfway)
template<class NDF, class = void> struct QuickHalfway { static constexpr bool value = false; };
template<class NDF> struct QuickHalfway<NDF, typename std::enable_if<NDF::kQuickHalfway>::type> { static constexpr bool value = true; };
template<class NDF, class = void> struct HasPdfVisible { static constexpr bool value = false; };
template<class NDF> struct HasPdfVisible<NDF, typename std::enable_if<NDF::kHasPdfVisible>::type> { static constexpr bool value = true; };
template<class NDF, class = void> struct HasSampleUnchecked { static constexpr bool value = false; };
template<class NDF> struct HasSampleUnchecked<NDF, typename std::enable_if<NDF::kHasSampleUnchecked>::type> { static constexpr bool value = true; };
template<class NDF> BBMCU_D f3 quick_halfway(f3 a, f3 b) { if constexpr (QuickHalfway<NDF>::value) return q_normalize(a + b); else return halfway(a, b); }

template<class NDF, class = void> struct NdfHasPre { static constexpr bool value = false; };
template<class NDF> struct NdfHasPre<NDF, typename std::enable_if<NDF::kHasPre>::type> { static constexpr bool value = true; };

template<class NDF, class = void> struct NdfFusedMinBlocks { static constexpr int value = 0; };
template<class NDF> struct NdfFusedMinBlocks<NDF, typename std::enable_if<(NDF::kFusedMinBlocks > 0)>::type> { static constexpr int value = NDF::kFusedMinBlocks; };

template<class NDF, class G, class F, int NORM, bool SCALED>
struct Microfacet
{
  static constexpr int kLaunchMinBlocksFused = NdfFusedMinBlocks<NDF>::value;      // 0: the default of bbmcu_bsdf.cuh
  static constexpr int SCALE = SCALED ? 0 : -1;
  static constexpr int OFF_NDF = SCALED ? 3 : 0;
  static constexpr int OFF_F = OFF_NDF + NDF::NA;
  static constexpr int NA = OFF_F + F::NA;
  BBMCU_HD static constexpr double norm() { return NORM == 0 ? 1.0 : (NORM == 1 ? 4.0 : kPiD); }

  // direction-only part of eval, hoisted out of the parameter loop of the batched loss kernels
  struct Geom { f3 h; float inh, outh; };
  BBMCU_D static Geom geom(f3 in, f3 out)
  {
    Geom g; g.h = make_f3(0, 0, 1); g.inh = 0.0f; g.outh = 0.0f;
    if((in.z > 0.0f) && (out.z > 0.0f)) { g.h = halfway(in, out); g.inh = dot(in, g.h); g.outh = dot(out, g.h); }
    return g;
  }
  // three identical channels before the scale whenever D, G and F are scalars
  static constexpr bool kGrayUnscaled = std::is_same<decltype(NDF::template D<float>(f3(), (const float*)nullptr)), float>::value
                                     && std::is_same<decltype(F::template eval<float>((const float*)nullptr, 0.0f)), float>::value;
  template<class T> BBMCU_D static Spec<T> eval_unscaled_g(const Geom& g, f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    auto D = NDF::template D<T>(g.h, a + OFF_NDF);
    auto Gv = G::template eval<NDF, T>(in, out, g.h, a + OFF_NDF);
    auto Fv = F::template evalq<T>(a + OFF_F, 0.5f*(g.inh + g.outh));
    Spec<T> dgf = to_spec(D) * to_spec(Gv) * to_spec(Fv);
    return divide_out(dgf, in.z*out.z);
  }
  // Value and parameter jacobian of the gray unscaled lobe u = D G F / (N z_i z_o) for the batched loss kernels.  D and G
  // depend on the NDF's parameters only, F on its own: their tangents never mix, so instead of pushing NA_ndf + NA_f
  // tangents through every product (eval_unscaled_g<Dual<N>>), D G carries NA_ndf tangents, F carries NA_f, and
  //   du/d(ndf) = (D G)' F k,   du/d(f) = D G F' k.
  static constexpr int kNdfParams = NDF::NA, kFresnelParams = F::NA;
  template<int N> BBMCU_D static Dual<N> unscaled_gray_jacobian(const Geom& g, f3 in, f3 out, const float* a, int component)
  {
    static_assert(kGrayUnscaled && N == NDF::NA + F::NA, "scalar D, G, F with every NDF and Fresnel parameter fitted");
    Dual<N> r(0.0f);
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return r;
    using DN = Dual<NDF::NA>;
    using DF = Dual<F::NA>;
    DN an[NDF::NA];
    DF af[F::NA];
#pragma unroll
    for(int i=0; i < NDF::NA; ++i) { an[i] = DN(a[OFF_NDF + i]); an[i].d[i] = 1.0f; }
#pragma unroll
    for(int i=0; i < F::NA; ++i) { af[i] = DF(a[OFF_F + i]); af[i].d[i] = 1.0f; }
    const DN dg = NDF::template D<DN>(g.h, an) * G::template eval<NDF, DN>(in, out, g.h, an);
    const DF fv = F::template evalq<DF>(af, 0.5f*(g.inh + g.outh));
    const float k = q_rcp((float)norm() * (in.z*out.z));
    const float dgk = dg.v * k, fk = fv.v * k;
    r.v = dgk * fv.v;
#pragma unroll
    for(int i=0; i < NDF::NA; ++i) r.d[i] = dg.d[i] * fk;
#pragma unroll
    for(int i=0; i < F::NA; ++i) r.d[NDF::NA + i] = dgk * fv.d[i];
    return r;
  }
  // eval without the leading scale (microfacet.h:74-102)
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    f3 h = std::is_same<T, float>::value ? quick_halfway<NDF>(in, out) : halfway(in, out);
    float inh = dot(in, h), outh = dot(out, h);
    auto D = NDF::template D<T>(h, a + OFF_NDF);
    auto Gv = G::template eval<NDF, T>(in, out, h, a + OFF_NDF);
    auto Fv = F::template evalq<T>(a + OFF_F, 0.5f*(inh + outh));
    // D*G*F / NormalizationFactor / (z_in z_out): the division chain runs in double in the reference
    Spec<T> dgf = to_spec(D) * to_spec(Gv) * to_spec(Fv);
    return divide_out(dgf, in.z*out.z);
  }
  BBMCU_D static Spec<float> divide_out(const Spec<float>& s, float zz)
  {
    // a double division chain in the reference (microfacet.h:100); a final value here: one quick reciprocal
    float k = q_rcp((float)norm() * zz);
    return Spec<float>(s.r*k, s.g*k, s.b*k);
  }
  template<int N> BBMCU_D static Spec<Dual<N>> divide_out(const Spec<Dual<N>>& s, float zz)
  {
    Spec<float> v = divide_out(Spec<float>(s.r.v, s.g.v, s.b.v), zz);
    float k = 1.0f / ((float)norm() * zz);
    Spec<Dual<N>> r = s * k;
    r.r.v = v.r; r.g.v = v.g; r.b.v = v.b;
    return r;
  }
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component)
  {
    Spec<T> r = eval_unscaled<T>(in, out, a, component);
    if(SCALED) r = r * load_spec(a);
    return r;
  }
  // eval with the NDF's parameter-only factors formed by the caller (NDF::pre): the same operations as eval<float>
  static constexpr bool kHasPre = NdfHasPre<NDF>::value && std::is_same<G, GUncorrelated>::value;
  BBMCU_D static auto precompute(const float* a) { if constexpr (kHasPre) return NDF::pre(a + OFF_NDF); else return 0; }
  template<class PRE> BBMCU_D static Spec<float> eval_pre(f3 in, f3 out, const float* a, int component, const PRE& q)
  {
    if constexpr (!kHasPre) return eval<float>(in, out, a, component);
    else
    {
      if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<float>(0.0f);
      f3 h = quick_halfway<NDF>(in, out);
      float inh = dot(in, h), outh = dot(out, h);
      auto D = NDF::template D<float>(h, a + OFF_NDF);
      float Gv = g_mask(in, out, h) ? NDF::G1_pre(in, h, a + OFF_NDF, q) * NDF::G1_pre(out, h, a + OFF_NDF, q) : 0.0f;   // GUncorrelated
      auto Fv = F::template evalq<float>(a + OFF_F, 0.5f*(inh + outh));
      Spec<float> r = divide_out(to_spec(D) * to_spec(Gv) * to_spec(Fv), in.z*out.z);
      if(SCALED) r = r * load_spec(a);
      return r;
    }
  }
  // pdf(h) / (4 |out.h|)  (microfacet.h:154-174)
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((out.z > 0.0f) && (in.z > 0.0f))) return 0.0f;
    f3 h = quick_halfway<NDF>(in, out);
    if(h.z < 0.0f) h = -h;
    return q_div(NDF::pdf(out, h, a + OFF_NDF), 4.0f * fabsf(dot(out, h)));
  }
  // Walter's GGX has a hand-merged eval + pdf below
  static constexpr bool kHandFusedEvalPdf = std::is_same<NDF, NdfGGX<false>>::value && std::is_same<G, GUncorrelated>::value && std::is_same<F, FresnelCookIor>::value && NORM == 1;
  // eval and pdf of ONE direction pair sharing the half vector, D and G1 (the fused sample -> eval -> pdf pass).
  // in.z > 0 and out.z > 0 imply h.z > 0, so pdf's flip of h (microfacet.h:163) never triggers here.
  BBMCU_D static void eval_pdf(f3 in, f3 out, const float* a, int component, Spec<float>& e, float& p)
  {
    e = Spec<float>(0.0f); p = 0.0f;
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return;
    if constexpr (kHandFusedEvalPdf)
    {
      // Walter's GGX (isotropic, uncorrelated Smith G, Cook Fresnel, N = 4) - the BASELINE headline model - with the
      // quotients of D, G1, F and the 1/(4 z_i z_o) chain merged algebraically: 9 MUFU operations for eval AND pdf.
      //   D   = alpha^2 / (pi (hx^2 + hy^2 + alpha^2 hz^2)^2)
      //   G1  = 2 z / (z + sqrt(z^2 + alpha^2 (1 - z^2)))
      //   F   = 1/2 a^2 (1 + b^2),  a = A d2 R,  b = n2 B R,  A = g - c, B = g + c, n2 = c B - 1, d2 = c A + 1, R = 1/(B d2)
      //   pdf = D G1(out) |o.h| / z_o / (4 |o.h|) = D G1(out) z_i / (4 z_i z_o)
      // Same mathematics as the generic path below, different (and fewer) roundings: final values, ~3e-7 relative.
      const float al = a[OFF_NDF], al2 = al*al, eta = a[OFF_F];
      f3 h = q_normalize(in + out);
      const float inh = q_dot(in, h), outh = q_dot(out, h);
      const float den = fmaf(al2, h.z*h.z, fmaf(h.x, h.x, h.y*h.y));
      const float Dv = sel_gt(h.z, 0.0f, al2 * q_rcp(kPi * den * den));
      const float gi = sel_gt(inh, 0.0f, 2.0f*in.z * q_rcp(in.z + q_sqrt(fmaf(al2, sinTheta2(in), in.z*in.z))));
      const float go = sel_gt(outh, 0.0f, 2.0f*out.z * q_rcp(out.z + q_sqrt(fmaf(al2, sinTheta2(out), out.z*out.z))));
      const float c = 0.5f*(inh + outh);
      const float g = q_sqrt(max0(eta*eta + c*c - 1.0f));
      const float A = g - c, B = g + c, n2 = fmaf(c, B, -1.0f), d2 = fmaf(c, A, 1.0f);
      const float R = q_rcp(B * d2);
      const float fa = A*d2*R, fb = n2*B*R;
      const float Fv = fmaxf(0.5f * (fa*fa) * fmaf(fb, fb, 1.0f), 0.0f);
      const float R4 = q_rcp(4.0f * in.z * out.z);
      const float u = Dv * (gi*go) * Fv * R4;
      e = Spec<float>(u*a[0], u*a[1], u*a[2]);
      const float pv = Dv * go * in.z * R4;
      p = (pv > 0.0f) ? pv : 0.0f;
      return;
    }
    if constexpr (std::is_same<G, GUncorrelated>::value && HasPdfVisible<NDF>::value && QuickHalfway<NDF>::value)
    {
      // scalar D and G1, uncorrelated G, visible-normal pdf (GGX): D, G1(out) and both dots are evaluated once
      f3 h = q_normalize(in + out);
      float inh = q_dot(in, h), outh = q_dot(out, h);
      auto D = NDF::template D<float>(h, a + OFF_NDF);
      auto gi = NDF::template G1<float>(in, h, a + OFF_NDF), go = NDF::template G1<float>(out, h, a + OFF_NDF);
      auto Gv = ((inh > 0.0f) && (outh > 0.0f)) ? gi*go : decltype(gi)(0.0f);
      auto Fv = F::template evalq<float>(a + OFF_F, 0.5f*(inh + outh));
      e = divide_out(to_spec(D) * to_spec(Gv) * to_spec(Fv), in.z*out.z);
      if(SCALED) e = e * load_spec(a);
      p = q_div(NDF::pdf_visible(D, go, fabsf(outh), out.z), 4.0f * fabsf(outh));
      return;
    }
    f3 h = quick_halfway<NDF>(in, out);
    float inh = dot(in, h), outh = dot(out, h);
    auto D = NDF::template D<float>(h, a + OFF_NDF);
    auto Gv = G::template eval<NDF, float>(in, out, h, a + OFF_NDF);
    auto Fv = F::template evalq<float>(a + OFF_F, 0.5f*(inh + outh));
    e = divide_out(to_spec(D) * to_spec(Gv) * to_spec(Fv), in.z*out.z);
    if(SCALED) e = e * load_spec(a);
    p = q_div(NDF::pdf(out, h, a + OFF_NDF), 4.0f * fabsf(outh));
  }
  // eval_pdf with the NDF's parameter-only factors formed by the caller: the general branch above, G1 through G1_pre
  template<class PRE> BBMCU_D static void eval_pdf_pre(f3 in, f3 out, const float* a, int component, Spec<float>& e, float& p, const PRE& q)
  {
    if constexpr (!kHasPre) eval_pdf(in, out, a, component, e, p);
    else
    {
      static_assert(!kHandFusedEvalPdf && !(HasPdfVisible<NDF>::value && QuickHalfway<NDF>::value), "eval_pdf takes its general branch for this model");
      e = Spec<float>(0.0f); p = 0.0f;
      if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return;
      f3 h = quick_halfway<NDF>(in, out);
      float inh = dot(in, h), outh = dot(out, h);
      auto D = NDF::template D<float>(h, a + OFF_NDF);
      float Gv = g_mask(in, out, h) ? NDF::G1_pre(in, h, a + OFF_NDF, q) * NDF::G1_pre(out, h, a + OFF_NDF, q) : 0.0f;   // GUncorrelated
      auto Fv = F::template evalq<float>(a + OFF_F, 0.5f*(inh + outh));
      e = divide_out(to_spec(D) * to_spec(Gv) * to_spec(Fv), in.z*out.z);
      if(SCALED) e = e * load_spec(a);
      p = q_div(NDF::pdf(out, h, a + OFF_NDF), 4.0f * fabsf(outh));
    }
  }
  // The whole fused element - sample, eval, pdf - of the hand-merged model without a branch: every validity test of
  // sample (microfacet.h:118-127) and eval / pdf (:74-81, :154-160) becomes a predicate that selects zeros at the end; the
  // arithmetic runs on whatever the element holds (the *_raw operations neither trap nor loop on garbage).
  BBMCU_D static void sample_eval_pdf_merged(f3 out, f2 xi, const float* a, int component, f3& dir, int& flag, Spec<float>& e, float& p)
  {
    float u;
    sample_eval_pdf_merged_u(out, xi, a, component, dir, flag, u, p);
    e = Spec<float>(u*a[0], u*a[1], u*a[2]);
  }
  // the same with the value before the leading RGB scale: eval = (u a[0], u a[1], u a[2]).  The host-pointer path sends u (4
  // bytes per element instead of 12) and lets host threads form the three products - IEEE single multiplications on both
  // sides, the same bits (bbmcu_api.cu, XFER_GRAY_SCALED)
  BBMCU_D static void sample_eval_pdf_merged_u(f3 out, f2 xi, const float* a, int component, f3& dir, int& flag, float& u, float& p)
  {
    static_assert(kHandFusedEvalPdf, "only the hand-merged model has this path");
    const bool ok_s = (component & FLAG_SPECULAR) && xi_valid(xi) && (out.z > 0.0f);
    const f3 m = NDF::sample_unchecked(out, xi, a + OFF_NDF);
    const float d2m = 2.0f * q_dot(m, out);                       // reflect(out, m) = m (2 m.out) - out, a final value: fused
    const f3 in = make_f3(fmaf(m.x, d2m, -out.x), fmaf(m.y, d2m, -out.y), fmaf(m.z, d2m, -out.z));
    const bool ok_e = ok_s && (in.z > 0.0f);
    const float al = a[OFF_NDF], al2 = al*al, eta = a[OFF_F];
    const f3 h = q_normalize(in + out);
    const float inh = q_dot(in, h), outh = q_dot(out, h);
    const float den = fmaf(al2, h.z*h.z, fmaf(h.x, h.x, h.y*h.y));
    const float Dv = sel_gt(h.z, 0.0f, al2 * q_rcp(kPi * den * den));
    const float gi = sel_gt(inh, 0.0f, 2.0f*in.z * q_rcp(in.z + q_sqrt(fmaf(al2, sinTheta2(in), in.z*in.z))));
    const float go = sel_gt(outh, 0.0f, 2.0f*out.z * q_rcp(out.z + q_sqrt(fmaf(al2, sinTheta2(out), out.z*out.z))));
    const float c = 0.5f*(inh + outh);
    const float g = q_sqrt(max0(eta*eta + c*c - 1.0f));
    const float A = g - c, B = g + c, n2 = fmaf(c, B, -1.0f), d2 = fmaf(c, A, 1.0f);
    const float R = q_rcp(B * d2);
    const float fa = A*d2*R, fb = n2*B*R;
    const float Fv = fmaxf(0.5f * (fa*fa) * fmaf(fb, fb, 1.0f), 0.0f);
    const float R4 = q_rcp(4.0f * in.z * out.z);
    u = ok_e ? Dv * (gi*go) * Fv * R4 : 0.0f;
    const float pv = Dv * go * in.z * R4;
    dir = make_f3(ok_s ? in.x : 0.0f, ok_s ? in.y : 0.0f, ok_s ? in.z : 0.0f);
    flag = ok_s ? FLAG_SPECULAR : FLAG_NONE;
    p = (ok_e && (pv > 0.0f)) ? pv : 0.0f;
  }
  // sample.pdf is pdf(sample.direction, out) (microfacet.h:138): fused sample -> eval -> pdf passes evaluate it once
  static constexpr bool kSamplePdfIsPdf = true;
  BBMCU_D static void sample_dir(f3 out, f2 xi, const float* a, int component, f3& dir, int& flag)
  {
    dir = make_f3(0, 0, 0); flag = FLAG_NONE;
    if(!(component & FLAG_SPECULAR) || !xi_valid(xi) || !(out.z > 0.0f)) return;
    f3 m;
    if constexpr (HasSampleUnchecked<NDF>::value) m = NDF::sample_unchecked(out, xi, a + OFF_NDF);
    else m = NDF::sample(out, xi, a + OFF_NDF);
    dir = reflect(out, m);
    flag = FLAG_SPECULAR;
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    sample_dir(out, xi, a, component, dir, flag);
    pdfv = (flag != FLAG_NONE) ? pdf(dir, out, a, component) : 0.0f;
  }
  // perfect-mirror approximation F(eta, out.z) / N * 4.0 (microfacet.h:186-199), times scale
  BBMCU_D static Spec<float> reflectance(f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !(out.z > 0.0f)) return Spec<float>(0.0f);
    Spec<float> f = to_spec(F::template eval<float>(a + OFF_F, out.z));
    double n = norm();
    Spec<float> r((float)((double)f.r / n * 4.0), (float)((double)f.g / n * 4.0), (float)((double)f.b / n * 4.0));
    if(SCALED) r = r * load_spec(a);
    return r;
  }
};

} // namespace bbmcu
