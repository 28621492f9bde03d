// The He-Torrance-Sillion-Greenberg directional-specular model in its four bbm flavours (He, HeWestin,
// HeHolzschuch, NganHe) and the data-driven importance sampler the reference wraps them in.
//
// Behaviour follows (restated, not copied):
//   include/bsdfmodel/he.h:141-170 (eval), :227-237 (reflectance), :266-291 (S, Eq. 24-25),
//   :306-352 (G, Eq. 76), :365-400 (apparent roughness, Eq. 80, 4 Newton steps),
//   :411-467 (D, Eq. 78-79: adaptive Taylor series blended with Beckmann's rough approximation), :489-496 (variants)
//   include/bsdfmodel/ngan.h:166-167 (NganHe), include/bsdfmodel/scaledmodel.h:50-67
//   include/bbm/ndf_sampler.h:79-156 (sample / pdf through a back-scatter "NDF" hsum(eval(h, h)))
//   include/ndf/sampler.h:63-92 (sample), :102-128 (pdf), :143-181 (90-bin CDF), include/util/cdf.h
//
// The reference builds the CDF lazily on the host whenever the parameters change; here every thread block
// rebuilds it in shared memory in its prologue (90 evaluations, one per thread; see bbmcu_kernels.cuh), so
// there is no host mathematics, no cache to invalidate and nothing to upload.  The device attribute block of a He
// lobe is [NA attributes][NT = 90 CDF entries].
//
// eval is a template over T: T = float reproduces the native backbone's float/double mix (the Taylor
// series runs in double, he.h:418-455); T = Dual<N> carries parameter derivatives in float.
#pragma once
#include "bbmcu_ndfsampler.cuh"

namespace bbmcu {

enum : int { HE_VARIANT_HE = 0, HE_VARIANT_WESTIN, HE_VARIANT_HOLZSCHUCH, HE_VARIANT_NGAN };

// the working type of the spots the reference evaluates in double
template<class T> struct WideOf { using type = T; };
template<> struct WideOf<float> { using type = double; };
BBMCU_D double he_wide(float a) { return (double)a; }
template<int N> BBMCU_D Dual<N> he_wide(const Dual<N>& a) { return a; }
BBMCU_D float he_narrow(double a) { return (float)a; }
template<int N> BBMCU_D Dual<N> he_narrow(const Dual<N>& a) { return a; }
BBMCU_D double m_exp(double a) { return exp(a); }
BBMCU_D double m_log(double a) { return log(a); }
BBMCU_D double m_sqrt(double a) { return sqrt(a); }
BBMCU_D double m_safe_sqrt(double a) { return safe_sqrt_d(a); }
BBMCU_D double val(double a) { return a; }

// exp of a double argument to float accuracy (the terms of Eq. 78 are rounded to float anyway): expf of the float part
// times (1 + remainder); ~2 ulp of a float, i.e. 1e-7 relative against the reference's double exp - the quick tier.
// On the host this is the plain double exp.
BBMCU_D double he_exp(double x)
{
#ifdef __CUDA_ARCH__
  const float hi = (float)x;
  const float lo = (float)(x - (double)hi);
  return (double)(expf(hi) * (1.0f + lo));
#else
  return exp(x);
#endif
}

// erfc and the double exp of the series, by tier.  STRICT is the path of the sampling CDF (backscatter below): the
// reference's CDF entries must come out bit for bit, or a xi within an ulp of an entry picks the neighbouring bin and the
// sampled direction jumps by a bin width (5e-6 of all samples before).  There erfc is the host libm's erfcf restated
// (bbmcu_libm.cuh), exp the true double exp, and the series keeps the reference's divisions.
template<bool STRICT> BBMCU_D float he_erfc(float a) { return STRICT ? glibc_erfcf(a) : erfcf(a); }
template<bool STRICT, int N> BBMCU_D Dual<N> he_erfc(const Dual<N>& a) { return m_erfc(a); }
template<bool STRICT> BBMCU_D double he_exp_t(double x) { return STRICT ? exp(x) : he_exp(x); }

template<int V> struct HeTraits;
template<> struct HeTraits<HE_VARIANT_HE>         { static constexpr bool ERRATA = false, WESTIN = false, ADAPTIVE = true,  ROUGH = true,  SCALED = false; static constexpr int TERMS = 64; using F = FresnelComplexRGB; };
template<> struct HeTraits<HE_VARIANT_WESTIN>     { static constexpr bool ERRATA = true,  WESTIN = true,  ADAPTIVE = true,  ROUGH = true,  SCALED = false; static constexpr int TERMS = 64; using F = FresnelComplexRGB; };
template<> struct HeTraits<HE_VARIANT_HOLZSCHUCH> { static constexpr bool ERRATA = true,  WESTIN = false, ADAPTIVE = false, ROUGH = false, SCALED = false; static constexpr int TERMS = 10; using F = FresnelComplexRGB; };
template<> struct HeTraits<HE_VARIANT_NGAN>       { static constexpr bool ERRATA = true,  WESTIN = true,  ADAPTIVE = true,  ROUGH = true,  SCALED = true;  static constexpr int TERMS = 64; using F = FresnelCookIor; };

template<int V>
struct HeModel : NdfSamplerCdf<HeModel<V>>
{
  using Tr = HeTraits<V>;
  using F = typename Tr::F;
  static constexpr int kLaunchMinBlocks = (V == HE_VARIANT_HOLZSCHUCH) ? 3 : 1;      // see LaunchMinBlocks (bbmcu_bsdf.cuh)
  static constexpr int SCALE = Tr::SCALED ? 0 : -1;
  static constexpr int OFF_R = Tr::SCALED ? 3 : 0;       // roughness (sigma0), then autocorrelation (tau)
  static constexpr int OFF_F = OFF_R + 2;
  static constexpr int NA = OFF_F + F::NA;
  static constexpr int NT = kHeCdfBins;                  // device-side CDF appended to the attribute block
  // value-only batched losses take the all-float templates (bbmcu_lossop.cuh) - except Holzschuch's variant, whose ten-term series
  // with the split exponent (one expf per term for the three channels) is already cheaper than three float exps per term
  // (configs[4] sweep, seconds per rank: HeWestin 14.6 -> 7.0, NganHe 14.0 -> 4.8, He 8.1 -> 6.5, HeHolzschuch 7.4 -> 9.1)
  static constexpr bool kQuickLossValue = (V != HE_VARIANT_HOLZSCHUCH);

  // ---- Eq. 24-25: mono-directional shadowing ---------------------------------------------------------
  template<class T, bool STRICT = false> BBMCU_D static T S1(f3 v, const T& rough, const T& tau)
  {
    using W = typename WideOf<T>::type;
    if(val(rough) < kEps) return T(1.0f);
    float cot = 1.0f / tanTheta(v);
    T scaledCot = he_narrow(he_wide(tau * cot) / (he_wide(rough) * 2.0f));
    T erfcv = he_narrow(he_wide(he_erfc<STRICT>(scaledCot)) * 0.5f);
    T Lambda = he_narrow((W)(0.5f * he_wide(T(kInvSqrtPi))) / he_wide(scaledCot));
    if(Tr::ERRATA) { W s = he_wide(scaledCot); Lambda = he_narrow(he_wide(Lambda) * m_exp(-(s*s))); }
    Lambda = Lambda - erfcv;
    return he_narrow((1.0f - he_wide(erfcv)) / (he_wide(Lambda) + 1.0f));
  }

  // ---- Eq. 76: geometrical factor (directions only) ---------------------------------------------------
  BBMCU_D static float G(f3 in, f3 out)
  {
    f3 v = in + out;
    double vs = (double)(dot(v, v) / v.z);
    float v_scale = (float)(vs*vs);
    float kixn2 = 1.0f - in.z*in.z, krxn2 = 1.0f - out.z*out.z;
    float kikr = dot(-in, out);
    float sikr = out.y*in.x - out.x*in.y, srki = in.y*out.x - in.x*out.y;
    float pikr = out.z + kikr*in.z, prki = in.z + kikr*out.z;
    double dd = 1.0 - (double)(kikr*kikr);
    float denom = (float)(dd*dd);
    float nom = (float)(((double)sikr*(double)sikr + (double)pikr*(double)pikr) * ((double)srki*(double)srki + (double)prki*(double)prki) / (double)(krxn2*kixn2));
    return (denom > kEps) ? v_scale * nom / denom : 1.0f;
  }

  // ---- Eq. 80: apparent roughness, Newton-Raphson ------------------------------------------------------
  template<class T, bool STRICT = false> BBMCU_D static T sigma(f3 in, f3 out, const T& rough, const T& tau)
  {
    using W = typename WideOf<T>::type;
    if(!(val(rough) > kEps)) return T(0.0f);
    float ti = tanTheta(in), to = tanTheta(out);
    T Ki = (ti > kEps) ? ti * he_erfc<STRICT>(tau / (2.0f * rough * ti)) : T(0.0f);
    T Ko = (to > kEps) ? to * he_erfc<STRICT>(tau / (2.0f * rough * to)) : T(0.0f);
    const float rs8pi = 1.0f / sqrtf((float)(8.0 * kPiD));
    T f0 = rs8pi * (Ki + Ko);
    T x = (val(f0) <= 1.0f) ? f0 : he_narrow(m_safe_sqrt(2.0f * he_wide(m_log(f0))));
#pragma unroll
    for(int step=0; step < 4; ++step)
    {
      W xw = he_wide(x);
      T expn = he_narrow(m_exp((0.5f * xw) * xw));
      T ev = x*expn - f0;
      T grad = (1.0f + x*x) * expn;
      if(val(grad) > kEps) x = x - ev / grad;
    }
    return rough / m_safe_sqrt(1.0f + x*x);
  }

  // ---- Eq. 78-79: distribution ------------------------------------------------------------------------
  BBMCU_D static float lerpf(float a, float b, float t)            // std::lerp<float>
  {
    if((a <= 0.0f && b >= 0.0f) || (a >= 0.0f && b <= 0.0f)) return t*b + (1.0f - t)*a;
    if(t == 1.0f) return b;
    float x = a + t*(b - a);
    return ((t > 1.0f) == (b > a)) ? (b < x ? x : b) : (b > x ? x : b);
  }
  template<bool STRICT = false> BBMCU_D static Spec<float> D(f3 in, f3 out, const float& rough, const float& tau)
  {
    const float wl[3] = {0.645f, 0.526f, 0.444f};            // floatRGB wavelengths (backbone/native/include/backbone.h:36)
    const float thr = Tr::ROUGH ? 18.0f : 3.402823466e+38f;
    float sx = in.x + out.x, sy = in.y + out.y;
    float v_xy2 = sx*sx + sy*sy;
    float base = kTwoPi * sigma<float, STRICT>(in, out, rough, tau) * (in.z + out.z);
    float tau2 = (float)((double)tau * (double)tau);
    const float pi2q = (0.25f * kPi) * kPi, pi2x4 = (4.0f * kPi) * kPi;
    double g[3], nrm[3]; float eb[3];
#pragma unroll
    for(int c=0; c < 3; ++c)
    {
      double q = (double)(base / wl[c]);  g[c] = q*q;
      double w2 = (double)wl[c] * (double)wl[c];
      nrm[c] = (double)(pi2q * tau2) / w2;
      eb[c] = v_xy2 * tau2 / 4.0f;
      if(Tr::WESTIN) eb[c] = (float)((double)eb[c] * ((double)pi2x4 / w2));
    }
    double gmin = fmin(fmin(g[0], g[1]), g[2]);
    if(!(g[0] == g[0]) || !(g[1] == g[1]) || !(g[2] == g[2])) gmin = g[0] + g[1] + g[2];   // std::min_element keeps the first on NaN compare; a NaN g poisons everything anyway
    float ra[3] = {0.0f, 0.0f, 0.0f}, weight = 0.0f;
    if(gmin > (double)thr)
    {
#pragma unroll
      for(int c=0; c < 3; ++c) ra[c] = (float)(exp((double)(-eb[c]) / g[c]) / g[c]);
      double wv = gmin - (double)thr;
      weight = (float)(wv < 0.0 ? 0.0 : (1.0 < wv ? 1.0 : wv));
    }
    float sum[3] = {0.0f, 0.0f, 0.0f}, gm[3] = {1.0f, 1.0f, 1.0f}, term[3] = {0.0f, 0.0f, 0.0f}, last_min = -1.0f;
    bool converged = (gmin - 1.0) > (double)thr;
    // quick tier, variants whose eb does not depend on the wavelength (He, HeHolzschuch): exp(-g - eb/m) = exp(-g) exp(-eb/m).
    // The first factor does not depend on the term (one double exp per channel, outside the loop), the second has a float
    // argument and is ONE expf per term for all three channels.  The series is what puts this model at the XU roof (exp +
    // float <-> double conversions per term and channel: profiles/r01_s14_pipe_utilisation_all_models.json): He eval 2.9 ->
    // 4.9 G/s, HeHolzschuch 7.4 -> 9.8.  The product differs from the reference's single exp by ~2e-7 relative, as he_exp
    // does.  Westin's variants (eb scaled per wavelength, exponents up to ~10^2) keep he_exp: with the split (+24 % / +43 %)
    // one evaluation in 2^21 ended its adaptive series one term apart from the reference (1.2e-5 relative).
    constexpr bool SPLIT = !STRICT && !Tr::WESTIN;
    double Eg[3] = {0.0, 0.0, 0.0};
    if(SPLIT && !converged) {
#pragma unroll
      for(int c=0; c < 3; ++c) Eg[c] = exp(-g[c]);
    }
    for(int m=1; m <= Tr::TERMS && !converged; ++m)
    {
      float tmin = 3.402823466e+38f;
      // one double reciprocal per term instead of two double divisions per channel: g/m and (.)/m become products with
      // 1/m, equal to the reference's quotients to 1 ulp of a double - invisible after the rounding to float
      const double inv_m = 1.0 / (double)m;
      float em_all = 0.0f;
      if(SPLIT) em_all = expf(-(eb[0] / (float)m));                              // the reference's float quotient (he.h:451)
#pragma unroll
      for(int c=0; c < 3; ++c)
      {
        if(STRICT)
        {
          gm[c] = (float)((double)gm[c] * (g[c] / (double)m));                                    // gm *= g / m   (he.h:450)
          term[c] = (float)(exp(-g[c] - (double)(eb[c] / (float)m)) * (double)gm[c] / (double)m);  // he.h:451
        }
        else
        {
          gm[c] = (float)((double)gm[c] * (g[c] * inv_m));
          if(SPLIT) term[c] = (float)((Eg[c] * (double)em_all) * ((double)gm[c] * inv_m));
          else term[c] = (float)(he_exp(-g[c] - (double)(eb[c] / (float)m)) * (double)gm[c] * inv_m);
        }
        sum[c] += term[c];
      }
      tmin = term[0]; if(term[1] < tmin) tmin = term[1]; if(term[2] < tmin) tmin = term[2];      // std::min_element order
      if(Tr::ADAPTIVE) converged = (tmin < kEps) && (tmin < last_min);
      last_min = tmin;
    }
    return Spec<float>((float)(nrm[0] * (double)lerpf(sum[0], ra[0], weight)), (float)(nrm[1] * (double)lerpf(sum[1], ra[1], weight)),
                       (float)(nrm[2] * (double)lerpf(sum[2], ra[2], weight)));
  }
  // derivative-carrying version (float arithmetic; same control flow, decided on the values)
  template<bool STRICT = false, int N> BBMCU_D static Spec<Dual<N>> D(f3 in, f3 out, const Dual<N>& rough, const Dual<N>& tau)
  {
    using T = Dual<N>;
    const float wl[3] = {0.645f, 0.526f, 0.444f};
    const float thr = Tr::ROUGH ? 18.0f : 3.402823466e+38f;
    float sx = in.x + out.x, sy = in.y + out.y;
    float v_xy2 = sx*sx + sy*sy;
    T base = kTwoPi * sigma<T>(in, out, rough, tau) * (in.z + out.z);
    T tau2 = tau * tau;
    const float pi2q = (0.25f * kPi) * kPi, pi2x4 = (4.0f * kPi) * kPi;
    T g[3], nrm[3], eb[3];
#pragma unroll
    for(int c=0; c < 3; ++c)
    {
      T q = base / wl[c];  g[c] = q*q;
      float w2 = wl[c]*wl[c];
      nrm[c] = (pi2q * tau2) / w2;
      eb[c] = v_xy2 * tau2 / 4.0f;
      if(Tr::WESTIN) eb[c] = eb[c] * (pi2x4 / w2);
    }
    float gmin = fminf(fminf(g[0].v, g[1].v), g[2].v);
    int cmin = (g[0].v <= g[1].v && g[0].v <= g[2].v) ? 0 : (g[1].v <= g[2].v ? 1 : 2);
    T ra[3] = {T(0.0f), T(0.0f), T(0.0f)}, weight(0.0f);
    if(gmin > thr)
    {
#pragma unroll
      for(int c=0; c < 3; ++c) ra[c] = m_exp(-eb[c] / g[c]) / g[c];
      float wv = gmin - thr;
      if(wv >= 1.0f) weight = T(1.0f); else weight = g[cmin] - thr;
    }
    T sum[3] = {T(0.0f), T(0.0f), T(0.0f)}, gm[3] = {T(1.0f), T(1.0f), T(1.0f)}, term[3] = {T(0.0f), T(0.0f), T(0.0f)};
    float last_min = -1.0f;
    bool converged = (gmin - 1.0f) > thr;
    for(int m=1; m <= Tr::TERMS && !converged; ++m)
    {
      // (the three quotients by m per channel share ONE quick reciprocal: the compiler merges the dual_rcp((float)m) calls.  A
      // correctly rounded reciprocal hoisted by hand - __frcp_rn - was 5 - 8 % slower on the He family's share of the configs[4]
      // sweep, profiles/r02_s46 against r02_s48)
#pragma unroll
      for(int c=0; c < 3; ++c)
      {
        gm[c] = gm[c] * (g[c] / (float)m);
        term[c] = m_exp(-g[c] - eb[c] / (float)m) * gm[c] / (float)m;
        sum[c] = sum[c] + term[c];
      }
      float tmin = fminf(fminf(term[0].v, term[1].v), term[2].v);
      if(Tr::ADAPTIVE) converged = (tmin < kEps) && (tmin < last_min);
      last_min = tmin;
    }
    Spec<T> r;
    r.r = nrm[0] * (sum[0] + weight*(ra[0] - sum[0]));
    r.g = nrm[1] * (sum[1] + weight*(ra[1] - sum[1]));
    r.b = nrm[2] * (sum[2] + weight*(ra[2] - sum[2]));
    return r;
  }

  // ---- the BSDF concept ----------------------------------------------------------------------------------
  template<class T, bool STRICT = false> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    const T& rough = a[OFF_R]; const T& tau = a[OFF_R + 1];
    T S = S1<T, STRICT>(in, rough, tau) * S1<T, STRICT>(out, rough, tau);
    float Gt = G(in, out);
    Spec<T> Dt = D<STRICT>(in, out, rough, tau);
    float cosHalf = (float)safe_sqrt_d((double)(1.0f + dot(in, out)) / 2.0);
    Spec<T> Ft = to_spec(F::template eval<T>(a + OFF_F, cosHalf));
    float nrm = 1.0f / (kPi * in.z * out.z);
    return (((Ft * nrm) * S) * Gt) * Dt;
  }
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component)
  {
    Spec<T> r = eval_unscaled<T>(in, out, a, component);
    if(Tr::SCALED) r = r * load_spec(a);
    return r;
  }
  BBMCU_D static Spec<float> reflectance(f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !(out.z > 0.0f)) return Spec<float>(0.0f);
    Spec<float> f = to_spec(F::template eval<float>(a + OFF_F, out.z));
    Spec<float> r((float)((double)(f.r / kPi) * 4.0), (float)((double)(f.g / kPi) * 4.0), (float)((double)(f.b / kPi) * 4.0));
    if(Tr::SCALED) r = r * load_spec(a);
    return r;
  }

  // ---- data-driven sampling: ndf::sampler over the back-scatter of the UNSCALED model (bbmcu_ndfsampler.cuh) ----------
  BBMCU_D static float backscatter(const float* a, int component, f3 h) { return hsum(eval_unscaled<float, true>(h, h, a, component)); }
};

} // namespace bbmcu
