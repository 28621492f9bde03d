// Closed-form BSDF lobes: Lambertian, Oren-Nayar, Phong, Lafortune (+Ngan normalisation),
// Ward / Ward-Duer / Ward-Duer-Geisler-Moroder, Ashikhmin-Shirley (3 Fresnel variants + the full
// model with coupled diffuse term) and Low's smooth-surface model.
//
// Behaviour follows (restated, not copied), with the reference's strict / non-strict horizon
// tests kept exactly (SURVEY.md section 8a "Horizon / component tests"):
//   include/bsdfmodel/lambertian.h:45-147, orennayar.h:43-120, phong.h:43-145,
//   include/bsdfmodel/lafortune.h:46-165, ngan.h:54-129,
//   include/bsdfmodel/ward.h:44-150, wardduer.h:55-77, wardduergeislermoroder.h:54-77,
//   include/bsdfmodel/ashikhminshirley.h:51-215, ashikhminshirleyfull.h:55-185,
//   include/bsdfmodel/lowsmooth.h:35-185
//
// Model interface (all static, attribute block `a` in reflection order, see bbmcu_models.cuh):
//   NA, SCALE (offset of a leading RGB scale the model multiplies its result with, or -1)
//   eval_unscaled<T>(in, out, a, component) -> Spec<T>     (T = float or Dual<N>)
//   eval<T>(...) = eval_unscaled * scale
//   sample(out, xi, a, component, dir&, pdf&, flag&), pdf(in, out, a, component), reflectance(out, a, component)
#pragma once
#include "bbmcu_microfacet.cuh"

namespace bbmcu {

#define BBMCU_SCALED_EVAL                                                                       \
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component)       \
  { Spec<T> r = eval_unscaled<T>(in, out, a, component); if(SCALE >= 0) r = r * load_spec(a + (SCALE >= 0 ? SCALE : 0)); return r; }

// cosine-weighted hemisphere sampling shared by Lambertian / OrenNayar / ASFull / He placeholders
// (lambertian.h:76-103,115-125)
BBMCU_D float lambert_pdf(f3 in, f3 out, int component)
{
  bool m = (component & FLAG_DIFFUSE) && (in.z >= 0.0f) && (out.z >= 0.0f);
  return m ? in.z * kInvPi : 0.0f;
}
BBMCU_D void lambert_sample(f3 out, f2 xi, int component, f3& dir, float& pdfv, int& flag)
{
  dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
  if(!(component & FLAG_DIFFUSE) || !xi_valid(xi)) return;
  float ph = xi.x * kTwoPi;
  float sinT = (float)safe_sqrt_d(1.0 - (double)xi.y);
  dir = make_f3(cosf(ph)*sinT, sinf(ph)*sinT, m_safe_sqrt(xi.y));
  pdfv = lambert_pdf(dir, out, component);
  flag = FLAG_DIFFUSE;
}

// ---- Lambertian: albedo / pi --------------------------------------------------------------------
struct Lambertian
{
  static constexpr int NA = 3, SCALE = 0;
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T*, int component)
  {
    bool m = (component & FLAG_DIFFUSE) && (in.z >= 0.0f) && (out.z >= 0.0f);
    return Spec<T>(T(m ? kInvPi : 0.0f));
  }
  BBMCU_SCALED_EVAL
  BBMCU_D static void sample(f3 out, f2 xi, const float*, int component, f3& dir, float& pdfv, int& flag) { lambert_sample(out, xi, component, dir, pdfv, flag); }
  BBMCU_D static float pdf(f3 in, f3 out, const float*, int component) { return lambert_pdf(in, out, component); }
  BBMCU_D static Spec<float> reflectance(f3, const float* a, int component) { return (component & FLAG_DIFFUSE) ? load_spec(a) : Spec<float>(0.0f); }
};

// ---- Oren-Nayar (orennayar.h:43-86) -------------------------------------------------------------
struct OrenNayar
{
  static constexpr int NA = 4, SCALE = 0;
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_DIFFUSE) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    T s2 = a[3]*a[3];
    T A = 1.0f - 0.5f*s2/(s2 + 0.33f);
    T B = 0.45f*s2/(s2 + 0.09f);
    float cosBeta = fmaxf(in.z, out.z);
    float dxy = in.x*out.x + in.y*out.y;
    T factor = A + (B * fmaxf(dxy, 0.0f) / cosBeta);
    // result = albedo / Pi * factor: the scale is applied first in the reference; 1/Pi*factor here
    return Spec<T>(factor / kPi);
  }
  // the reference divides albedo by Pi before multiplying by factor (orennayar.h:62)
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_DIFFUSE) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    Spec<T> u = eval_unscaled<T>(in, out, a, component);
    T factor = u.r * kPi;
    return Spec<T>(a[0]/kPi*factor, a[1]/kPi*factor, a[2]/kPi*factor);
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float*, int component, f3& dir, float& pdfv, int& flag) { lambert_sample(out, xi, component, dir, pdfv, flag); }
  BBMCU_D static float pdf(f3 in, f3 out, const float*, int component) { return lambert_pdf(in, out, component); }
  BBMCU_D static Spec<float> reflectance(f3, const float* a, int component) { return (component & FLAG_DIFFUSE) ? load_spec(a) : Spec<float>(0.0f); }
};

// sample a cosine-power lobe about +Z: phi = 2 pi xi0, cos = xi1^(1/(n+1))   (phong.h:90-94)
BBMCU_D f3 sample_power_lobe(f2 xi, float n)
{
  float ph = xi.x * kTwoPi;
  float cosT = (float)pow((double)xi.y, 1.0 / (double)(n + 1.0f));
  float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
  return make_f3(cosf(ph)*sinT, sinf(ph)*sinT, cosT);
}

// ---- Phong / NganBlinnPhong (phong.h:43-145) ----------------------------------------------------
struct Phong
{
  static constexpr int NA = 4, SCALE = 0;
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return Spec<T>(T(0.0f));
    float cosAlpha = fmaxf(dot(reflect_z(in), out), 0.0f);
    const T& n = a[3];
    return Spec<T>((n + 2.0f) * (0.5f*kInvPi) * m_pow(cosAlpha, n));
  }
  BBMCU_SCALED_EVAL
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return 0.0f;
    float cosAlpha = fmaxf(dot(reflect_z(in), out), 0.0f);
    return (a[3] + 1.0f) * (0.5f*kInvPi) * powf(cosAlpha, a[3]);
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    if(!(component & FLAG_SPECULAR) || !xi_valid(xi)) return;
    dir = to_global_frame(reflect_z(out), sample_power_lobe(xi, a[3]));
    pdfv = pdf(dir, out, a, component);
    flag = FLAG_SPECULAR;
  }
  BBMCU_D static Spec<float> reflectance(f3, const float* a, int component) { return (component & FLAG_SPECULAR) ? load_spec(a) : Spec<float>(0.0f); }
};

// ---- Lafortune (lafortune.h) and NganLafortune (ngan.h:54-129) ----------------------------------
// ANISO: albedo[3] Cxy[2] Cz n ; iso: albedo[3] Cxy Cz n.  NGAN adds (n+2)/(2 pi max(Cz^2,Cxy^2)^(n/2)).
template<bool ANISO, bool NGAN>
struct Lafortune
{
  static constexpr int NA = ANISO ? 7 : 6, SCALE = 0;
  static constexpr int OCZ = ANISO ? 5 : 4, ON = ANISO ? 6 : 5;
  template<class T> BBMCU_D static T ngan_norm(const T* a)
  {
    const T& n = a[ON];
    T cz2 = a[OCZ]*a[OCZ], cxy2 = ANISO ? (a[3]*a[3] + a[4]*a[4]) : a[3]*a[3];
    return (n + 2.0f) * (0.5f*kInvPi) / m_pow(m_max(cz2, cxy2), n*0.5f);
  }
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    const T& cx = a[3]; const T& cy = ANISO ? a[4] : a[3];
    T d = cx*(in.x*out.x) + cy*(in.y*out.y) + a[OCZ]*(in.z*out.z);
    T fr = m_pow(m_max(d, 0.0f), a[ON]);
    if(NGAN) fr = fr * ngan_norm(a);
    return Spec<T>(fr);
  }
  // albedo * fr first, then the Ngan factor (ngan.h:93-94): order kept for the float path
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    const T& cx = a[3]; const T& cy = ANISO ? a[4] : a[3];
    T d = cx*(in.x*out.x) + cy*(in.y*out.y) + a[OCZ]*(in.z*out.z);
    T fr = m_pow(m_max(d, 0.0f), a[ON]);
    Spec<T> r = load_spec(a) * fr;
    if(NGAN) r = r * ngan_norm(a);
    return r;
  }
  BBMCU_D static f3 lobe_axis(f3 out, const float* a)
  {
    float cx = a[3], cy = ANISO ? a[4] : a[3];
    return normalize(make_f3(cx*out.x, cy*out.y, a[OCZ]*out.z));
  }
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return 0.0f;
    float cosAlpha = fmaxf(dot(lobe_axis(out, a), in), 0.0f);
    return (a[ON] + 1.0f) / kTwoPi * powf(cosAlpha, a[ON]);
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    if(!(component & FLAG_SPECULAR) || !xi_valid(xi)) return;
    dir = to_global_frame(lobe_axis(out, a), sample_power_lobe(xi, a[ON]));
    pdfv = pdf(dir, out, a, component);
    flag = FLAG_SPECULAR;
  }
  BBMCU_D static Spec<float> reflectance(f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR)) return Spec<float>(0.0f);
    float cx = a[3], cy = ANISO ? a[4] : a[3];
    f3 c = make_f3(cx*out.x, cy*out.y, a[OCZ]*out.z);
    float nrm = powf(sqrtf(dot(c, c)), a[ON]) * kTwoPi / (a[ON] + 2.0f);
    Spec<float> r = load_spec(a) * nrm;
    if(NGAN) r = r * ngan_norm<float>(a);
    return r;
  }
};

// ---- Ward family (ward.h, wardduer.h, wardduergeislermoroder.h).  VARIANT 0 Ward, 1 Duer, 2 GM ----
template<bool ANISO, int VARIANT>
struct Ward
{
  static constexpr int NA = ANISO ? 5 : 4, SCALE = 0;
  static constexpr float kFourPi = 12.566370614359172f;     // Constants::Pi(4)
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    // non-strict horizon test: z == 0 produces Inf/NaN exactly like the reference
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return Spec<T>(T(0.0f));
    const T& rx = a[3]; const T& ry = ANISO ? a[4] : a[3];
    f3 H = in + out;
    T sx = H.x/rx, sy = H.y/ry;
    T nf, ex;
    if(VARIANT == 0)      { nf = kFourPi * sqrtf(in.z*out.z) * rx * ry;  ex = (sx*sx + sy*sy) / (H.z*H.z); }
    else if(VARIANT == 1) { nf = kFourPi * rx * ry * (in.z*out.z);        ex = (sx*sx + sy*sy) / (H.z*H.z); }
    else { float zH2 = H.z*H.z; nf = kFourPi * rx * ry * (zH2*zH2) / dot(H, H); ex = (sx*sx + sy*sy) / zH2; }
    return Spec<T>(m_exp(-ex) / nf);
  }
  BBMCU_SCALED_EVAL
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return 0.0f;
    float rx = a[3], ry = ANISO ? a[4] : a[3];
    f3 h = halfway(in, out);
    float c3 = (float)((double)h.z*(double)h.z*(double)h.z);          // powf(h.z, 3)
    float nf = kFourPi * rx * ry * dot(in, h) * c3;
    float sx = h.x/rx, sy = h.y/ry;
    float ex = (sx*sx + sy*sy) / (h.z*h.z);
    return expf(-ex) / nf;
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    if(!(component & FLAG_SPECULAR) || !xi_valid(xi)) return;
    float rx = a[3], ry = ANISO ? a[4] : a[3];
    float ph = kTwoPi * xi.x;
    float sph_, cph_; glibc_sincosf_both(ph, sph_, cph_);            // the host libm's float functions (bbmcu_libm.cuh): bit-equal directions
    float cx = cph_*rx, cy = sph_*ry;
    float rn = 1.0f / sqrtf(cx*cx + cy*cy); cx *= rn; cy *= rn;
    float qx = cx/rx, qy = cy/ry;
    float cosT = (float)(1.0 / sqrt(1.0 - (double)(glibc_logf(xi.y) / (qx*qx + qy*qy))));
    float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
    dir = reflect(out, make_f3(cx*sinT, cy*sinT, cosT));
    pdfv = pdf(dir, out, a, component);
    flag = FLAG_SPECULAR;
  }
  BBMCU_D static Spec<float> reflectance(f3, const float* a, int component) { return (component & FLAG_SPECULAR) ? load_spec(a) : Spec<float>(0.0f); }
};

// ---- Ashikhmin-Shirley specular lobe (ashikhminshirley.h) --------------------------------------
// FRESNEL policy gives NA_F leading attribute floats; then sharpness[2|1].
// SCALED = wrapped in scaledmodel (albedo[3] first).
template<class FRESNEL, bool ANISO, bool SCALED>
struct AshikhminShirley
{
  static constexpr int SCALE = SCALED ? 0 : -1;
  static constexpr int OFF_F = SCALED ? 3 : 0;
  static constexpr int OFF_N = OFF_F + FRESNEL::NA;
  static constexpr int NA = OFF_N + (ANISO ? 2 : 1);
  static constexpr float kEightPi = 25.132741228718345f;    // Constants::Pi(8)

  template<class T> BBMCU_D static T exponent(f3 h, const T* a, bool double_one)
  {
    if(!ANISO) return a[OFF_N];
    if(!(h.z < 0.99999988079071044921875f)) return T(0.0f);
    // (nu hx^2 + nv hy^2) / (1 - hz^2); the pdf variant forms 1.0 - hz^2 in double
    float den = double_one ? (float)(1.0 - (double)(h.z*h.z)) : (1.0f - h.z*h.z);
    return ((a[OFF_N]*(h.x*h.x)) + (a[OFF_N+1]*(h.y*h.y))) / den;
  }
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    f3 h = halfway(in, out);
    float hin = dot(h, in);
    float den = hin * fmaxf(in.z, out.z);
    Spec<T> F = to_spec(FRESNEL::template eval<T>(a + OFF_F, hin));
    T ex = exponent<T>(h, a, false);
    T nrm = ANISO ? m_sqrt((a[OFF_N] + 1.0f)*(a[OFF_N+1] + 1.0f)) / kEightPi : (a[OFF_N] + 1.0f) / kEightPi;
    T lobe = nrm * m_pow(h.z, ex);
    return F * lobe / den;
  }
  BBMCU_SCALED_EVAL
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return 0.0f;
    f3 h = halfway(in, out);
    float hin = dot(h, in);
    float ex = exponent<float>(h, a, true);
    float nrm = ANISO ? sqrtf((a[OFF_N] + 1.0f)*(a[OFF_N+1] + 1.0f)) / kTwoPi : (float)(((double)a[OFF_N] + 1.0) / (double)kTwoPi);
    return (float)((double)(nrm * powf(h.z, ex)) / (4.0 * (double)hin));
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    if(!(component & FLAG_SPECULAR) || !xi_valid(xi)) return;
    float cp, sp, cosT;
    if(ANISO)
    {
      float nu = a[OFF_N], nv = a[OFF_N+1];
      float ph = atanf((float)sqrt(((double)nu + 1.0) / ((double)nv + 1.0)) * tanf(xi.x * kTwoPi));
      if((xi.x > 0.25f) && (xi.x < 0.75f)) ph = ph + kPi;
      cp = cosf(ph); sp = sinf(ph);
      cosT = (float)pow((double)xi.y, 1.0 / ((double)((nu*(cp*cp)) + (nv*(sp*sp))) + 1.0));
    }
    else
    {
      float ph = xi.x * kTwoPi;
      cp = cosf(ph); sp = sinf(ph);
      cosT = (float)pow((double)xi.y, 1.0 / ((double)a[OFF_N] + 1.0));
    }
    float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
    dir = reflect(out, make_f3(cp*sinT, sp*sinT, cosT));
    pdfv = pdf(dir, out, a, component);
    flag = FLAG_SPECULAR;
  }
  BBMCU_D static Spec<float> reflectance(f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !(out.z > 0.0f)) return Spec<float>(0.0f);
    Spec<float> r = to_spec(FRESNEL::template eval<float>(a + OFF_F, out.z));
    if(SCALED) r = r * load_spec(a);
    return r;
  }
};

// ---- AshikhminShirleyFull (ashikhminshirleyfull.h): diffuseReflectance[3] then the base attributes ----
struct AshikhminShirleyFull
{
  using Base = AshikhminShirley<FresnelSchlickRGB, true, false>;
  static constexpr int NA = 3 + Base::NA, SCALE = -1;
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    if(!((in.z > 0.0f) && (out.z > 0.0f))) return Spec<T>(T(0.0f));
    Spec<T> spec = Base::template eval<T>(in, out, a + 3, component);
    if(!(component & FLAG_DIFFUSE)) return spec;
    // Scalar(1.0) - 0.5*Vec2d(z): a scalar on the LEFT of a native array is converted to the array's float
    // (backbone/native array.h:96-100), so the base is float and only pow(., 5.0) runs in double
    double si = (double)(1.0f - 0.5f*in.z), so = (double)(1.0f - 0.5f*out.z);
    float scale = (float)((1.0 - si*si*si*si*si) * (1.0 - so*so*so*so*so));
    float nrm = (float)(28.0 / (23.0 * (double)kPi));
    float ns = nrm * scale;
    return Spec<T>(ns*a[0]*(1.0f - a[3]) + spec.r, ns*a[1]*(1.0f - a[4]) + spec.g, ns*a[2]*(1.0f - a[5]) + spec.b);
  }
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component) { return eval_unscaled<T>(in, out, a, component); }
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    if(!(component & FLAG_DIFFUSE)) return Base::pdf(in, out, a + 3, component);
    if(!(component & FLAG_SPECULAR)) return lambert_pdf(in, out, component);
    float spec_albedo = (a[3] + a[4]) + a[5];
    float diff_albedo = (float)((double)((a[0] + a[1]) + a[2]) * (1.0 - (double)spec_albedo));
    float dw = (diff_albedo > kEps) ? diff_albedo / (diff_albedo + spec_albedo) : 0.0f;
    float sw = (float)(1.0 - (double)dw);
    return sw*Base::pdf(in, out, a + 3, component) + dw*lambert_pdf(in, out, component);
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    if(!(component & FLAG_DIFFUSE)) { Base::sample(out, xi, a + 3, component, dir, pdfv, flag); return; }
    if(!(component & FLAG_SPECULAR)) { lambert_sample(out, xi, component, dir, pdfv, flag); return; }
    float spec_albedo = (a[3] + a[4]) + a[5];
    float diff_albedo = ((a[0] + a[1]) + a[2]) * (1.0f - spec_albedo);
    float dw = diff_albedo / (diff_albedo + spec_albedo);
    float sw = 1.0f - dw;
    f3 sd, dd; float sp, dp; int sf, df;
    Base::sample(out, make_f2((sw > kEps) ? xi.x / sw : 0.0f, xi.y), a + 3, component, sd, sp, sf);
    lambert_sample(out, make_f2((dw > kEps) ? (xi.x - sw) / dw : 0.0f, xi.y), component, dd, dp, df);
    bool pick_spec = xi.x <= sw;
    dir = pick_spec ? sd : dd;
    flag = pick_spec ? sf : df;
    pdfv = sw*sp + dw*dp;                    // mixes the pdfs of two different directions, as the reference does
  }
  BBMCU_D static Spec<float> reflectance(f3 out, const float* a, int component)
  {
    if(!(out.z > 0.0f)) return Spec<float>(0.0f);
    Spec<float> spec = Base::reflectance(out, a + 3, component);
    if(!(component & FLAG_DIFFUSE)) return spec;
    return Spec<float>(a[0]*(1.0f - a[3]) + spec.r, a[1]*(1.0f - a[4]) + spec.g, a[2]*(1.0f - a[5]) + spec.b);
  }
};

// ---- Low smooth-surface model (lowsmooth.h): A[3] B C eta ---------------------------------------
struct LowSmooth
{
  static constexpr int NA = 6, SCALE = 0;
  BBMCU_D static float S_(float Dp2, float B, float C) { return (float)pow(1.0 + (double)(B*Dp2), -(double)C); }
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component);
  // A*S*Q: (A*S)*Q per channel
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return Spec<T>(T(0.0f));
    T S, Q; sq<T>(in, out, a, S, Q);
    return Spec<T>(a[0]*S*Q, a[1]*S*Q, a[2]*S*Q);
  }
  template<class T> BBMCU_D static void sq(f3 in, f3 out, const T* a, T& S, T& Q)
  {
    float sx = in.x + out.x, sy = in.y + out.y, dx = in.x - out.x, dy = in.y - out.y;
    float Dp2 = sx*sx + sy*sy;
    float cosD = (float)safe_sqrt_d(1.0 - 0.25*(double)(dx*dx + dy*dy));
    S = pow_s(a[3], a[4], Dp2);
    Q = fresnel_cook(a[5], cosD);
  }
  BBMCU_D static float pow_s(float B, float C, float Dp2) { return S_(Dp2, B, C); }
  template<int N> BBMCU_D static Dual<N> pow_s(const Dual<N>& B, const Dual<N>& C, float Dp2)
  { Dual<N> r = m_pow(1.0f + B*Dp2, -C); r.v = S_(Dp2, B.v, C.v); return r; }

  BBMCU_D static float md_temp(f3 out, float B, float& ro2)
  {
    ro2 = sinTheta2(out);
    float bp = B*(1.0f - ro2);
    float t = (float)(1.0 + (double)(2.0f*B*(float)(1.0 + (double)ro2)) + (double)(bp*bp));
    return t;
  }
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return 0.0f;
    float B = a[3], ro2;
    float t = md_temp(out, B, ro2);
    t = -logf(2.0f) + logf(1.0f + B*(1.0f - ro2) + m_safe_sqrt(t));
    float Md = B * kInvPi * (1.0f / t);
    float sx = in.x + out.x, sy = in.y + out.y;
    float p = (float)((double)Md / (1.0 + (double)(B*(sx*sx + sy*sy))));
    return p * in.z;
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    if(!(component & FLAG_SPECULAR) || !xi_valid(xi)) return;
    float B = a[3], ro2;
    // (lowsmooth.h:95: here the reference writes "1 - ro2" and "pow(.., 2)" with INT literals, so - unlike in pdf() - the
    // square is a float powf and only 2 B (1.0 + ro2) runs in double.  Every libm call below is the host library's float
    // function restated (bbmcu_libm.cuh): E - 2 cancels for small xi and would amplify a last-bit difference by 1e3.)
    ro2 = sinTheta2(out);
    const float bp = B*(1.0f - ro2);
    float t = (float)((1.0 + (double)(2.0f*B) * (1.0 + (double)ro2)) + (double)(bp*bp));
    t = (float)(-0x1.62e42fefa39efp-1 + (double)glibc_logf(1.0f + bp + m_safe_sqrt(t)));
    float MdPi = B * (1.0f / t);
    float E = (float)(2.0 * (double)glibc_expf(xi.x * B * (1.0f / MdPi)));
    float ri = m_safe_sqrt((E - 2.0f)*(E + 2.0f*B*ro2) / (2.0f*E*B));
    float ro = sqrtf(ro2);
    double rp = (double)(ri + ro), rm = (double)(ri - ro);
    float scale = (float)sqrt((1.0 + (double)B*(rp*rp)) / (1.0 + (double)B*(rm*rm)));
    float phi_o = glibc_atan2f(out.y, out.x); if(phi_o < 0.0f) phi_o += kTwoPi;                   // spherical::phi (core/spherical.h:42-46)
    float phi_i = (float)(2.0 * (double)glibc_atanf(glibc_tanf(xi.y * kPi) * scale) + (double)phi_o);
    float sp, cp; glibc_sincosf_both(phi_i, sp, cp);
    dir = make_f3(cp*ri, sp*ri, (float)safe_sqrt_d(1.0 - (double)(ri*ri)));
    pdfv = pdf(dir, out, a, component);
    flag = FLAG_SPECULAR;
  }
  BBMCU_D static Spec<float> reflectance(f3 out, const float* a, int component)
  {
    if(!(component & FLAG_SPECULAR) || !(out.z > 0.0f)) return Spec<float>(0.0f);
    float B = a[3], C = a[4], eta = a[5];
    float factor = (fabsf(C - 1.0f) < kEps) ? logf(B + 1.0f) / (2.0f*B)
                                           : (float)((1.0 - (double)powf(B + 1.0f, 1.0f - C)) / (double)(2.0f*B*(C - 1.0f)));
    float r = (eta - 1.0f) / (eta + 1.0f);
    float R0 = r*r;
    return Spec<float>(kTwoPi*a[0]*factor*R0, kTwoPi*a[1]*factor*R0, kTwoPi*a[2]*factor*R0);
  }
};
template<class T> BBMCU_D Spec<T> LowSmooth::eval_unscaled(f3 in, f3 out, const T* a, int component)
{
  if(!(component & FLAG_SPECULAR) || !((in.z >= 0.0f) && (out.z >= 0.0f))) return Spec<T>(T(0.0f));
  T S, Q; sq<T>(in, out, a, S, Q);
  return Spec<T>(S*Q);
}

} // namespace bbmcu
