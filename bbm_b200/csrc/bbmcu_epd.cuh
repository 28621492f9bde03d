// Exponential Power Distribution NDF ("Holzschuch-Pacanowski" in bbm is the EPD microfacet model only,
// SURVEY.md fact 5): D, table-driven G1, and importance sampling through the inverse of the regularised
// upper incomplete gamma function.
//
// Behaviour follows (restated, not copied):
//   include/ndf/epd.h:60-76 (eval), :87-104 (sample), :117-133 (pdf), :142-152 (G1), :157-176 (normalisation)
//   include/core/precompute.h:89-97,132-141,193-199 (tab<>::interpolate: map, clamp, multilinear lerp)
//   include/precomputed/holzschuchpacanowski/G1.h (the 100 x 1000 table and its two coordinate maps)
//   include/util/gamma.h:45-150,620-660 (series / continued fraction of P and Q, MaxTerms = 100)
//   include/util/invgamma.h (DiDonato & Morris 1986 initial estimate + <= 3 Schroeder/Newton steps)
// The reference instantiates all of this with T = float; double appears only through double literals,
// which the restatement keeps.  Not restated: the Temme large-a expansion (gamma.h:155-470, a = 1/p > 20,
// i.e. p < 0.05, outside the G1 table's range); those lanes use the series / continued fraction.
//
// The G1 table is model data (bbm_b200/data/epd_g1.f32, linked into libbbmcu.so); each translation unit
// holds a device pointer to the context's copy, bound before a launch that contains an EPD lobe.
#pragma once
#include "bbmcu_microfacet.cuh"

namespace bbmcu {

constexpr int kEpdRows = 100, kEpdCols = 1000;

#ifdef __CUDACC__
static __device__ const float* g_epd_g1_dev = nullptr;
#endif
static const float* g_epd_g1_host = nullptr;            // host-compiled tests only

BBMCU_D const float* epd_table()
{
#ifdef __CUDA_ARCH__
  return g_epd_g1_dev;
#else
  return g_epd_g1_host;
#endif
}

// ---- the two G1 rows of a launch-uniform p, staged in shared memory ---------------------------------------------------
// A lookup touches the rows floor(5/p - 1) and ceil(5/p - 1) only (include/ndf/epd.h:142-152,
// include/precomputed/holzschuchpacanowski/G1.h:14-16).  In every element-wise kernel the BSDF - hence p - is the same for
// the whole launch, so the kernel prologue (epd_stage_rows, bbmcu_bsdf.cuh) copies those two rows (2 x 1000 floats, 8 KB)
// from the device table into shared memory and the eight taps of an eval come from there.  Kernels that evaluate MANY
// parameter sets per launch (the batched losses: p differs per set) are compiled with BBMCU_EPD_NO_STAGE and read the
// table - 400 KB, L2-resident - through the read-only path instead.
struct EpdStage { int r0, r1; float rows[2*kEpdCols]; };
#if defined(__CUDACC__) && !defined(BBMCU_EPD_NO_STAGE)
__device__ __forceinline__ EpdStage& epd_stage() { __shared__ EpdStage st; return st; }
#endif

// ---- tab<float, {100,1000}>::interpolate(p, t) with derivative of the bilinear patch ------------------
BBMCU_D int epd_clamp_index(double v, int dim)
{
  if(!(v == v)) return 0;
  double c = v < 0.0 ? 0.0 : ((double)(dim - 1) < v ? (double)(dim - 1) : v);
  return (int)c;
}
BBMCU_D double std_lerp(double a, double b, double t)
{
  if((a <= 0.0 && b >= 0.0) || (a >= 0.0 && b <= 0.0)) return t*b + (1.0 - t)*a;
  if(t == 1.0) return b;
  double x = a + t*(b - a);
  return ((t > 1.0) == (b > a)) ? (b < x ? x : b) : (b > x ? x : b);
}
// value of G1 at (p, t); optionally d/d(row coordinate) and d/d(column coordinate)
BBMCU_D float epd_g1_lookup(double ip, double it, float* d_ip, float* d_it)
{
  const float* tab = epd_table();
  double fr = floor(ip), cr = ceil(ip), wr = ip - fr;
  double fc = floor(it), cc = ceil(it), wc = it - fc;
  int r0 = epd_clamp_index(fr, kEpdRows), r1 = epd_clamp_index(cr, kEpdRows);
  int c0 = epd_clamp_index(fc, kEpdCols), c1 = epd_clamp_index(cc, kEpdCols);
  float v00, v01, v10, v11;
#if defined(__CUDA_ARCH__) && !defined(BBMCU_EPD_NO_STAGE)
  const EpdStage& st = epd_stage();
  if(st.r0 == r0 && st.r1 == r1)                               // uniform: the staged rows are this launch's rows
  { v00 = st.rows[c0]; v01 = st.rows[c1]; v10 = st.rows[kEpdCols + c0]; v11 = st.rows[kEpdCols + c1]; }
  else
#endif
  {
#ifdef __CUDA_ARCH__
    v00 = __ldg(tab + r0*kEpdCols + c0); v01 = __ldg(tab + r0*kEpdCols + c1); v10 = __ldg(tab + r1*kEpdCols + c0); v11 = __ldg(tab + r1*kEpdCols + c1);
#else
    v00 = tab[r0*kEpdCols + c0]; v01 = tab[r0*kEpdCols + c1]; v10 = tab[r1*kEpdCols + c0]; v11 = tab[r1*kEpdCols + c1];
#endif
  }
  float a = (float)std_lerp((double)v00, (double)v01, wc);
  float b = (float)std_lerp((double)v10, (double)v11, wc);
  if(d_ip) *d_ip = (ip >= 0.0 && ip <= (double)(kEpdRows - 1)) ? (b - a) : 0.0f;
  if(d_it) *d_it = (it >= 0.0 && it <= (double)(kEpdCols - 1)) ? (float)((1.0 - wr)*(double)(v01 - v00) + wr*(double)(v11 - v10)) : 0.0f;
  return (float)std_lerp((double)a, (double)b, wr);
}

// ---- regularised incomplete gamma functions in float (util/gamma.h, MaxTerms = 100) -----------------
// `norm` = lgammaf(a): a = 1/p is the same for every sample of a launch, so callers evaluate it once per inverse instead of
// once per series / continued fraction (three to five times per sample).  The reciprocal inside the series loop takes the
// unguarded IEEE fast path (ieee_rcp_raw: the same bits as 1.0f / x for operands that are normal and far from the
// exponent limits, bbmcu_math.cuh); the guarded form in the continued fraction measured slower than the plain operator
// (the sampler is bound by the latency of these dependent chains, not by instruction issue: profiles/r02_s5_ncu_epd_sample.txt).
BBMCU_D float epd_gamma_series_p(float a, float x, float norm)          // detail::gamma<100, true>
{
  if(!((x >= 0.0f) && (a > 0.0f))) return 0.0f;
  float ap = a + 1.0f;
  float sum = 1.0f / a, term = sum;
  bool converged = false;
  const bool fast = (a > 1e-30f) && (a < 1e29f);              // a + m, m <= 101, is then in range as well
  for(int m=1; m <= 100 && !converged; ++m, ap += 1.0f)
  {
    term *= x * (fast ? ieee_rcp_raw(ap) : 1.0f / ap);
    sum += term;
    converged = fabsf(term) < fabsf(sum)*kEps;
  }
  sum *= expf(-x + a*logf(x) - norm);
  return sum;
}
BBMCU_D float epd_gamma_cf_q(float a, float x, float norm)  // detail::Gamma<100, true> (modified Lentz)
{
  if(!((x >= 0.0f) && (a > 0.0f))) return 0.0f;
  const float tiny = 1.17549435e-38f / kEps;
  float b_k = x + 1.0f - a;
  float c = 1.0f / tiny, d = 1.0f / b_k, G = d;
  bool converged = false;
  for(int k=1; k < 100 && !converged; ++k)
  {
    float a_k = (float)k * (a - (float)k);
    b_k += 2.0f;
    d = b_k + a_k*d;  if(fabsf(d) < tiny) d = tiny;
    c = b_k + a_k/c;  if(fabsf(c) < tiny) c = tiny;
    d = 1.0f / d;
    float delta = c*d;
    G = G*delta;
    converged = fabsf(delta - 1.0f) <= kEps;
  }
  return expf(-x + a*logf(x) - norm) * G;
}
BBMCU_D void epd_gamma_pq(float a, float x, float norm, float& p, float& q)   // gamma_pq (gamma.h:620-660) without the Temme branch; norm = lgammaf(a)
{
  if(x <= a + 1.0f) { p = epd_gamma_series_p(a, x, norm); q = 1.0f - p; }
  else if(x > a + 1.0f) { q = epd_gamma_cf_q(a, x, norm); p = 1.0f - q; }
  else { p = 0.0f; q = 0.0f; }                                  // NaN x
}

// ---- DiDonato-Morris initial estimates (util/invgamma.h) ----------------------------------------------------
constexpr double kEulerGamma = 0.57721566490153286061;

BBMCU_D float epd_eq25(float a, float y)
{
  float a2 = a*a, a3 = a2*a;
  float c1 = (a - 1.0f) * logf(y);
  float c2 = (a - 1.0f) * (1.0f + c1);
  double x = (double)c1;
  float c3 = (float)((double)(a - 1.0f) * ((-0.5*x + (double)(a - 2.0f))*x + 0.5*(double)(3.0f*a - 5.0f)));
  float c4 = (float)((double)(a - 1.0f) * ((((1.0/3.0)*x + -0.5*(double)(3.0f*a - 5.0f))*x + (double)(a2 - 6.0f*a + 7.0f))*x
                                           + (double)(11.0f*a2 - 46.0f*a + 47.0f) / 6.0));
  float c5 = (float)((double)(a - 1.0f) * (((((-0.25)*x + (double)(11.0f*a - 17.0f) / 6.0)*x + (double)(-3.0f*a2 + 13.0f*a - 13.0f))*x
                                            + (double)(2.0f*a3 - 25.0f*a2 + 72.0f*a - 61.0f) * 0.5)*x
                                           + (double)(25.0f*a3 - 195.0f*a2 + 477.0f*a - 379.0f) / 12.0));
  float ry = 1.0f / y;
  return y + ((((c5*ry + c4)*ry + c3)*ry + c2)*ry + c1);
}
BBMCU_D float epd_a_less_one(float a, float p, float q)
{
  float gamma = tgammaf(a);
  float b = q * gamma;
  if((b > 0.6f) || ((double)b >= 0.45 && (double)a >= 0.3))
  {
    float u;
    if(((double)(b*q) > 1e-8) && ((double)q > 1e-5)) u = powf(p*gamma*a, 1.0f/a);
    else u = (float)exp((double)(-q / a) - kEulerGamma);
    return u / (1.0f - (u / (a + 1.0f)));
  }
  if(((double)a < 0.3) && ((double)b >= 0.35))
  {
    float t = (float)exp(-kEulerGamma - (double)b);
    float u = t * expf(t);
    return t * expf(u);
  }
  float y = -logf(b);
  if(((double)b >= 0.15) || ((double)a >= 0.3))
  {
    float u = y - ((1.0f - a) * logf(y));
    return y - ((1.0f - a) * logf(u)) - logf(1.0f + ((1.0f - a) / (1.0f + u)));
  }
  if((double)b > 0.1)
  {
    float u = y - ((1.0f - a) * logf(y));
    float num = (1.0f*u + 2.0f*(3.0f - a))*u + (2.0f - a)*(3.0f - a);
    float den = (1.0f*u + (5.0f - a))*u + 2.0f;
    return y - ((1.0f - a) * logf(u)) - logf(num / den);
  }
  return epd_eq25(a, y);
}
BBMCU_D float epd_eq31(float a, float p, float q)
{
  float sqrta = sqrtf(a);
  bool lo = (double)p < 0.5;
  float t = lo ? sqrtf(-2.0f*logf(p)) : sqrtf(-2.0f*logf(q));
  double td = (double)t;
  double num = ((0.213623493715853*td + 4.28342155967104)*td + 11.6616720288968)*td + 3.31125922108741;
  double den = (((0.3611708101884203e-1*td + 1.27364489782223)*td + 6.40691597760039)*td + 6.61053765625462)*td + 1.0;
  float s = (float)(td - num/den);
  if(lo) s = -s;
  double c0 = (double)a - 1.0/3.0 + (double)(16.0f/(810.0f*a));
  float  c1 = sqrta - 7.0f/(36.0f*sqrta) - 433.0f/(38880.0f*a*sqrta);
  double c2 = 1.0/3.0 - (double)(7.0f/(810.0f*a));
  float  c3 = 1.0f/(36.0f*sqrta) + 256.0f/(38880.0f*a*sqrta);
  float  c4 = -3.0f/(810.0f*a);
  float  c5 = 9.0f/(38880.0f*a*sqrta);
  // Horner from the top: float until the first double coefficient joins
  float h = (c5*s + c4)*s + c3;
  double w = (((double)(h*s) + c2)*(double)s + (double)c1)*(double)s + c0;
  return (float)w;
}
BBMCU_D float epd_eq33(float a, float y, float w)
{
  float u = y + ((a - 1.0f) * logf(w)) - logf(1.0f + (1.0f - a)/(1.0f + w));
  return y + ((a - 1.0f) * logf(u)) - logf(1.0f + (1.0f - a)/(1.0f + u));
}
BBMCU_D float epd_Sn(int N, float x, float a, float tol)
{
  float sum = 1.0f, partial = 1.0f;
  bool go = partial > tol;
  for(int i=1; i <= N && go; ++i)
  {
    partial *= x / (a + (float)i);
    sum += partial;
    go = partial > tol;
  }
  return sum;
}
BBMCU_D float epd_Fn(int N, float x, float a, float v) { return expf((v + x - logf(epd_Sn(N, x, a, 0.0f))) / a); }
BBMCU_D float epd_a_greater_one(float a, float p, float q, float lg, bool& converged)      // lg = lgammaf(a)
{
  float w = epd_eq31(a, p, q);
  if((a >= 500.0f) && ((double)fabsf(1.0f - w/a) < 1e-6)) { converged = true; return w; }
  if((double)p > 0.5)
  {
    if(w < 3.0f*a) return w;
    float D = fmaxf(a*(a - 1.0f), 2.0f);
    float lb = logf(q) + lg;
    if((double)lb <= (double)(-D)*2.3) return epd_eq25(a, -lb);
    return epd_eq33(a, -lb, w);
  }
  float z = w;
  if((double)w < 0.15*(double)(a + 1.0f))
  {
    float v = logf(p) + lgammaf(a + 1.0f);
    float u1 = epd_Fn(0, w, a, v), u2 = epd_Fn(1, u1, a, v), u3 = epd_Fn(2, u2, a, v);
    z = epd_Fn(3, u3, a, v);
  }
  bool done = ((double)z < 0.01*(double)(a + 1.0f)) || ((double)z > 0.7*(double)(a + 1.0f));
  if((double)z <= 0.002*(double)(a + 1.0f)) converged = true;
  if(done) return z;
  float lnSn = logf(epd_Sn(100, z, a, 1e-4f));
  float v = logf(p) + lgammaf(a + 1.0f);
  float zbar = expf((v + z - lnSn) / a);
  return zbar * (1.0f - (a*logf(zbar) - z - v + lnSn) / (a - zbar));
}

// gamma_q_inv(a, q) (invgamma.h:440-447 -> inverse, :386-417)
BBMCU_D float epd_gamma_q_inv(float a, float q)
{
  if(!((a > 0.0f) && (q > 0.0f))) return 0.0f;
  float p = 1.0f - q;
  bool converged = false;
  float x;
  const float lg = lgammaf(a);
  if(a == 1.0f) { x = -logf(q); converged = true; }
  else if(a < 1.0f) x = epd_a_less_one(a, p, q);
  else if(a > 1.0f) x = epd_a_greater_one(a, p, q, lg, converged);
  else x = 0.0f;
  for(int itr=0; itr < 3 && !converged; ++itr)
  {
    float r;
    if(a < 20.0f) r = expf(-x - lg + logf(x)*a);
    else
    {
      float lambda = x / a;
      float delta = (float)((double)lg - (((double)a - 0.5)*(double)logf(a)) + (double)a - 0.5*log(2.0*kPiD));
      float phi = lambda - 1.0f - logf(lambda);
      r = (float)(sqrt(0.5*(double)a/kPiD) * (double)expf(-a*phi - delta));
    }
    float P, Q; epd_gamma_pq(a, x, lg, P, Q);
    float t = (((double)p <= 0.5) ? (P - p) : (q - Q)) / r;
    float w = (float)(0.5 * (double)(a - 1.0f - x));
    bool m = ((double)fabsf(t) <= 0.1) && ((double)fabsf(w*t) <= 0.1);
    x *= 1.0f - (t + (m ? w*t*t : 0.0f));
  }
  return x;
}

// ---- the NDF -------------------------------------------------------------------------------------------------
struct NdfEPD
{
  static constexpr int NA = 2;         // beta, p
  static constexpr int kFusedMinBlocks = 1;      // the fused sample + eval + pdf kernel spills at 80 registers (4.1 -> 2.6 G/s)
  template<class T> BBMCU_D static T normalization(const T& beta, const T& p)
  {
    if(!(val(p) > kEps)) return T(0.0f) / (beta*beta);
    T r = p * kInvPi * m_rcp(m_tgamma(m_rcp(p)));
    return r / (beta*beta);
  }
  template<class T> BBMCU_D static T D(f3 h, const T* a)
  {
    if(!(h.z > 0.0f)) return T(0.0f);
    float c2 = h.z*h.z;
    float t2 = (1.0f - c2) / c2;
    T beta2 = a[0]*a[0];
    T nrm = normalization(a[0], a[1]);
    return nrm * m_exp(-m_pow(t2 / beta2, a[1])) / (c2*c2);
  }
  // column coordinate of the table and its derivative with respect to t = tan(theta) * beta
  BBMCU_D static double col_coord(float t, float* dcol_dt)
  {
    float L = logf(1.0f / t);
    double E = exp((double)L * 0.05);
    double ex = exp(-E);
    if(dcol_dt) *dcol_dt = (float)(1000.0 * ex * E * 0.05 / (double)t);
    return ex * 1000.0 - 1.0;
  }
  BBMCU_D static float G1v(f3 v, f3 m, float beta, float p, float* d_beta, float* d_p)
  {
    if(d_beta) *d_beta = 0.0f;
    if(d_p) *d_p = 0.0f;
    if(!((v.z > 0.0f) && (dot(v, m) > 0.0f))) return 0.0f;
    float tanT = tanTheta(v);
    float t = tanT * beta;
    double ip = 5.0 / (double)p - 1.0;
    float dcol;
    double it = col_coord(t, d_beta ? &dcol : nullptr);
    float dr, dc;
    float g = epd_g1_lookup(ip, it, d_p ? &dr : nullptr, d_beta ? &dc : nullptr);
    if(d_beta) { float s = dc * dcol * tanT; *d_beta = (s == s) ? s : 0.0f; }
    if(d_p) *d_p = dr * (-5.0f / (p*p));
    return g;
  }
  template<class T> BBMCU_D static T G1(f3 v, f3 m, const T* a);
  BBMCU_D static float pdf(f3, f3 m, const float* a)
  {
    if(!(m.z > 0.0f)) return 0.0f;
    float p = D<float>(m, a) * m.z;
    return (p > 0.0f) ? p : 0.0f;
  }
  BBMCU_D static f3 sample(f3, f2 xi, const float* a)
  {
    if(!xi_valid(xi)) return make_f3(0, 0, 0);
    float ph = kTwoPi * xi.x;
    float cp, sp; glibc_sincosf_both(ph, sp, cp);              // the host libm's sinf / cosf restated (bbmcu_libm.cuh)
    float inv_p = 1.0f / a[1];
    float tan2 = a[0]*a[0] * powf(epd_gamma_q_inv(inv_p, xi.y), inv_p);
    float cosT = (float)(1.0 / sqrt(1.0 + (double)tan2));
    float sinT = (float)safe_sqrt_d(1.0 - (double)(cosT*cosT));
    return make_f3(cp*sinT, sp*sinT, cosT);
  }
};
template<> BBMCU_D float NdfEPD::G1<float>(f3 v, f3 m, const float* a) { return G1v(v, m, a[0], a[1], nullptr, nullptr); }
template<class T> BBMCU_D T NdfEPD::G1(f3 v, f3 m, const T* a)
{
  float db, dp;
  float g = G1v(v, m, a[0].v, a[1].v, &db, &dp);
  T r; r.v = g;
#pragma unroll
  for(int i=0; i < (int)(sizeof(r.d)/sizeof(float)); ++i) r.d[i] = db*a[0].d[i] + dp*a[1].d[i];
  return r;
}

} // namespace bbmcu
