// The reference's data-driven importance sampler, shared by the He family and the measured MERL model:
//   include/bbm/ndf_sampler.h:79-156   sample / pdf of a BSDF through a back-scatter "NDF" hsum(eval(h, h))
//   include/ndf/sampler.h:63-92 (sample), :102-128 (pdf), :143-181 (90-bin CDF over theta = (i/90)^2 pi/2)
//   include/util/cdf.h:36-44,80-117    partial sum, normalisation, lower-bound search, per-bin pdf
// The reference builds the CDF lazily on the host whenever the parameters change; here every thread block rebuilds it in
// shared memory in its prologue (bsdf_tables_phase1/2 in bbmcu_bsdf.cuh: 90 evaluations, one per thread), so there is no
// host mathematics, no cache to invalidate and nothing to upload.  A lobe's device block is [NA attributes][90 CDF bins].
//
// M provides NA and  static float backscatter(const float* a, int component, f3 h)  = hsum(eval(h, h)) before any scale.
#pragma once
#include "bbmcu_microfacet.cuh"

namespace bbmcu {

constexpr int kHeCdfBins = 90;

template<class M>
struct NdfSamplerCdf
{
  // un-normalised CDF sample of bin `idx` for `component` (sampler.h:143-175)
  BBMCU_D static float cdf_sample(const float* a, int component, int idx)
  {
    float q = (float)idx / (float)kHeCdfBins;
    float theta = (float)(((double)q*(double)q) * (double)kHalfPi);
    // (the host libm's sinf / cosf restated, bbmcu_libm.cuh: the CDF must reproduce the reference's entries bit for bit)
    f3 h = make_f3(1.0f*glibc_sinf(theta), 0.0f*glibc_sinf(theta), glibc_cosf(theta));
    float s = 0.0f + M::backscatter(a, component, h);
    s /= 1.0f;
    float q1 = (float)(idx + 1) / (float)kHeCdfBins;
    float theta1 = (float)(((double)q1*(double)q1) * (double)kHalfPi);
    return s * (glibc_sinf(theta1) * sqrtf(theta1));
  }
  // cdf(samples): sequential partial sum, then normalise by the last entry (util/cdf.h:36-44)
  BBMCU_D static void cdf_finish(float* cdf)
  {
    float run = cdf[0];
    for(int i=1; i < kHeCdfBins; ++i) { run = run + cdf[i]; cdf[i] = run; }
    float norm = cdf[kHeCdfBins - 1];
    for(int i=0; i < kHeCdfBins; ++i) cdf[i] = cdf[i] / norm;
  }
  BBMCU_D static float cdf_pdf(const float* cdf, int idx) { return cdf[idx] - (idx >= 1 ? cdf[idx - 1] : 0.0f); }

  // ndf::sampler::pdf(view, m)
  BBMCU_D static float h_pdf(f3 m, const float* cdf)
  {
    if(!(m.z > 0.0f)) return 0.0f;
    float theta = sph_theta(m);
    float ti = (float)((double)(sqrtf(theta / kHalfPi) * (float)kHeCdfBins) - 0.5);
    float fl = floorf(ti), ce = ceilf(ti);
    float w = ti - fl;
    // cast<size_t>(negative) wraps to a huge value on x86-64 and clamps to the LAST bin (sampler.h:119)
    int lidx = (fl < 0.0f) ? (kHeCdfBins - 1) : (fl > (float)(kHeCdfBins - 1) ? kHeCdfBins - 1 : (int)fl);
    int uidx = (ce < 0.0f) ? (kHeCdfBins - 1) : (ce > (float)(kHeCdfBins - 1) ? kHeCdfBins - 1 : (int)ce);
    float p = cdf_pdf(cdf, lidx) * (1.0f - w) + cdf_pdf(cdf, uidx) * w;
    float jac = (sqrtf(theta) * ((0.25f*kPi)*kPi) / (float)kHeCdfBins) * fabsf(sinf(theta)) * kTwoPi;
    return (jac > kEps) ? p / jac : 0.0f;
  }
  // ndf::sampler::sample(view, xi)
  BBMCU_D static f3 h_sample(f2 xi, const float* cdf)
  {
    // std::lower_bound with predicate (val < xi): first entry that is not < xi
    int lo = 0, count = kHeCdfBins;
    while(count > 0) { int step = count / 2, mid = lo + step; if(cdf[mid] < xi.x) { lo = mid + 1; count -= step + 1; } else count = step; }
    int idx = lo;
    float residual = 0.0f;
    if(idx < kHeCdfBins)
    {
      float prev = (idx >= 1) ? cdf[idx - 1] : 0.0f;
      float pdfv = cdf[idx] - prev;
      residual = (xi.x - prev) / pdfv;
    }
    double rc = (double)residual - 0.5;
    double xi_r = fabs(rc);
    double offs = 1.0 - safe_sqrt_d(1.0 - 2.0*xi_r);
    double q = ((double)idx + 0.5 + copysign(1.0, rc)*offs) / (double)kHeCdfBins;
    float theta = (float)((q*q) * (double)kHalfPi);
    float phi = kTwoPi * xi.y;
    if(theta > kHalfPi) theta = kPi - theta;
    return sph_to_vec(phi, theta);
  }
  BBMCU_D static float pdf(f3 in, f3 out, const float* a, int component)
  {
    (void)component;                                       // ndf_sampler::pdf has no component test (ndf_sampler.h:131)
    if(!((out.z > 0.0f) && (in.z > 0.0f))) return 0.0f;
    f3 h = halfway(in, out);
    float p = h_pdf(h, a + M::NA);
    return (float)((double)p / fabs(4.0 * (double)dot(out, h)));
  }
  BBMCU_D static void sample(f3 out, f2 xi, const float* a, int component, f3& dir, float& pdfv, int& flag)
  {
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    if(!xi_valid(xi) || !(out.z > 0.0f)) return;           // no Specular test (ndf_sampler.h:84-87)
    f3 h = h_sample(xi, a + M::NA);
    dir = reflect(out, h);
    pdfv = pdf(dir, out, a, component);
    flag = component;
  }
};

} // namespace bbmcu
