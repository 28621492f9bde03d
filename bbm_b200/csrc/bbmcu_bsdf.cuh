// A BSDF as the kernels see it: a short list of lobes (model id + offset into one flat attribute
// block) with the reference's aggregate semantics for eval / sample / pdf / reflectance
// (include/bbm/aggregatebsdf.h:77-211; bsdfmodel/aggregatemodel.h:60-172).  A single model is
// a one-lobe list and skips the aggregate arithmetic, exactly like calling the model directly.
#pragma once
#include "bbmcu_models.cuh"

namespace bbmcu {

constexpr int kMaxLobes = 4;
constexpr int kMaxAttrs = 64;

struct BsdfDesc
{
  int n_lobes;
  int aggregate;               // 1: wrapped in Aggregate(...) (weights / sum even for one lobe)
  int model[kMaxLobes];
  int offset[kMaxLobes];       // first attribute of the lobe in attrs[]
  float attrs[kMaxAttrs];
};

// ---- single lobe helpers ----------------------------------------------------------------------
BBMCU_D Spec<float> lobe_eval(int model, const float* a, f3 in, f3 out, int component)
{
  Spec<float> r(0.0f);
  dispatch_model(model, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; r = M::template eval<float>(in, out, a, component); });
  return r;
}
BBMCU_D float lobe_pdf(int model, const float* a, f3 in, f3 out, int component)
{
  float r = 0.0f;
  dispatch_model(model, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; r = M::pdf(in, out, a, component); });
  return r;
}
BBMCU_D Spec<float> lobe_reflectance(int model, const float* a, f3 out, int component)
{
  Spec<float> r(0.0f);
  dispatch_model(model, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; r = M::reflectance(out, a, component); });
  return r;
}
BBMCU_D void lobe_sample(int model, const float* a, f3 out, f2 xi, int component, f3& dir, float& pdfv, int& flag)
{
  dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
  dispatch_model(model, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; M::sample(out, xi, a, component, dir, pdfv, flag); });
}

// ---- run-time lobe list -----------------------------------------------------------------------
struct BsdfGeneric
{
  BBMCU_D static Spec<float> eval(const BsdfDesc& b, f3 in, f3 out, int component)
  {
    if(!b.aggregate) return lobe_eval(b.model[0], b.attrs + b.offset[0], in, out, component);
    Spec<float> r(0.0f);                                   // std::accumulate from Spectrum(0) (aggregatebsdf.h:96-100)
    for(int l=0; l < b.n_lobes; ++l) r = r + lobe_eval(b.model[l], b.attrs + b.offset[l], in, out, component);
    return r;
  }
  BBMCU_D static Spec<float> reflectance(const BsdfDesc& b, f3 out, int component)
  {
    if(!b.aggregate) return lobe_reflectance(b.model[0], b.attrs + b.offset[0], out, component);
    Spec<float> r(0.0f);
    for(int l=0; l < b.n_lobes; ++l) r = r + lobe_reflectance(b.model[l], b.attrs + b.offset[l], out, component);
    return r;
  }
  BBMCU_D static float pdf(const BsdfDesc& b, f3 in, f3 out, int component)
  {
    if(!b.aggregate) return lobe_pdf(b.model[0], b.attrs + b.offset[0], in, out, component);
    float w[kMaxLobes], sum = 0.0f;
    for(int l=0; l < b.n_lobes; ++l) { w[l] = hsum(lobe_reflectance(b.model[l], b.attrs + b.offset[l], out, component)); sum += w[l]; }
    if(!(sum > kEps)) return 0.0f;
    float p = 0.0f;                                        // pdf += w * pdf_l / sum, term by term (aggregatebsdf.h:183)
    for(int l=0; l < b.n_lobes; ++l) p += w[l] * lobe_pdf(b.model[l], b.attrs + b.offset[l], in, out, component) / sum;
    return p;
  }
  BBMCU_D static void sample(const BsdfDesc& b, f3 out, f2 xi, int component, f3& dir, float& pdfv, int& flag)
  {
    if(!b.aggregate) { lobe_sample(b.model[0], b.attrs + b.offset[0], out, xi, component, dir, pdfv, flag); return; }
    // the reference returns an uninitialised sample when the weight sum is <= eps
    // (aggregatebsdf.h:104,116-117); we return {0, 0, None} there (documented deviation).
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    float w[kMaxLobes], sum = 0.0f;
    for(int l=0; l < b.n_lobes; ++l) { w[l] = hsum(lobe_reflectance(b.model[l], b.attrs + b.offset[l], out, component)); sum += w[l]; }
    if(!(sum > kEps)) return;
    float residual = xi.x * sum;
    for(int l=0; l < b.n_lobes; ++l)
    {
      if((residual >= 0.0f) && (residual <= w[l]))
      {
        float nr = (w[l] > kEps) ? residual / w[l] : 0.0f;
        lobe_sample(b.model[l], b.attrs + b.offset[l], out, make_f2(nr, xi.y), component, dir, pdfv, flag);
      }
      residual -= w[l];
    }
    float p = 0.0f;
    for(int l=0; l < b.n_lobes; ++l) p += w[l] * lobe_pdf(b.model[l], b.attrs + b.offset[l], dir, out, component) / sum;
    pdfv = p;
  }
};

// ---- compile-time single model (no dispatch, smallest register footprint) ------------------------
template<class M>
struct BsdfSingle
{
  BBMCU_D static Spec<float> eval(const BsdfDesc& b, f3 in, f3 out, int component) { return M::template eval<float>(in, out, b.attrs, component); }
  BBMCU_D static Spec<float> reflectance(const BsdfDesc& b, f3 out, int component) { return M::reflectance(out, b.attrs, component); }
  BBMCU_D static float pdf(const BsdfDesc& b, f3 in, f3 out, int component) { return M::pdf(in, out, b.attrs, component); }
  BBMCU_D static void sample(const BsdfDesc& b, f3 out, f2 xi, int component, f3& dir, float& pdfv, int& flag) { M::sample(out, xi, b.attrs, component, dir, pdfv, flag); }
};

} // namespace bbmcu
