// A BSDF as the kernels see it: a short list of lobes (model id + offset into one flat attribute
// block) with the reference's aggregate semantics for eval / sample / pdf / reflectance
// (include/bbm/aggregatebsdf.h:77-211; bsdfmodel/aggregatemodel.h:60-172).  A single model is
// a one-lobe list and skips the aggregate arithmetic, exactly like calling the model directly.
#pragma once
#include "bbmcu_models.cuh"

namespace bbmcu {

constexpr int kMaxLobes = 4;
constexpr int kMaxAttrs = 256;         // attribute floats + device-side tables (90 CDF bins per He lobe)

struct BsdfDesc
{
  int n_lobes;
  int aggregate;               // 1: wrapped in Aggregate(...) (weights / sum even for one lobe)
  int n_tables;                // lobes whose block ends in a table the kernel prologue has to fill (He family)
  int n_floats;                // used floats of attrs[]
  int model[kMaxLobes];
  int offset[kMaxLobes];       // first attribute of the lobe in attrs[]
  float attrs[kMaxAttrs];
};

// ---- lobe lists ---------------------------------------------------------------------------------
// How lobe l of a descriptor reaches its model's code: through the switch over every model id (any BSDF string), or - for
// the shape of every entry of the reference's fits/*.fit, Aggregate(M0, M1) - through a two-way branch between two models
// known at compile time (the kernel then holds two models instead of thirty-five: 5-10 x faster, 80 instead of 255 registers).
struct DispatchAll
{
  static constexpr int kMinBlocks = 1;                     // every model's code behind one switch: let it have its registers
  static constexpr int kLobes = 0;                         // taken from the descriptor
  static constexpr bool kTables = true;
  template<class F> BBMCU_D static void apply(const BsdfDesc& b, int l, F&& f) { dispatch_model(b.model[l], f); }
};
template<class M0, class M1>
struct DispatchPair
{
  static constexpr int kMinBlocks = 3;                     // swept like the single-model kernels: 3 >= 2 = 1 (+0..18 %), fused Ward -1 %
  static constexpr int kLobes = 2;
  static constexpr bool kTables = false;                   // pairs with a He-family or measured lobe stay on the run-time path
  template<class F> BBMCU_D static void apply(const BsdfDesc&, int l, F&& f) { if(l == 0) f((M0*)nullptr); else f((M1*)nullptr); }
};

// the reference's aggregate semantics over a lobe list (aggregatebsdf.h:77-211)
template<class Dsp>
struct BsdfLobes
{
  static constexpr int kMinBlocks = Dsp::kMinBlocks, kMinBlocksFused = Dsp::kMinBlocks;
  static constexpr bool kTables = Dsp::kTables;
  static constexpr bool kAggregatePdfFromSample = true;
  BBMCU_D static Spec<float> lobe_eval(const BsdfDesc& b, int l, f3 in, f3 out, int component)
  {
    Spec<float> r(0.0f); const float* a = b.attrs + b.offset[l];
    Dsp::apply(b, l, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; r = M::template eval<float>(in, out, a, component); });
    return r;
  }
  BBMCU_D static float lobe_pdf(const BsdfDesc& b, int l, f3 in, f3 out, int component)
  {
    float r = 0.0f; const float* a = b.attrs + b.offset[l];
    Dsp::apply(b, l, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; r = M::pdf(in, out, a, component); });
    return r;
  }
  BBMCU_D static Spec<float> lobe_reflectance(const BsdfDesc& b, int l, f3 out, int component)
  {
    Spec<float> r(0.0f); const float* a = b.attrs + b.offset[l];
    Dsp::apply(b, l, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; r = M::reflectance(out, a, component); });
    return r;
  }
  BBMCU_D static void lobe_sample(const BsdfDesc& b, int l, f3 out, f2 xi, int component, f3& dir, float& pdfv, int& flag)
  {
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE; const float* a = b.attrs + b.offset[l];
    Dsp::apply(b, l, [&](auto* tag) { using M = typename std::remove_pointer<decltype(tag)>::type; M::sample(out, xi, a, component, dir, pdfv, flag); });
  }
  static constexpr bool kFusedSample = false;
  static constexpr bool kHandFused = false;
  BBMCU_D static void sample_dir(const BsdfDesc&, f3, f2, int, f3&, int&) {}
  BBMCU_D static void eval_pdf(const BsdfDesc&, f3, f3, int, Spec<float>&, float&) {}
  // f(l) for every lobe: unrolled over a compile-time count (the weights then live in registers and DispatchPair's branch
  // folds away), a plain loop over the descriptor's count otherwise
  template<class F> BBMCU_D static void for_lobes(const BsdfDesc& b, F&& f)
  {
    if constexpr (Dsp::kLobes > 0) {
#pragma unroll
      for(int l=0; l < Dsp::kLobes; ++l) f(l);
    }
    else for(int l=0; l < b.n_lobes; ++l) f(l);
  }
  BBMCU_D static Spec<float> eval(const BsdfDesc& b, f3 in, f3 out, int component)
  {
    if(!b.aggregate) return lobe_eval(b, 0, in, out, component);
    Spec<float> r(0.0f);                                   // std::accumulate from Spectrum(0) (aggregatebsdf.h:96-100)
    for_lobes(b, [&](int l) { r = r + lobe_eval(b, l, in, out, component); });
    return r;
  }
  BBMCU_D static Spec<float> reflectance(const BsdfDesc& b, f3 out, int component)
  {
    if(!b.aggregate) return lobe_reflectance(b, 0, out, component);
    Spec<float> r(0.0f);
    for_lobes(b, [&](int l) { r = r + lobe_reflectance(b, l, out, component); });
    return r;
  }
  BBMCU_D static float pdf(const BsdfDesc& b, f3 in, f3 out, int component)
  {
    if(!b.aggregate) return lobe_pdf(b, 0, in, out, component);
    float w[kMaxLobes], sum = 0.0f;
    for_lobes(b, [&](int l) { w[l] = hsum(lobe_reflectance(b, l, out, component)); sum += w[l]; });
    if(!(sum > kEps)) return 0.0f;
    float p = 0.0f;                                        // pdf += w * pdf_l / sum, term by term (aggregatebsdf.h:183)
    for_lobes(b, [&](int l) { p += w[l] * lobe_pdf(b, l, in, out, component) / sum; });
    return p;
  }
  BBMCU_D static void sample(const BsdfDesc& b, f3 out, f2 xi, int component, f3& dir, float& pdfv, int& flag)
  {
    if(!b.aggregate) { lobe_sample(b, 0, out, xi, component, dir, pdfv, flag); return; }
    // the reference returns an uninitialised sample when the weight sum is <= eps
    // (aggregatebsdf.h:104,116-117); we return {0, 0, None} there (documented deviation).
    dir = make_f3(0, 0, 0); pdfv = 0.0f; flag = FLAG_NONE;
    float w[kMaxLobes], sum = 0.0f;
    for_lobes(b, [&](int l) { w[l] = hsum(lobe_reflectance(b, l, out, component)); sum += w[l]; });
    if(!(sum > kEps)) return;
    float residual = xi.x * sum;
    for_lobes(b, [&](int l) {
      if((residual >= 0.0f) && (residual <= w[l]))
      {
        float nr = (w[l] > kEps) ? residual / w[l] : 0.0f;
        lobe_sample(b, l, out, make_f2(nr, xi.y), component, dir, pdfv, flag);
      }
      residual -= w[l];
    });
    float p = 0.0f;
    for_lobes(b, [&](int l) { p += w[l] * lobe_pdf(b, l, dir, out, component) / sum; });
    pdfv = p;
  }
};
using BsdfGeneric = BsdfLobes<DispatchAll>;
template<class M0, class M1> using BsdfPair = BsdfLobes<DispatchPair<M0, M1>>;

// ---- device-side tables: the sampling CDF of He-family lobes (ndf/sampler.h:143-181) ------------------------
// phase 1: thread `tid` of `nthreads` fills the un-normalised samples; phase 2 (after a barrier): one thread per
// lobe runs the sequential partial sum + normalisation.  The host-compiled tests call both with (0, 1).
BBMCU_D void bsdf_tables_phase1(BsdfDesc& b, int component, int tid, int nthreads)
{
  for(int l=0; l < b.n_lobes; ++l)
    dispatch_model(b.model[l], [&](auto* tag) {
      using M = typename std::remove_pointer<decltype(tag)>::type;
      if constexpr (TableFloats<M>::N > 0)
      {
        float* a = b.attrs + b.offset[l];
        for(int i = tid; i < kHeCdfBins; i += nthreads) a[M::NA + i] = M::cdf_sample(a, component, i);
      }
    });
}
BBMCU_D void bsdf_tables_phase2(BsdfDesc& b, int tid)
{
  if(tid < b.n_lobes)
    dispatch_model(b.model[tid], [&](auto* tag) {
      using M = typename std::remove_pointer<decltype(tag)>::type;
      if constexpr (TableFloats<M>::N > 0) M::cdf_finish(b.attrs + b.offset[tid] + M::NA);
    });
}

// models whose sample() returns pdf(direction, out) as the sample's pdf, and expose the direction-only half
template<class M, class = void> struct HandFused { static constexpr bool value = false; };
template<class M> struct HandFused<M, typename std::enable_if<M::kHandFusedEvalPdf>::type> { static constexpr bool value = true; };
template<class M, class = void> struct SamplePdfIsPdf { static constexpr bool value = false; };
template<class M> struct SamplePdfIsPdf<M, typename std::enable_if<M::kSamplePdfIsPdf>::type> { static constexpr bool value = true; };

// ---- compile-time single model (no dispatch, smallest register footprint) ------------------------
// Resident 256-thread blocks per SM the single-model kernels are compiled for.  3 (<= 80 registers) measured best or
// within 2 % of best for every model whose kernels were swept (tools/microbench: Bagher eval 5.8 -> 9.7 G/s, Low
// microfacet sample 21 -> 38, Phong / Lafortune / Ribardiere sample + fused +10..40 %); models whose series evaluation
// loses more to spills than it gains in occupancy say so with kLaunchMinBlocks (He's Taylor series: 3.8 -> 3.2 G/s at 3).
template<class M, class = void> struct LaunchMinBlocks { static constexpr int value = 3; };
template<class M> struct LaunchMinBlocks<M, typename std::enable_if<(M::kLaunchMinBlocks > 0)>::type> { static constexpr int value = M::kLaunchMinBlocks; };

// the fused sample + eval + pdf kernel may differ (EPD: eval +16 %, sample +23 % at 3, but fused 4.1 -> 2.6 G/s)
template<class M, class = void> struct LaunchMinBlocksFused { static constexpr int value = LaunchMinBlocks<M>::value; };
template<class M> struct LaunchMinBlocksFused<M, typename std::enable_if<(M::kLaunchMinBlocksFused > 0)>::type> { static constexpr int value = M::kLaunchMinBlocksFused; };

// models whose eval has parameter-only factors worth forming once per thread (M::precompute / M::eval_pre: Student-t)
template<class M, class = void> struct ModelHasPre { static constexpr bool value = false; };
template<class M> struct ModelHasPre<M, typename std::enable_if<M::kHasPre>::type> { static constexpr bool value = true; };

template<class M>
struct BsdfSingle
{
  static constexpr bool kHasPre = ModelHasPre<M>::value;
  BBMCU_D static auto precompute(const BsdfDesc& b) { if constexpr (kHasPre) return M::precompute(b.attrs); else return 0; }
  template<class PRE> BBMCU_D static Spec<float> eval_pre(const BsdfDesc& b, const PRE& q, f3 in, f3 out, int component)
  { if constexpr (kHasPre) return M::eval_pre(in, out, b.attrs, component, q); else return M::template eval<float>(in, out, b.attrs, component); }
  template<class PRE> BBMCU_D static void eval_pdf_pre(const BsdfDesc& b, const PRE& q, f3 in, f3 out, int component, Spec<float>& e, float& p)
  { if constexpr (kHasPre && SamplePdfIsPdf<M>::value) M::eval_pdf_pre(in, out, b.attrs, component, e, p, q); }
  static constexpr int kMinBlocks = LaunchMinBlocks<M>::value, kMinBlocksFused = LaunchMinBlocksFused<M>::value;
  static constexpr bool kAggregatePdfFromSample = false;
  static constexpr bool kTables = TableFloats<M>::N > 0;
  static constexpr bool kFusedSample = SamplePdfIsPdf<M>::value;
  static constexpr bool kHandFused = HandFused<M>::value;
  BBMCU_D static void sample_dir(const BsdfDesc& b, f3 out, f2 xi, int component, f3& dir, int& flag)
  { if constexpr (kFusedSample) M::sample_dir(out, xi, b.attrs, component, dir, flag); }
  BBMCU_D static void eval_pdf(const BsdfDesc& b, f3 in, f3 out, int component, Spec<float>& e, float& p)
  { if constexpr (kFusedSample) M::eval_pdf(in, out, b.attrs, component, e, p); }
  BBMCU_D static void sample_eval_pdf_merged(const BsdfDesc& b, f3 out, f2 xi, int component, f3& dir, int& flag, Spec<float>& e, float& p)
  { if constexpr (kHandFused) M::sample_eval_pdf_merged(out, xi, b.attrs, component, dir, flag, e, p); }
  BBMCU_D static void sample_eval_pdf_merged_u(const BsdfDesc& b, f3 out, f2 xi, int component, f3& dir, int& flag, float& u, float& p)
  { if constexpr (kHandFused) M::sample_eval_pdf_merged_u(out, xi, b.attrs, component, dir, flag, u, p); }
  BBMCU_D static Spec<float> eval(const BsdfDesc& b, f3 in, f3 out, int component) { return M::template eval<float>(in, out, b.attrs, component); }
  BBMCU_D static Spec<float> reflectance(const BsdfDesc& b, f3 out, int component) { return M::reflectance(out, b.attrs, component); }
  BBMCU_D static float pdf(const BsdfDesc& b, f3 in, f3 out, int component) { return M::pdf(in, out, b.attrs, component); }
  BBMCU_D static void sample(const BsdfDesc& b, f3 out, f2 xi, int component, f3& dir, float& pdfv, int& flag) { M::sample(out, xi, b.attrs, component, dir, pdfv, flag); }
};

// ---- EPD: which BSDF types can reach the G1 table, and the prologue that stages its two rows (bbmcu_epd.cuh) ----------------
using EpdModel = typename ModelOf<M_EPD>::type;
template<class B> struct UsesEpd { static constexpr bool value = false; };
template<> struct UsesEpd<BsdfSingle<EpdModel>> { static constexpr bool value = true; };
template<> struct UsesEpd<BsdfGeneric> { static constexpr bool value = true; };
template<class M0> struct UsesEpd<BsdfPair<M0, EpdModel>> { static constexpr bool value = true; };
#if defined(__CUDACC__) && !defined(BBMCU_EPD_NO_STAGE)
// every thread of the block calls this before the first evaluation; callers __syncthreads() afterwards
__device__ __forceinline__ void epd_stage_rows(const BsdfDesc& b, int tid, int nthreads)
{
  EpdStage& st = epd_stage();
  int lobe = -1;
  for(int l = 0; l < b.n_lobes; ++l) if(b.model[l] == M_EPD) { lobe = l; break; }
  int r0 = -1, r1 = -1;
  if(lobe >= 0)
  {
    const float p = b.attrs[b.offset[lobe] + EpdModel::OFF_NDF + 1];
    const double ip = 5.0 / (double)p - 1.0;                   // the row map of G1.h, as NdfEPD::G1v forms it
    r0 = epd_clamp_index(floor(ip), kEpdRows); r1 = epd_clamp_index(ceil(ip), kEpdRows);
    const float* tab = epd_table();
    for(int i = tid; i < kEpdCols; i += nthreads) { st.rows[i] = __ldg(tab + r0*kEpdCols + i); st.rows[kEpdCols + i] = __ldg(tab + r1*kEpdCols + i); }
  }
  if(tid == 0) { st.r0 = r0; st.r1 = r1; }
}
#endif

} // namespace bbmcu
