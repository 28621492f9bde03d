// fused sample -> eval -> pdf pass of the hand-merged single-lobe models, eval written before its RGB scale (host-pointer path)
#include "bbmcu_launch.cuh"
namespace bbmcu {
bool launch_sample_eval_pdf_gray_capable(const BsdfDesc& d)
{
  if(d.aggregate || d.n_lobes != 1) return false;
  bool ok = false;
  dispatch_model_host(d.model[0], [&](auto* m) { using M = typename std::remove_pointer<decltype(m)>::type; ok = HandFused<M>::value; });
  return ok;
}
bool launch_sample_eval_pdf_gray(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* out, const float* xi,
                                 float* dir, float* spdf, int32_t* flag, float* gray, float* pdf, size_t n)
{
  if(d.aggregate || d.n_lobes != 1) return false;
  const bool al = aligned16(out) && aligned16(xi) && aligned16(dir) && aligned16(spdf) && aligned16(flag) && aligned16(gray) && aligned16(pdf);
  bool launched = false;
  dispatch_model_host(d.model[0], [&](auto* m) {
    using M = typename std::remove_pointer<decltype(m)>::type;
    if constexpr (HandFused<M>::value)
    {
      SampleEvalPdfGrayOp<BsdfSingle<M>> op; op.bsdf = d; op.component = component; op.out = out; op.xi = xi; op.dir = dir; op.spdf = spdf;
      op.flag = flag; op.gray = gray; op.pdf = pdf; op.n = n; op.aligned = al;
      launch_foreach4(ctx, s, op, n);
      launched = true;
    }
  });
  return launched;
}
}
