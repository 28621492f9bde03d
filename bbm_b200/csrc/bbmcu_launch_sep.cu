#include "bbmcu_launch.cuh"
namespace bbmcu {
void launch_sample_eval_pdf(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* out, const float* xi,
                            float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf, size_t n, const GenArgs& g)
{
  bool al = aligned16(out) && aligned16(xi) && aligned16(dir) && aligned16(spdf) && aligned16(flag) && aligned16(rgb) && aligned16(pdf) && aligned16(g.out) && aligned16(g.xi);
  if(g.gen) { launch_sample_eval_pdf_gen(ctx, s, d, component, dir, spdf, flag, rgb, pdf, n, al, g); return; }
  if(launch_pair_sample_eval_pdf(ctx, s, d, component, out, xi, dir, spdf, flag, rgb, pdf, n, al)) return;
  launch_bsdf_op<SampleEvalPdfOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.out = out; op.xi = xi; op.dir = dir; op.spdf = spdf;
                                                              op.flag = flag; op.rgb = rgb; op.pdf = pdf; op.n = n; op.aligned = al; });
}
}
