#include "bbmcu_launch.cuh"
namespace bbmcu {
void launch_pdf(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* in, const float* out, float* pdf, size_t n)
{
  bool al = aligned16(in) && aligned16(out) && aligned16(pdf);
  if(launch_pair_pdf(ctx, s, d, component, in, out, pdf, n, al)) return;
  launch_bsdf_op<PdfOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.in = in; op.out = out; op.pdf = pdf; op.n = n; op.aligned = al; });
}
void launch_reflectance(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* out, float* rgb, size_t n)
{
  bool al = aligned16(out) && aligned16(rgb);
  launch_bsdf_op<ReflectanceOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.out = out; op.rgb = rgb; op.n = n; op.aligned = al; });
}
}
