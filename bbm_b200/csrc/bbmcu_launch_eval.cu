#include "bbmcu_launch.cuh"
namespace bbmcu {
void launch_eval(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* in, const float* out, float* rgb, size_t n)
{
  bool al = aligned16(in) && aligned16(out) && aligned16(rgb);
  if(launch_pair_eval(ctx, s, d, component, in, out, rgb, n, al)) return;
  launch_bsdf_op<EvalOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.in = in; op.out = out; op.rgb = rgb; op.n = n; op.aligned = al; });
}
}
