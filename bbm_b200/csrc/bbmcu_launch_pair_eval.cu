#include "bbmcu_launch.cuh"
namespace bbmcu {
bool launch_pair_eval(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* in, const float* out, float* rgb, size_t n, bool al)
{
  return launch_pair_op<EvalOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.in = in; op.out = out; op.rgb = rgb; op.n = n; op.aligned = al; });
}
}
