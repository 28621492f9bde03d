// EvalGridOp for Aggregate(Lambertian, M): the two-model kernels
#include "bbmcu_launch.cuh"
namespace bbmcu {
bool launch_pair_eval_grid(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* tab, uint32_t first, float* rgb, float* in, float* out, size_t n, bool al)
{
  return launch_pair_op<EvalGridOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.lin_tab = tab; op.first = first; op.rgb = rgb; op.in = in; op.out = out; op.n = n; op.aligned = al; });
}
}
