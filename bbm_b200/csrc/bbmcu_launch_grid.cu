// eval over the MERL grid with the linearizer fused into the kernel (EvalGridOp): single models and run-time lobe lists
#include "bbmcu_launch.cuh"
namespace bbmcu {
void launch_eval_grid(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, uint32_t first, float* rgb, float* in, float* out, size_t n)
{
  const float* tab = merl_lin_table_device(ctx->device);
  bool al = aligned16(rgb) && aligned16(in) && aligned16(out);
  if(launch_pair_eval_grid(ctx, s, d, component, tab, first, rgb, in, out, n, al)) return;
  launch_bsdf_op<EvalGridOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.lin_tab = tab; op.first = first; op.rgb = rgb; op.in = in; op.out = out; op.n = n; op.aligned = al; });
}
}
