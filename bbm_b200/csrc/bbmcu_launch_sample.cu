#include "bbmcu_launch.cuh"
namespace bbmcu {
void launch_sample(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* out, const float* xi, float* dir, float* pdf, int32_t* flag, size_t n)
{
  bool al = aligned16(out) && aligned16(xi) && aligned16(dir) && aligned16(pdf) && aligned16(flag);
  if(launch_pair_sample(ctx, s, d, component, out, xi, dir, pdf, flag, n, al)) return;
  launch_bsdf_op<SampleOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.out = out; op.xi = xi; op.dir = dir; op.pdf = pdf; op.flag = flag; op.n = n; op.aligned = al; });
}
}
