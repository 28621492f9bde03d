#include "bbmcu_launch.cuh"
namespace bbmcu {
bool launch_pair_sample(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* out, const float* xi, float* dir, float* pdf, int32_t* flag, size_t n, bool al)
{
  return launch_pair_op<SampleOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.out = out; op.xi = xi; op.dir = dir; op.pdf = pdf; op.flag = flag; op.n = n; op.aligned = al; });
}
}
