#include "bbmcu_launch.cuh"
namespace bbmcu {
bool launch_pair_sample_eval_pdf(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* out, const float* xi,
                                 float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf, size_t n, bool al)
{
  return launch_pair_op<SampleEvalPdfOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.out = out; op.xi = xi; op.dir = dir; op.spdf = spdf;
                                                                      op.flag = flag; op.rgb = rgb; op.pdf = pdf; op.n = n; op.aligned = al; });
}
}
