#include "bbmcu_launch.cuh"
namespace bbmcu {
bool launch_pair_pdf(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component, const float* in, const float* out, float* pdf, size_t n, bool al)
{
  return launch_pair_op<PdfOp>(ctx, s, d, n, [&](auto& op) { op.component = component; op.in = in; op.out = out; op.pdf = pdf; op.n = n; op.aligned = al; });
}
}
