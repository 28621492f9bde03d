// the fused sample + eval + pdf pass over inputs drawn in the kernel (SampleEvalPdfGenOp): single models and run-time lobe lists
#include "bbmcu_launch.cuh"
namespace bbmcu {
void launch_sample_eval_pdf_gen(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, int component,
                                float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf, size_t n, bool al, const GenArgs& g)
{
  auto fill = [&](auto& op) { op.component = component; op.out = nullptr; op.xi = nullptr; op.dir = dir; op.spdf = spdf; op.flag = flag; op.rgb = rgb; op.pdf = pdf;
                              op.n = n; op.aligned = al; op.gen_seed = g.seed; op.gen_first = g.first; op.gen_out = g.out; op.gen_xi = g.xi; };
  if(launch_pair_op<SampleEvalPdfGenOp>(ctx, s, d, n, fill)) return;
  launch_bsdf_op<SampleEvalPdfGenOp>(ctx, s, d, n, fill);
}
}
