// compile check of one pair instantiation (development tool)
#include "bbmcu_launch.cuh"
namespace bbmcu { const float* epd_table_device(int) { return nullptr; } }
using namespace bbmcu;
using Op = SampleEvalPdfOp<BsdfPair<Lambertian, ModelOf<M_CookTorrance>::type>>;
template __global__ void bbmcu::k_foreach4<Op>(const Op, size_t);
using Op2 = EvalOp<BsdfPair<Lambertian, ModelOf<M_GGX>::type>>;
template __global__ void bbmcu::k_foreach4<Op2>(const Op2, size_t);
