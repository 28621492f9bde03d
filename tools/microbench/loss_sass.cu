// Stand-alone instantiation of the headline loss kernel for SASS inspection (development tool):
//   nvcc ... -fmad=true -cubin loss_sass.cu -o _bin/loss.cubin; nvdisasm -g -c _bin/loss.cubin > _bin/loss.sass; python sass_lines.py _bin/loss.sass
#include "bbmcu_launch.cuh"
#include "bbmcu_losskernel.cuh"
namespace bbmcu { const float* epd_table_device(int) { return nullptr; } }
using namespace bbmcu;
using L = LossPair<Lambertian, ModelOf<M_CookTorrance>::type>;
template __global__ void bbmcu::k_loss_tile<L, true>(const LossArgs, int, int);
