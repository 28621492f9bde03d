// (batched losses evaluate many parameter sets per launch: the EPD G1 rows are not launch-uniform here - read the table, not a staged copy)
#define BBMCU_EPD_NO_STAGE
// loss(+gradient) kernel instantiations, single lobe, model group 0
#include "bbmcu_losskernel.cuh"
namespace bbmcu {
bool launch_loss_single_g0(int model, cudaStream_t s, const LossArgs& a, unsigned bx, unsigned K)
{
  switch(model) {
    BBMCU_LOSS_CASE_SINGLE(M_Lambertian)
    BBMCU_LOSS_CASE_SINGLE(M_OrenNayar)
    BBMCU_LOSS_CASE_SINGLE(M_Phong)
    BBMCU_LOSS_CASE_SINGLE(M_NganBlinnPhong)
    BBMCU_LOSS_CASE_SINGLE(M_Lafortune)
    BBMCU_LOSS_CASE_SINGLE(M_NganLafortune)
    BBMCU_LOSS_CASE_SINGLE(M_Ward)
    BBMCU_LOSS_CASE_SINGLE(M_WardDuer)
    BBMCU_LOSS_CASE_SINGLE(M_WardDuerGeislerMoroder)
    BBMCU_LOSS_CASE_SINGLE(M_NganWard)
    BBMCU_LOSS_CASE_SINGLE(M_NganWardDuer)
    default: return false;
  }
}
}
