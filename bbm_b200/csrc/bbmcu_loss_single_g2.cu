// (batched losses evaluate many parameter sets per launch: the EPD G1 rows are not launch-uniform here - read the table, not a staged copy)
#define BBMCU_EPD_NO_STAGE
// loss(+gradient) kernel instantiations, single lobe, model group 2
#include "bbmcu_losskernel.cuh"
namespace bbmcu {
bool launch_loss_single_g2(int model, cudaStream_t s, const LossArgs& a, unsigned bx, unsigned K)
{
  switch(model) {
    BBMCU_LOSS_CASE_SINGLE(M_GGX)
    BBMCU_LOSS_CASE_SINGLE(M_GGXHeitz)
    BBMCU_LOSS_CASE_SINGLE(M_PhongWalter)
    BBMCU_LOSS_CASE_SINGLE(M_LowMicrofacet)
    BBMCU_LOSS_CASE_SINGLE(M_LowMicrofacetFit)
    BBMCU_LOSS_CASE_SINGLE(M_LowSmooth)
    BBMCU_LOSS_CASE_SINGLE(M_Ribardiere)
    BBMCU_LOSS_CASE_SINGLE(M_RibardiereAnisotropic)
    default: return false;
  }
}
}
