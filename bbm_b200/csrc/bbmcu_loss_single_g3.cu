// (batched losses evaluate many parameter sets per launch: the EPD G1 rows are not launch-uniform here - read the table, not a staged copy)
#define BBMCU_EPD_NO_STAGE
// loss(+gradient) kernel instantiations, single lobe, model group 3
#include "bbmcu_losskernel.cuh"
namespace bbmcu {
bool launch_loss_single_g3(int model, cudaStream_t s, const LossArgs& a, unsigned bx, unsigned K)
{
  switch(model) {
    BBMCU_LOSS_CASE_SINGLE(M_Bagher)
    BBMCU_LOSS_CASE_SINGLE(M_EPD)
    BBMCU_LOSS_CASE_SINGLE(M_He) BBMCU_LOSS_CASE_SINGLE(M_HeWestin) BBMCU_LOSS_CASE_SINGLE(M_HeHolzschuch) BBMCU_LOSS_CASE_SINGLE(M_NganHe)
    default: return false;
  }
}
}
