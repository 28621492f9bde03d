// (batched losses evaluate many parameter sets per launch: the EPD G1 rows are not launch-uniform here - read the table, not a staged copy)
#define BBMCU_EPD_NO_STAGE
// loss(+gradient) kernel instantiations, pair lobe after a Lambertian lobe, model group 3
#include "bbmcu_losskernel.cuh"
namespace bbmcu {
bool launch_loss_pair_g3(int model, cudaStream_t s, const LossArgs& a, unsigned bx, unsigned K)
{
  switch(model) {
    BBMCU_LOSS_CASE_PAIR(M_Bagher)
    BBMCU_LOSS_CASE_PAIR(M_EPD)
    BBMCU_LOSS_CASE_PAIR(M_He) BBMCU_LOSS_CASE_PAIR(M_HeWestin) BBMCU_LOSS_CASE_PAIR(M_HeHolzschuch) BBMCU_LOSS_CASE_PAIR(M_NganHe)
    default: return false;
  }
}
}
