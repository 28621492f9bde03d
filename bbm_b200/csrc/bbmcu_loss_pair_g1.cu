// (batched losses evaluate many parameter sets per launch: the EPD G1 rows are not launch-uniform here - read the table, not a staged copy)
#define BBMCU_EPD_NO_STAGE
// loss(+gradient) kernel instantiations, pair lobe after a Lambertian lobe, model group 1
#include "bbmcu_losskernel.cuh"
namespace bbmcu {
bool launch_loss_pair_g1(int model, cudaStream_t s, const LossArgs& a, unsigned bx, unsigned K)
{
  switch(model) {
    BBMCU_LOSS_CASE_PAIR(M_AshikhminShirley)
    BBMCU_LOSS_CASE_PAIR(M_AshikhminShirleyFull)
    BBMCU_LOSS_CASE_PAIR(M_NganAshikhminShirley)
    BBMCU_LOSS_CASE_PAIR(M_LowAshikhminShirley)
    BBMCU_LOSS_CASE_PAIR(M_CookTorrance)
    BBMCU_LOSS_CASE_PAIR(M_LowCookTorrance)
    BBMCU_LOSS_CASE_PAIR(M_NganCookTorrance)
    BBMCU_LOSS_CASE_PAIR(M_CookTorranceWalter)
    BBMCU_LOSS_CASE_PAIR(M_CookTorranceHeitz)
    default: return false;
  }
}
}
