// compact loss(+gradient) kernels of Aggregate(Lambertian, model), and of the model alone, for the models that have one
// (bbmcu_losscompact.cuh)
#define BBMCU_EPD_NO_STAGE
#include "bbmcu_losscompact.cuh"
namespace bbmcu {
namespace {
template<int MODEL> bool try_compact(cudaStream_t s, const LossArgs& a, unsigned K)
{
  using M = typename ModelOf<MODEL>::type;
  if constexpr (CompactOf<M>::value) { launch_loss_compact_static<typename CompactOf<M>::type>(s, a, K); return true; }
  else return false;
}
template<int MODEL> bool try_compact_single(cudaStream_t s, const LossArgs& a, unsigned K)
{
  using M = typename ModelOf<MODEL>::type;
  if constexpr (CompactSingleOf<M>::value) { launch_loss_compact_static<typename CompactSingleOf<M>::type>(s, a, K); return true; }
  else return false;
}
}
#define BBMCU_COMPACT_MODELS(X) X(M_CookTorrance) X(M_LowCookTorrance) X(M_NganCookTorrance) X(M_GGX) X(M_LowMicrofacet) X(M_LowMicrofacetFit) \
  X(M_NganAshikhminShirley) X(M_LowAshikhminShirley) X(M_Phong) X(M_NganBlinnPhong) X(M_NganLafortune) X(M_LowSmooth) X(M_NganWard) X(M_NganWardDuer)
bool launch_loss_pair_compact(int model, cudaStream_t s, const LossArgs& a, unsigned K)
{
  switch(model) {
#define X(m) case m: return try_compact<m>(s, a, K);
    BBMCU_COMPACT_MODELS(X)
#undef X
    default: return false;
  }
}
bool launch_loss_single_compact(int model, cudaStream_t s, const LossArgs& a, unsigned K)
{
  switch(model) {
#define X(m) case m: return try_compact_single<m>(s, a, K);
    BBMCU_COMPACT_MODELS(X)
#undef X
    default: return false;
  }
}
}
