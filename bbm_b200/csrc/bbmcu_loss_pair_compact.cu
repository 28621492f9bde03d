// compact loss(+gradient) kernels of Aggregate(Lambertian, model) for the models that have one (bbmcu_losscompact.cuh)
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
}
bool launch_loss_pair_compact(int model, cudaStream_t s, const LossArgs& a, unsigned K)
{
  switch(model) {
    case M_CookTorrance:     return try_compact<M_CookTorrance>(s, a, K);
    case M_LowCookTorrance:  return try_compact<M_LowCookTorrance>(s, a, K);
    case M_NganCookTorrance: return try_compact<M_NganCookTorrance>(s, a, K);
    case M_GGX:              return try_compact<M_GGX>(s, a, K);
    case M_LowMicrofacet:    return try_compact<M_LowMicrofacet>(s, a, K);
    case M_LowMicrofacetFit: return try_compact<M_LowMicrofacetFit>(s, a, K);
    case M_NganAshikhminShirley: return try_compact<M_NganAshikhminShirley>(s, a, K);
    case M_LowAshikhminShirley:  return try_compact<M_LowAshikhminShirley>(s, a, K);
    case M_Phong:            return try_compact<M_Phong>(s, a, K);
    case M_NganBlinnPhong:   return try_compact<M_NganBlinnPhong>(s, a, K);
    case M_NganLafortune:    return try_compact<M_NganLafortune>(s, a, K);
    case M_LowSmooth:        return try_compact<M_LowSmooth>(s, a, K);
    case M_NganWard:         return try_compact<M_NganWard>(s, a, K);
    case M_NganWardDuer:     return try_compact<M_NganWardDuer>(s, a, K);
    default: return false;
  }
}
}
