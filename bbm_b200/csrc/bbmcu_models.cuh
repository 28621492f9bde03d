// The 34 bbm models as device types, their ids and the run-time (uniform) dispatch used for
// aggregates of arbitrary lobes.  Ids follow SURVEY.md section 8(a3); names are exactly what the
// reference's bbm_info prints (include/export/bbm_fromstring.h:48-49).
#pragma once
#include "bbmcu_lobes.cuh"
#include "bbmcu_epd.cuh"
#include "bbmcu_he.cuh"
#include "bbmcu_merl.cuh"

namespace bbmcu {

enum ModelId : int {
  M_Lambertian = 0, M_OrenNayar, M_Phong, M_NganBlinnPhong, M_Lafortune, M_NganLafortune,
  M_Ward, M_WardDuer, M_WardDuerGeislerMoroder, M_NganWard, M_NganWardDuer,
  M_AshikhminShirley, M_AshikhminShirleyFull, M_NganAshikhminShirley, M_LowAshikhminShirley,
  M_CookTorrance, M_LowCookTorrance, M_NganCookTorrance, M_CookTorranceWalter, M_CookTorranceHeitz,
  M_GGX, M_GGXHeitz, M_PhongWalter, M_LowMicrofacet, M_LowMicrofacetFit, M_LowSmooth,
  M_Ribardiere, M_RibardiereAnisotropic, M_Bagher, M_EPD, M_He, M_HeWestin, M_HeHolzschuch, M_NganHe,
  M_Merl,                       // the measured-data model (include/staticmodel/merl.h), listed after the 34 analytic ones
  M_COUNT
};

template<int ID> struct ModelOf;
#define BBMCU_MODEL(id, ...) template<> struct ModelOf<id> { using type = __VA_ARGS__; };
BBMCU_MODEL(M_Lambertian, Lambertian)
BBMCU_MODEL(M_OrenNayar, OrenNayar)
BBMCU_MODEL(M_Phong, Phong)
BBMCU_MODEL(M_NganBlinnPhong, Phong)                                   // ngan.h:43-44
BBMCU_MODEL(M_Lafortune, Lafortune<true, false>)
BBMCU_MODEL(M_NganLafortune, Lafortune<false, true>)                  // ngan.h:54-129
BBMCU_MODEL(M_Ward, Ward<true, 0>)
BBMCU_MODEL(M_WardDuer, Ward<true, 1>)
BBMCU_MODEL(M_WardDuerGeislerMoroder, Ward<true, 2>)
BBMCU_MODEL(M_NganWard, Ward<false, 0>)                               // ngan.h:30-38
BBMCU_MODEL(M_NganWardDuer, Ward<false, 1>)
BBMCU_MODEL(M_AshikhminShirley, AshikhminShirley<FresnelSchlickRGB, true, false>)
BBMCU_MODEL(M_AshikhminShirleyFull, AshikhminShirleyFull)
BBMCU_MODEL(M_NganAshikhminShirley, AshikhminShirley<FresnelSchlickR0, false, true>)   // ngan.h:157-158
BBMCU_MODEL(M_LowAshikhminShirley, AshikhminShirley<FresnelCookIor, false, true>)      // low.h:24-25
BBMCU_MODEL(M_CookTorrance, Microfacet<NdfBeckmann<false, false>, GVGroove, FresnelCookIor, 2, true>)     // cooktorrance.h:29-34
BBMCU_MODEL(M_LowCookTorrance, Microfacet<NdfBeckmann<false, false>, GVGroove, FresnelCookIor, 2, true>)  // low.h:32-33
BBMCU_MODEL(M_NganCookTorrance, Microfacet<NdfBeckmann<false, true>, GVGroove, FresnelSchlickR0, 2, true>) // ngan.h:141-147
BBMCU_MODEL(M_CookTorranceWalter, Microfacet<NdfBeckmann<false, true>, GUncorrelated, FresnelCookIor, 1, true>)
BBMCU_MODEL(M_CookTorranceHeitz, Microfacet<NdfBeckmann<true, true>, GHeightCorrelated, FresnelCookIor, 1, true>)
BBMCU_MODEL(M_GGX, Microfacet<NdfGGX<false>, GUncorrelated, FresnelCookIor, 1, true>)
BBMCU_MODEL(M_GGXHeitz, Microfacet<NdfGGX<true>, GHeightCorrelated, FresnelCookIor, 1, true>)
BBMCU_MODEL(M_PhongWalter, Microfacet<NdfPhong, GUncorrelated, FresnelCookIor, 1, true>)
BBMCU_MODEL(M_LowMicrofacet, Microfacet<NdfLow, GVGroove, FresnelCookIor, 2, true>)      // lowmicrofacet.h:38 (default Cook)
BBMCU_MODEL(M_LowMicrofacetFit, Microfacet<NdfLow, GVGroove, FresnelCookIor, 2, true>)   // low.h:40-41
BBMCU_MODEL(M_LowSmooth, LowSmooth)
BBMCU_MODEL(M_Ribardiere, Microfacet<NdfStudentT<false>, GUncorrelated, FresnelCookIor, 1, true>)
BBMCU_MODEL(M_RibardiereAnisotropic, Microfacet<NdfStudentT<true>, GUncorrelated, FresnelCookIor, 1, true>)
BBMCU_MODEL(M_Bagher, Microfacet<NdfSGD, GUncorrelated, FresnelBagher, 2, true>)          // bagher.h:62-68
BBMCU_MODEL(M_EPD, Microfacet<NdfEPD, GVanGinneken, FresnelComplexScalar, 1, false>)      // holzschuchpacanowski.h:34-42
BBMCU_MODEL(M_He, HeModel<HE_VARIANT_HE>)
BBMCU_MODEL(M_HeWestin, HeModel<HE_VARIANT_WESTIN>)
BBMCU_MODEL(M_HeHolzschuch, HeModel<HE_VARIANT_HOLZSCHUCH>)
BBMCU_MODEL(M_NganHe, HeModel<HE_VARIANT_NGAN>)
BBMCU_MODEL(M_Merl, MerlModel)

// floats of device-side tables appended to a lobe's attribute block (the He family's 90-bin sampling CDF)
template<class M> struct TableFloats { static constexpr int N = 0; };
template<int V> struct TableFloats<HeModel<V>> { static constexpr int N = HeModel<V>::NT; };
template<> struct TableFloats<MerlModel> { static constexpr int N = MerlModel::NT; };
BBMCU_HD constexpr int table_floats_of(int model) { return (model >= M_He && model <= M_Merl) ? kHeCdfBins : 0; }

// uniform (per-launch) dispatch on a model id: calls f((ModelOf<id>::type*)nullptr)
#define BBMCU_CASES_EPD BBMCU_CASE(M_EPD)
#define BBMCU_CASES_HE BBMCU_CASE(M_He) BBMCU_CASE(M_HeWestin) BBMCU_CASE(M_HeHolzschuch) BBMCU_CASE(M_NganHe) BBMCU_CASE(M_Merl)
#define BBMCU_ALL_CASES \
    BBMCU_CASE(M_Lambertian) BBMCU_CASE(M_OrenNayar) BBMCU_CASE(M_Phong) BBMCU_CASE(M_NganBlinnPhong) \
    BBMCU_CASE(M_Lafortune) BBMCU_CASE(M_NganLafortune) BBMCU_CASE(M_Ward) BBMCU_CASE(M_WardDuer) \
    BBMCU_CASE(M_WardDuerGeislerMoroder) BBMCU_CASE(M_NganWard) BBMCU_CASE(M_NganWardDuer) \
    BBMCU_CASE(M_AshikhminShirley) BBMCU_CASE(M_AshikhminShirleyFull) BBMCU_CASE(M_NganAshikhminShirley) \
    BBMCU_CASE(M_LowAshikhminShirley) BBMCU_CASE(M_CookTorrance) BBMCU_CASE(M_LowCookTorrance) \
    BBMCU_CASE(M_NganCookTorrance) BBMCU_CASE(M_CookTorranceWalter) BBMCU_CASE(M_CookTorranceHeitz) \
    BBMCU_CASE(M_GGX) BBMCU_CASE(M_GGXHeitz) BBMCU_CASE(M_PhongWalter) BBMCU_CASE(M_LowMicrofacet) \
    BBMCU_CASE(M_LowMicrofacetFit) BBMCU_CASE(M_LowSmooth) BBMCU_CASE(M_Ribardiere) \
    BBMCU_CASE(M_RibardiereAnisotropic) BBMCU_CASE(M_Bagher) BBMCU_CASES_EPD BBMCU_CASES_HE

#define BBMCU_CASE(m) case m: f((typename ModelOf<m>::type*)nullptr); break;
template<class F> BBMCU_D void dispatch_model(int id, F&& f) { switch(id) { BBMCU_ALL_CASES default: break; } }
// host-only twin (kernel launches); returns false for an id that is not compiled in
template<class F> inline bool dispatch_model_host(int id, F&& f) { switch(id) { BBMCU_ALL_CASES default: return false; } return true; }
#undef BBMCU_CASE

} // namespace bbmcu
