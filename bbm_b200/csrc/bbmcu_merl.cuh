// The measured MERL model: nearest-bin lookup in a 90 x 90 x 180 half/difference-angle table, with the reference's
// data-driven importance sampling.
//
// Behaviour follows (restated, not copied):
//   include/staticmodel/merl.h:78-96 (eval: component must be exactly All, z >= 0 on both sides, merl_linearizer(in, out),
//   lookup), :152-155 (reflectance = 1 for All), :224-225 (merl = ndf_sampler<merl_data, 90, 1>), :173-206 (file format;
//   the reader lives in bbmcu_host.cpp).
// The table (3 planes of 1 458 000 floats) lives in device memory, owned by the BSDF object; the lobe's attribute block
// is [pointer low word, pointer high word][90 CDF bins] - two 32-bit patterns carried in float slots, never used as numbers.
// A NaN direction pair (antipodal grazing vectors) makes the reference throw from lookup (backbone/native control.h:75);
// kernels cannot throw, so those elements evaluate to NaN (reported).
#pragma once
#include "bbmcu_ndfsampler.cuh"
#include "bbmcu_linearizer.cuh"

namespace bbmcu {

struct MerlModel : NdfSamplerCdf<MerlModel>
{
  static constexpr int NA = 2, NT = kHeCdfBins, SCALE = -1;
  BBMCU_D static const float* table(const float* a)
  {
    uint64_t p = (uint64_t)f2u(a[0]) | ((uint64_t)f2u(a[1]) << 32);
    return reinterpret_cast<const float*>(p);
  }
  BBMCU_D static Spec<float> lookup(const float* a, f3 in, f3 out, int component)
  {
    if((component & FLAG_ALL) != FLAG_ALL || !((in.z >= 0.0f) && (out.z >= 0.0f))) return Spec<float>(0.0f);
    uint32_t idx = merl_index(in, out);
    if(!(idx < kMerlBins)) { float nan = u2f(0x7fc00000u); return Spec<float>(nan); }
    const float* t = table(a);
    return Spec<float>(t[idx], t[kMerlBins + idx], t[2*kMerlBins + idx]);
  }
  template<class T> BBMCU_D static Spec<T> eval_unscaled(f3 in, f3 out, const T* a, int component)
  {
    static_assert(std::is_same<T, float>::value, "the measured model has no parameters to differentiate");
    return lookup(a, in, out, component);
  }
  template<class T> BBMCU_D static Spec<T> eval(f3 in, f3 out, const T* a, int component) { return eval_unscaled<T>(in, out, a, component); }
  BBMCU_D static Spec<float> reflectance(f3, const float*, int component) { return Spec<float>(((component & FLAG_ALL) == FLAG_ALL) ? 1.0f : 0.0f); }
  BBMCU_D static float backscatter(const float* a, int component, f3 h) { return hsum(lookup(a, h, h, component)); }
};

} // namespace bbmcu
