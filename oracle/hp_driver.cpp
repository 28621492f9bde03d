// TEST INFRASTRUCTURE ONLY.  C-ABI door to the reference's Holzschuch-Pacanowski table generators, compiled from the
// sources where they lie under /root/reference (oracle/Makefile -> oracle/_ref/libbbmref_hp.so): the generator's
// translation unit is included as it is, with its main() renamed, so that single table entries can be asked for
// (the generator itself only writes the whole 100 x 100 x 100 header, minutes of one core).
#define main bbm_hp_normalization_generator_main
#include "precompute/HolzschuchPacanowski/normalization.cpp"
#undef main

extern "C" {

// entry (bIndex, cIndex, sinThetaIndex) of the renormalisation table, formed exactly as the generator's main() does
// (precompute/HolzschuchPacanowski/normalization.cpp:214-236)
float ref_hp_normalization_entry(int bIndex, int cIndex, int sinThetaIndex)
{
  const size_t samples = 100;
  auto inv_b = 0.1 * (Value(bIndex)+1) / (Value(samples) - Value(bIndex));
  auto c = (Value(samples)+1) / (Value(samples) - Value(cIndex));
  auto sinTheta = Value(sinThetaIndex) / Value(samples);
  return integralSH(sinTheta, bbm::rcp(inv_b*inv_b), c);
}

}
