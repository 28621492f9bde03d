// One entry of the Holzschuch-Pacanowski renormalisation table, in the generator's own operation order
// (precompute/HolzschuchPacanowski/normalization.cpp:133-155 integralSH, :214-236 the index -> (b, c, sin theta) maps).
// Host and device: the kernel of bbmcu_hp.cu gives one thread one entry; tests/cpp/test_hpnorm.cpp runs the same
// function on the host against the generator compiled from the reference tree.  Compiled without FMA contraction.
#pragma once
#include <cmath>
#include "bbmcu_math.cuh"

namespace bbmcu {

constexpr int kHpNormN = 100;

BBMCU_D float hp_normalization_entry(int bi, int ci, int si)
{
  const float samples = (float)kHpNormN;
  // main(): inv_b = 0.1 * (bIndex + 1) / (samples - bIndex) stays a double; b2 = rcp(inv_b^2) is handed over as a float
  const double inv_b = 0.1 * (double)((float)bi + 1.0f) / (double)(samples - (float)bi);
  const float b2 = (float)(1.0 / (inv_b*inv_b));
  const float c = (samples + 1.0f) / (samples - (float)ci);
  const float sinTheta = (float)si / samples;
  const double expo = (double)(-(c + 1.0f)) * 0.5;
  float integral = 0.0f;
  const float df = (float)(0.01 * (double)kPi / 180.0);
  const float f_end = sinTheta + 1.0f;
  // float f and float running sum, double terms - as the generator
  for(float f = 1.0f - sinTheta; f <= f_end; f += df)
  {
    const double x = ((double)(f*f) - 1.0 + (double)(sinTheta*sinTheta)) / (2.0*(double)sinTheta);
    const double alpha = 2.0 * acos(fmin(1.0, fmax(-1.0, x / (double)f)));
    const double S = pow((double)(1.0f + b2*f*f), expo);
    integral = (float)((double)integral + alpha * (double)f * (double)df * S);
  }
  integral = integral * ((b2*(c - 1.0f)) * (float)(0.5 * 0.31830988618379067154));        // float: (b2 (c - 1)) InvPi(0.5)
  const float one_minus = 1.0f - sinTheta;                     // bbm::pow(1 - sin, 2) is powf: the correctly rounded square
  const double inner = pow((double)(1.0f + b2*(one_minus*one_minus)), 0.5*(double)(1.0f - c));
  return (float)((double)integral + (1.0 - inner));
}

} // namespace bbmcu
