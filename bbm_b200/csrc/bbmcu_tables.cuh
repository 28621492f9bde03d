// Per-translation-unit binding of device-resident model tables.
//
// libbbmcu.so is built without relocatable device code, so a `static __device__` pointer in a header is a
// separate symbol in every .cu that includes it.  bind_device_tables() (internal linkage on purpose: one
// copy per translation unit) points this unit's symbols at the per-device tables owned by bbmcu_api.cu.
// It is called by every host function that launches a kernel which may evaluate an EPD lobe.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include "bbmcu_epd.cuh"

namespace bbmcu {

// the 100 x 1000 Holzschuch-Pacanowski G1 table on `device` (uploaded on first use, lives for the process)
const float* epd_table_device(int device);

#ifdef __CUDACC__
static void bind_device_tables()
{
  static bool bound[64] = {};
  int dev = 0;
  if(cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) throw std::runtime_error("BBM: cudaGetDevice failed");
  if(bound[dev]) return;
  const float* p = epd_table_device(dev);
  cudaError_t e = cudaMemcpyToSymbol(g_epd_g1_dev, &p, sizeof(p));
  if(e != cudaSuccess) throw std::runtime_error(std::string("BBM: binding the EPD G1 table failed: ") + cudaGetErrorString(e));
  bound[dev] = true;
}
#endif

} // namespace bbmcu
