// BSDF object (host) -> kernel descriptor
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include "bbmcu_host.hpp"
#include "bbmcu_bsdf.cuh"

namespace bbmcu {

// device copy of a MERL table for `device` (uploaded on first use, released with the MerlData); implemented by the
// CUDA side (bbmcu_api.cu) - the host-compiled tests return the host table instead
const float* merl_device_table(const bbmcu_host::MerlData& m, int device);

inline BsdfDesc make_desc(const bbmcu_host::Bsdf& b, int device = 0)      // device < 0: validate only, resolve no device pointers
{
  BsdfDesc d;
  std::memset(&d, 0, sizeof(d));
  if((int)b.lobes.size() > kMaxLobes) throw std::invalid_argument("BBM: at most " + std::to_string(kMaxLobes) + " lobes per BSDF are supported on the CUDA backbone");
  int need = 0;
  for(auto& l : b.lobes) need += (int)l.values.size() + (l.merl ? 2 : 0) + table_floats_of(l.model->id);
  if(need > kMaxAttrs) throw std::invalid_argument("BBM: attribute block exceeds " + std::to_string(kMaxAttrs) + " floats");
  d.n_lobes = (int)b.lobes.size();
  d.aggregate = b.aggregate ? 1 : 0;
  int off = 0;
  for(int l=0; l < d.n_lobes; ++l)
  {
    d.model[l] = b.lobes[l].model->id;
    d.offset[l] = off;
    for(double v : b.lobes[l].values) d.attrs[off++] = (float)v;
    if(b.lobes[l].merl)
    {
      // the measured model's "attributes" are the two halves of the table's device address (bbmcu_merl.cuh)
      const uint64_t p = device >= 0 ? (uint64_t)reinterpret_cast<uintptr_t>(merl_device_table(*b.lobes[l].merl, device)) : 0;
      const uint32_t lo = (uint32_t)p, hi = (uint32_t)(p >> 32);
      std::memcpy(&d.attrs[off], &lo, 4); std::memcpy(&d.attrs[off + 1], &hi, 4);
      off += 2;
    }
    const int nt = table_floats_of(d.model[l]);          // filled by the kernel prologue (bsdf_tables_phase1/2)
    if(nt) { ++d.n_tables; off += nt; }
  }
  d.n_floats = off;
  return d;
}

} // namespace bbmcu
