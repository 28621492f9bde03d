// BSDF object (host) -> kernel descriptor
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include "bbmcu_host.hpp"
#include "bbmcu_bsdf.cuh"

namespace bbmcu {

inline BsdfDesc make_desc(const bbmcu_host::Bsdf& b)
{
  BsdfDesc d;
  std::memset(&d, 0, sizeof(d));
  if((int)b.lobes.size() > kMaxLobes) throw std::invalid_argument("BBM: at most " + std::to_string(kMaxLobes) + " lobes per BSDF are supported on the CUDA backbone");
  int need = 0;
  for(auto& l : b.lobes) need += (int)l.values.size() + table_floats_of(l.model->id);
  if(need > kMaxAttrs) throw std::invalid_argument("BBM: attribute block exceeds " + std::to_string(kMaxAttrs) + " floats");
  d.n_lobes = (int)b.lobes.size();
  d.aggregate = b.aggregate ? 1 : 0;
  int off = 0;
  for(int l=0; l < d.n_lobes; ++l)
  {
    d.model[l] = b.lobes[l].model->id;
    d.offset[l] = off;
    for(double v : b.lobes[l].values) d.attrs[off++] = (float)v;
    const int nt = table_floats_of(d.model[l]);          // filled by the kernel prologue (bsdf_tables_phase1/2)
    if(nt) { ++d.n_tables; off += nt; }
  }
  d.n_floats = off;
  return d;
}

} // namespace bbmcu
