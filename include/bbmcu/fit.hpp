// bbmcu/fit.hpp - io::importFIT / io::exportFIT (include/io/fit.h:34-77) over the C ABI
#pragma once
#include <map>
#include "bsdf.hpp"

namespace bbmcu {

inline std::map<std::string, cuda_bsdf> importFIT(const context& ctx, const std::string& filename)
{
  bbmcu_fit* f = nullptr;
  check(bbmcu_fit_import(ctx.get(), filename.c_str(), &f), ctx.get());
  std::shared_ptr<bbmcu_fit> guard(f, bbmcu_fit_free);
  std::map<std::string, cuda_bsdf> out;
  for(int i = 0; i < bbmcu_fit_count(f); ++i)
  {
    bbmcu_bsdf* b = nullptr;
    check(bbmcu_fit_bsdf(f, i, &b));
    out.emplace(bbmcu_fit_key(f, i), cuda_bsdf::adopt(ctx, b));
  }
  return out;
}

inline void exportFIT(const context& ctx, const std::string& filename, const std::map<std::string, cuda_bsdf>& data, const std::string& comment = "")
{
  bbmcu_fit* f = nullptr;
  check(bbmcu_fit_create(&f));
  std::shared_ptr<bbmcu_fit> guard(f, bbmcu_fit_free);
  for(auto& [key, b] : data) check(bbmcu_fit_add(f, key.c_str(), b.get()));
  check(bbmcu_fit_export(ctx.get(), f, filename.c_str(), comment.c_str()), ctx.get());
}

} // namespace bbmcu
