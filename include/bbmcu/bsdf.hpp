// bbmcu/bsdf.hpp - C++ adapter over the C ABI (include/bbmcu.h): the reference's bsdf_ptr surface on the CUDA backbone.
//
// Mirrors include/bbm/bsdf_ptr.h:20-165 / include/bbm/bsdf_base.h:76-129 of bsdfbenchmark/bbm: eval, sample, pdf,
// reflectance (defaults component = All, unit = Radiance), toString, and the four parameter enumerations of
// include/bbm/bsdf_enumerate.h:102-237.  Scalar calls take and return the reference's array-of-structs types
// (std::array<float,3>); the batched overloads take struct-of-arrays spans in host or device memory and are what
// the hot path uses.  Errors surface as the reference's exception classes (std::invalid_argument from parsers,
// std::runtime_error elsewhere; core/error.h:42-46).
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../bbmcu.h"

namespace bbmcu {

using vec3 = std::array<float, 3>;
using vec2 = std::array<float, 2>;
using spectrum = std::array<float, 3>;

enum class bsdf_flag : int { None = BBMCU_NONE, Diffuse = BBMCU_DIFFUSE, Specular = BBMCU_SPECULAR, All = BBMCU_ALL };
enum class unit_t : int { Radiance = BBMCU_RADIANCE, Importance = BBMCU_IMPORTANCE };
enum class bsdf_attr : int { DiffuseScale = 1, DiffuseParameter = 2, SpecularScale = 4, SpecularParameter = 8, Dependent = 16, All = 15 };

struct bsdf_sample { vec3 direction; float pdf; bsdf_flag flag; };          // include/bbm/bsdfsample.h

inline void check(int rc, bbmcu_ctx* ctx = nullptr)
{
  if(rc == BBMCU_OK) return;
  std::string msg = bbmcu_last_error(ctx);
  if(rc == BBMCU_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  if(rc == BBMCU_OUT_OF_RANGE) throw std::out_of_range(msg);
  throw std::runtime_error(msg);
}

// one CUDA device + stream; all calls made through it are serialised on that stream
class context
{
public:
  explicit context(int device) { bbmcu_ctx* c = nullptr; check(bbmcu_init(device, &c)); _ctx.reset(c, bbmcu_destroy); }
  context() = default;                                 // empty handle (no device touched)
  bbmcu_ctx* get() const { return _ctx.get(); }
  void synchronize() const { check(bbmcu_synchronize(get()), get()); }
  void* stream() const { return bbmcu_stream(get()); }
  uint64_t launches() const { return bbmcu_launch_count(get()); }
private:
  std::shared_ptr<bbmcu_ctx> _ctx;
};

// bbm::bsdf_ptr<floatRGB> on the CUDA backbone
class cuda_bsdf
{
public:
  cuda_bsdf() = default;
  cuda_bsdf(const context& ctx, const std::string& str) : _ctx(ctx)
  {
    bbmcu_bsdf* b = nullptr;
    check(bbmcu_bsdf_from_string(_ctx.get(), str.c_str(), &b), _ctx.get());
    _bsdf.reset(b, bbmcu_bsdf_free);
  }
  static cuda_bsdf adopt(const context& ctx, bbmcu_bsdf* b) { cuda_bsdf r; r._ctx = ctx; r._bsdf.reset(b, bbmcu_bsdf_free); return r; }
  explicit operator bool() const { return (bool)_bsdf; }
  bbmcu_bsdf* get() const { return _bsdf.get(); }
  const context& ctx() const { return _ctx; }

  // ---- concepts::bsdfmodel, one direction at a time (host convenience; n = 1 launches) ---------------------------
  spectrum eval(const vec3& in, const vec3& out, bsdf_flag component = bsdf_flag::All, unit_t unit = unit_t::Radiance) const
  { spectrum r; eval(in.data(), out.data(), 1, r.data(), component, unit); return r; }
  bsdf_sample sample(const vec3& out, const vec2& xi, bsdf_flag component = bsdf_flag::All, unit_t unit = unit_t::Radiance) const
  { bsdf_sample s; int32_t f = 0; sample(out.data(), xi.data(), 1, s.direction.data(), &s.pdf, &f, component, unit); s.flag = (bsdf_flag)f; return s; }
  float pdf(const vec3& in, const vec3& out, bsdf_flag component = bsdf_flag::All, unit_t unit = unit_t::Radiance) const
  { float r; pdf(in.data(), out.data(), 1, &r, component, unit); return r; }
  spectrum reflectance(const vec3& out, bsdf_flag component = bsdf_flag::All, unit_t unit = unit_t::Radiance) const
  { spectrum r; reflectance(out.data(), 1, r.data(), component, unit); return r; }

  // ---- batched (struct of arrays: xyz = x[n] y[n] z[n]; host or device pointers) ----------------------------------
  void eval(const float* in_xyz, const float* out_xyz, size_t n, float* rgb, bsdf_flag c = bsdf_flag::All, unit_t u = unit_t::Radiance) const
  { check(bbmcu_eval(_ctx.get(), get(), (int)c, (int)u, in_xyz, out_xyz, n, rgb), _ctx.get()); }
  void sample(const float* out_xyz, const float* xi_uv, size_t n, float* dir_xyz, float* pdf, int32_t* flag, bsdf_flag c = bsdf_flag::All, unit_t u = unit_t::Radiance) const
  { check(bbmcu_sample(_ctx.get(), get(), (int)c, (int)u, out_xyz, xi_uv, n, dir_xyz, pdf, flag), _ctx.get()); }
  void pdf(const float* in_xyz, const float* out_xyz, size_t n, float* pdf, bsdf_flag c = bsdf_flag::All, unit_t u = unit_t::Radiance) const
  { check(bbmcu_pdf(_ctx.get(), get(), (int)c, (int)u, in_xyz, out_xyz, n, pdf), _ctx.get()); }
  void reflectance(const float* out_xyz, size_t n, float* rgb, bsdf_flag c = bsdf_flag::All, unit_t u = unit_t::Radiance) const
  { check(bbmcu_reflectance(_ctx.get(), get(), (int)c, (int)u, out_xyz, n, rgb), _ctx.get()); }
  void sample_eval_pdf(const float* out_xyz, const float* xi_uv, size_t n, float* dir_xyz, float* sample_pdf, int32_t* flag, float* rgb, float* pdf,
                       bsdf_flag c = bsdf_flag::All, unit_t u = unit_t::Radiance) const
  { check(bbmcu_sample_eval_pdf(_ctx.get(), get(), (int)c, (int)u, out_xyz, xi_uv, n, dir_xyz, sample_pdf, flag, rgb, pdf), _ctx.get()); }

  // ---- bsdf_base::toString and the parameter enumeration (forward order) ---------------------------
  std::string toString() const
  {
    std::vector<char> buf(1 << 14);
    check(bbmcu_bsdf_to_string(get(), buf.data(), buf.size()));
    return std::string(buf.data());
  }
  std::vector<double> parameter_values(bsdf_attr flags = bsdf_attr::All) const { return vec(BBMCU_PARAM_VALUE, flags); }
  std::vector<double> parameter_default_values(bsdf_attr flags = bsdf_attr::All) const { return vec(BBMCU_PARAM_DEFAULT, flags); }
  std::vector<double> parameter_lower_bound(bsdf_attr flags = bsdf_attr::All) const { return vec(BBMCU_PARAM_LOWER, flags); }
  std::vector<double> parameter_upper_bound(bsdf_attr flags = bsdf_attr::All) const { return vec(BBMCU_PARAM_UPPER, flags); }
  void set_parameter_values(const std::vector<double>& v, bsdf_attr flags = bsdf_attr::All) { check(bbmcu_bsdf_set_params(get(), (int)flags, v.data(), (int)v.size())); }

private:
  std::vector<double> vec(int which, bsdf_attr flags) const
  {
    int n = 0;
    check(bbmcu_bsdf_get_params(get(), which, (int)flags, nullptr, &n));
    std::vector<double> v(n);
    check(bbmcu_bsdf_get_params(get(), which, (int)flags, v.data(), &n));
    return v;
  }
  context _ctx;
  std::shared_ptr<bbmcu_bsdf> _bsdf;
};

// bbm::bsdf_import<floatRGB>(str) (include/bbm/bsdf_import.h:22-26)
inline cuda_bsdf bsdf_import(const context& ctx, const std::string& str) { return cuda_bsdf(ctx, str); }

} // namespace bbmcu
