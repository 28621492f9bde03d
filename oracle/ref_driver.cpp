// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI driver around the UNMODIFIED reference (bsdfbenchmark/bbm, native backbone),
// compiled from the headers where they lie under /root/reference by oracle/Makefile
// into oracle/_ref/libbbmref_{float,double}.so.  It exposes the reference's own
// eval / sample / pdf / reflectance, parameter enumeration, both linearizers and
// the six fitting metrics over flat arrays so the parity tests, the golden-vector
// generator (oracle/gen_golden.py) and bench.py's reference arm can call them.
//
// Two things make the snapshot compile without editing it:
//  * two empty stub headers (oracle/stubs) for the blobs listed in .MISSING_LARGE_BLOBS;
//  * core/vec_transform.h:120-121 calls phi(half)/theta(half) unqualified.  The two
//    forwarding overloads below live in the namespace of the argument type
//    (backbone::array) so ADL finds them at instantiation; they forward to
//    bbm::spherical::phi/theta, which is what the author meant.  No arithmetic changes.
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <chrono>

#include "bbm.h"
#include "linearizer/merl_linearizer.h"
#include "linearizer/spherical_linearizer.h"
#include "loss/cosine_weighted_l2.h"
#include "loss/cosine_weighted_log.h"
#include "optimizer/compass.h"
#include "io/fit.h"

namespace backbone {
  // deliberately less specialised than bbm::spherical::phi/theta(const vec3d<T>&), so that inside
  // namespace bbm::spherical (where both are visible) partial ordering still picks the original.
  template<typename V> requires (sizeof(V) == 3*sizeof(typename V::value_type)) inline auto phi(const V& v)   { return bbm::spherical::phi(v); }
  template<typename V> requires (sizeof(V) == 3*sizeof(typename V::value_type)) inline auto theta(const V& v) { return bbm::spherical::theta(v); }
}

using namespace bbm;

#ifndef REF_CONFIG
#define REF_CONFIG floatRGB
#endif
#ifndef REF_REAL
#define REF_REAL float
#endif
#define CAT2(a,b) a##b
#define CAT(a,b) CAT2(a,b)
#define FN(name) CAT(CAT(bbmref_, name), REF_SUFFIX)
#ifndef REF_SUFFIX
#define REF_SUFFIX _f
#endif

using C = REF_CONFIG;
using real_t_ = REF_REAL;
BBM_IMPORT_CONFIG(C);

static thread_local std::string g_err;
#define GUARD(...) try { __VA_ARGS__; return 0; } catch(const std::exception& e) { g_err = e.what(); return 1; } catch(...) { g_err = "unknown"; return 2; }

static inline Vec3d v3(const real_t_* p) { return Vec3d(p[0], p[1], p[2]); }

template<typename F> static void par_for(size_t n, int threads, F&& f)
{
  if(threads <= 1) { f(0, n, 0); return; }
  std::vector<std::thread> pool;
  size_t chunk = (n + threads - 1) / threads;
  for(int t=0; t < threads; ++t)
  {
    size_t b = std::min(n, t*chunk), e = std::min(n, b+chunk);
    pool.emplace_back([=,&f]() { f(b, e, t); });
  }
  for(auto& t : pool) t.join();
}


extern "C" const char* FN(last_error)(void) { return g_err.c_str(); }
extern "C" int FN(sizeof_real)(void) { return sizeof(real_t_); }

// directions are AoS (n x 3), xi AoS (n x 2); one model object per thread (the
// reference's data-driven samplers keep mutable caches).
extern "C" int FN(eval)(const char* bsdf, int component, int unit, size_t n, const real_t_* in, const real_t_* out, real_t_* rgb, int threads)
{
  GUARD(
    par_for(n, threads, [&](size_t b, size_t e, int) {
      auto m = bsdf_import<C>(bsdf);
      for(size_t i=b; i < e; ++i) {
        Spectrum s = m.eval(v3(in+3*i), v3(out+3*i), BsdfFlag(component), unit_t(unit));
        rgb[3*i] = s[0]; rgb[3*i+1] = s[1]; rgb[3*i+2] = s[2];
      }
    })
  )
}

extern "C" int FN(sample)(const char* bsdf, int component, int unit, size_t n, const real_t_* out, const real_t_* xi, real_t_* dir, real_t_* pdf, int* flag, int threads)
{
  GUARD(
    par_for(n, threads, [&](size_t b, size_t e, int) {
      auto m = bsdf_import<C>(bsdf);
      for(size_t i=b; i < e; ++i) {
        BsdfSample s = m.sample(v3(out+3*i), Vec2d(xi[2*i], xi[2*i+1]), BsdfFlag(component), unit_t(unit));
        dir[3*i] = s.direction[0]; dir[3*i+1] = s.direction[1]; dir[3*i+2] = s.direction[2];
        pdf[i] = s.pdf; flag[i] = int(s.flag);
      }
    })
  )
}

extern "C" int FN(pdf)(const char* bsdf, int component, int unit, size_t n, const real_t_* in, const real_t_* out, real_t_* pdf, int threads)
{
  GUARD(
    par_for(n, threads, [&](size_t b, size_t e, int) {
      auto m = bsdf_import<C>(bsdf);
      for(size_t i=b; i < e; ++i)
        pdf[i] = m.pdf(v3(in+3*i), v3(out+3*i), BsdfFlag(component), unit_t(unit));
    })
  )
}

extern "C" int FN(reflectance)(const char* bsdf, int component, int unit, size_t n, const real_t_* out, real_t_* rgb, int threads)
{
  GUARD(
    par_for(n, threads, [&](size_t b, size_t e, int) {
      auto m = bsdf_import<C>(bsdf);
      for(size_t i=b; i < e; ++i) {
        Spectrum s = m.reflectance(v3(out+3*i), BsdfFlag(component), unit_t(unit));
        rgb[3*i] = s[0]; rgb[3*i+1] = s[1]; rgb[3*i+2] = s[2];
      }
    })
  )
}

// the BASELINE config-2 unit of work: s = sample(out, xi); e = eval(s.dir, out); p = pdf(s.dir, out)
extern "C" int FN(sample_eval_pdf)(const char* bsdf, size_t n, const real_t_* out, const real_t_* xi, real_t_* dir, real_t_* spdf, int* flag, real_t_* rgb, real_t_* pdf, int threads)
{
  GUARD(
    par_for(n, threads, [&](size_t b, size_t e, int) {
      auto m = bsdf_import<C>(bsdf);
      for(size_t i=b; i < e; ++i) {
        Vec3d o = v3(out+3*i);
        BsdfSample s = m.sample(o, Vec2d(xi[2*i], xi[2*i+1]));
        Spectrum v = m.eval(s.direction, o);
        Value p = m.pdf(s.direction, o);
        dir[3*i] = s.direction[0]; dir[3*i+1] = s.direction[1]; dir[3*i+2] = s.direction[2];
        spdf[i] = s.pdf; flag[i] = int(s.flag);
        rgb[3*i] = v[0]; rgb[3*i+1] = v[1]; rgb[3*i+2] = v[2];
        pdf[i] = p;
      }
    })
  )
}

// which: 0 current values, 1 defaults, 2 lower bound, 3 upper bound.  flag = bsdf_attr bits.
extern "C" int FN(params)(const char* bsdf, int which, int flag, double* values, int* count)
{
  GUARD(
    auto m = bsdf_import<C>(bsdf);
    int k = 0;
    auto f = bsdf_attr(flag);
    if(which == 0) { auto p = m.parameter_values(f); for(size_t i=0; i < p.size(); ++i) values[k++] = Value(p[i]); }
    else {
      auto p = (which == 1) ? m.parameter_default_values(f) : (which == 2) ? m.parameter_lower_bound(f) : m.parameter_upper_bound(f);
      for(size_t i=0; i < p.size(); ++i) values[k++] = p[i];
    }
    *count = k;
  )
}

extern "C" int FN(to_string)(const char* bsdf, char* buf, size_t cap)
{
  GUARD(
    auto m = bsdf_import<C>(bsdf);
    std::string s = m.toString();
    std::strncpy(buf, s.c_str(), cap); if(cap) buf[cap-1] = 0;
  )
}

// merl_linearizer: direction pair -> bin index (linearizer/merl_linearizer.h:94-123)
extern "C" int FN(merl_index)(size_t n, const real_t_* in, const real_t_* out, uint64_t* idx, int threads)
{
  GUARD(
    par_for(n, threads, [&](size_t b, size_t e, int) {
      merl_linearizer<C> lin;
      for(size_t i=b; i < e; ++i) idx[i] = lin(v3(in+3*i), v3(out+3*i));
    })
  )
}

// merl_linearizer: bin index -> direction pair (linearizer/merl_linearizer.h:50-83)
extern "C" int FN(merl_dirs)(size_t first, size_t n, real_t_* in, real_t_* out, int threads)
{
  GUARD(
    par_for(n, threads, [&](size_t b, size_t e, int) {
      merl_linearizer<C> lin;
      for(size_t i=b; i < e; ++i) {
        Vec3dPair d = lin(size_t(first+i));
        for(int c=0; c < 3; ++c) { in[3*i+c] = d.in[c]; out[3*i+c] = d.out[c]; }
      }
    })
  )
}

struct sph_desc { uint64_t samplesIn[2], samplesOut[2]; double startIn[2], endIn[2], startOut[2], endOut[2]; };

static spherical_linearizer<C> make_sph(const sph_desc* d)
{
  return spherical_linearizer<C>(vec2d<size_t>(d->samplesIn[0], d->samplesIn[1]), vec2d<size_t>(d->samplesOut[0], d->samplesOut[1]),
                                 Vec2d(d->startIn[0], d->startIn[1]), Vec2d(d->endIn[0], d->endIn[1]),
                                 Vec2d(d->startOut[0], d->startOut[1]), Vec2d(d->endOut[0], d->endOut[1]));
}

extern "C" int FN(spherical_dirs)(const sph_desc* d, size_t first, size_t n, real_t_* in, real_t_* out)
{
  GUARD(
    auto lin = make_sph(d);
    for(size_t i=0; i < n; ++i) {
      Vec3dPair p = lin(size_t(first+i));
      for(int c=0; c < 3; ++c) { in[3*i+c] = p.in[c]; out[3*i+c] = p.out[c]; }
    }
  )
}

extern "C" int FN(spherical_index)(const sph_desc* d, size_t n, const real_t_* in, const real_t_* out, uint64_t* idx)
{
  GUARD(
    auto lin = make_sph(d);
    for(size_t i=0; i < n; ++i) idx[i] = lin(v3(in+3*i), v3(out+3*i));
  )
}

// metric: 0 nganL2, 1 lowL2, 2 bieronL2, 3 lowLog, 4 bieronLog, 5 standardLog
// (loss/cosine_weighted_l2.h:25-34,96-105,166-176; loss/cosine_weighted_log.h:32-43,101-112,170-181)
// linearizer: d == NULL -> merl_linearizer, else spherical_linearizer(d).
// terms (may be NULL): per-sample l(idx) for idx in [first, first+n); total (may be NULL): the
// reference's own operator()() = sequential Value accumulation / N (bbm/sampledlossfunction.h:78-87).
template<typename ERR, typename LIN>
static void run_loss(const bsdf_ptr<C>& fit, const bsdf_ptr<C>& ref, const LIN& lin, size_t first, size_t n, real_t_* terms, real_t_* total, int threads)
{
  if(terms)
    par_for(n, threads, [&](size_t b, size_t e, int) {
      // one loss object per thread; bsdf_ptr eval is const and cache-free for analytic models
      sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, ERR, LIN> L(fit, ref, ERR(), lin);
      for(size_t i=b; i < e; ++i) terms[i] = L(size_t(first+i));
    });
  if(total) {
    sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, ERR, LIN> L(fit, ref, ERR(), lin);
    *total = L();
  }
}

template<typename LIN>
static void dispatch_loss(int metric, const bsdf_ptr<C>& fit, const bsdf_ptr<C>& ref, const LIN& lin, size_t first, size_t n, real_t_* terms, real_t_* total, int threads)
{
  switch(metric) {
    case 0: run_loss<nganL2_error<C>>(fit, ref, lin, first, n, terms, total, threads); break;
    case 1: run_loss<lowL2_error<C>>(fit, ref, lin, first, n, terms, total, threads); break;
    case 2: run_loss<bieronL2_error<C>>(fit, ref, lin, first, n, terms, total, threads); break;
    case 3: run_loss<lowLog_error<C>>(fit, ref, lin, first, n, terms, total, threads); break;
    case 4: run_loss<bieronLog_error<C>>(fit, ref, lin, first, n, terms, total, threads); break;
    case 5: run_loss<standardLog_error<C>>(fit, ref, lin, first, n, terms, total, threads); break;
    default: throw std::invalid_argument("unknown metric");
  }
}

extern "C" int FN(loss)(int metric, const sph_desc* d, const char* fitted, const char* reference, size_t first, size_t n, real_t_* terms, real_t_* total, int threads)
{
  GUARD(
    auto fit = bsdf_import<C>(fitted);
    auto ref = bsdf_import<C>(reference);
    if(d) dispatch_loss(metric, fit, ref, make_sph(d), first, n, terms, total, threads);
    else  dispatch_loss(metric, fit, ref, merl_linearizer<C>(), first, n, terms, total, threads);
  )
}

// K loss evaluations at K parameter vectors (row-major K x P, forward enumeration order of
// parameter_values(bsdf_attr::All) on a SINGLE model or written lobe by lobe -- see note) --
// used for finite-difference gradients and the CPU baseline of the loss pass.
// Parameters are written through the reference's own parameter_values() reference vector.
extern "C" int FN(loss_at)(int metric, const sph_desc* d, const char* fitted, const char* reference, size_t K, size_t P, const double* params, double* losses, int accumulate_double, int threads)
{
  GUARD(
    par_for(K, threads, [&](size_t b, size_t e, int) {
      auto fit = bsdf_import<C>(fitted);
      auto ref = bsdf_import<C>(reference);
      auto pv = fit.parameter_values();
      if(pv.size() != P) throw std::invalid_argument("parameter count mismatch");
      for(size_t k=b; k < e; ++k) {
        for(size_t j=0; j < P; ++j) pv[j] = Value(params[k*P + j]);
        if(!accumulate_double) {
          real_t_ tot;
          if(d) dispatch_loss(metric, fit, ref, make_sph(d), 0, 0, nullptr, &tot, 1);
          else  dispatch_loss(metric, fit, ref, merl_linearizer<C>(), 0, 0, nullptr, &tot, 1);
          losses[k] = tot;
        } else {
          size_t N = d ? size_t(make_sph(d).size()) : size_t(merl_linearizer<C>().size());
          std::vector<real_t_> t(N);
          if(d) dispatch_loss(metric, fit, ref, make_sph(d), 0, N, t.data(), nullptr, 1);
          else  dispatch_loss(metric, fit, ref, merl_linearizer<C>(), 0, N, t.data(), nullptr, 1);
          double s = 0; for(size_t i=0; i < N; ++i) s += double(t[i]);
          losses[k] = s / double(N);
        }
      }
    })
  )
}

// The reference's own compass search (optimizer/compass.h:82-140) on metric over the given
// linearizer; returns the loss after every step and the final parameters.
extern "C" int FN(compass)(int metric, const sph_desc* d, const char* fitted, const char* reference, int max_steps, double* loss_trace, int* steps_done, double* final_params, int* P, char* final_str, size_t cap)
{
  GUARD(
    auto fit = bsdf_import<C>(fitted);
    auto ref = bsdf_import<C>(reference);
    auto param = fit.parameter_values();
    auto low = fit.parameter_lower_bound();
    auto up = fit.parameter_upper_bound();
    auto run = [&](auto& loss) {
      compass opt(loss, param, low, up);
      int t = 0;
      for(; t < max_steps && !bbm::all(opt.is_converged()); ++t) loss_trace[t] = opt.step();
      *steps_done = t;
    };
    auto with_lin = [&](auto lin) {
      using LIN = decltype(lin);
      switch(metric) {
        case 0: { sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, nganL2_error<C>, LIN> L(fit, ref, {}, lin); run(L); } break;
        case 1: { sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, lowL2_error<C>, LIN> L(fit, ref, {}, lin); run(L); } break;
        case 2: { sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, bieronL2_error<C>, LIN> L(fit, ref, {}, lin); run(L); } break;
        case 3: { sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, lowLog_error<C>, LIN> L(fit, ref, {}, lin); run(L); } break;
        case 4: { sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, bieronLog_error<C>, LIN> L(fit, ref, {}, lin); run(L); } break;
        case 5: { sampledlossfunction<bsdf_ptr<C>, bsdf_ptr<C>, standardLog_error<C>, LIN> L(fit, ref, {}, lin); run(L); } break;
        default: throw std::invalid_argument("unknown metric");
      }
    };
    if(d) with_lin(make_sph(d)); else with_lin(merl_linearizer<C>());
    *P = int(param.size());
    for(size_t j=0; j < param.size(); ++j) final_params[j] = Value(param[j]);
    std::string s = fit.toString();
    std::strncpy(final_str, s.c_str(), cap); if(cap) final_str[cap-1] = 0;
  )
}

// .fit import (io/fit.h:34-77): returns "key\tBSDF-string\n" lines.
extern "C" int FN(import_fit)(const char* path, char* buf, size_t cap, int* entries)
{
  GUARD(
    std::map<std::string, bsdf_ptr<C>> m;
    io::importFIT(path, m);
    std::string s;
    for(auto& [k, v] : m) s += k + "\t" + v.toString() + "\n";
    *entries = int(m.size());
    std::strncpy(buf, s.c_str(), cap); if(cap) buf[cap-1] = 0;
  )
}

