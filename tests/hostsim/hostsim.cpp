// TEST INFRASTRUCTURE ONLY.  The device headers of bbm_b200/csrc compiled for the HOST with g++
// (every model function is __host__ __device__), so the model / linearizer / loss arithmetic can be
// checked against the reference on a machine without a GPU.  Nothing in the product loads this
// library; the product path is libbbmcu.so and fails without a CUDA device.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include "bbmcu_desc.hpp"
#include "bbmcu_kernels.cuh"
#include "bbmcu_lossop.cuh"
#include "bbmcu_losscompact.cuh"
#include "bbmcu_hpnorm.cuh"

using namespace bbmcu;

static thread_local std::string g_err;
#define GUARD(...) try { __VA_ARGS__; return 0; } catch(const std::exception& e) { g_err = e.what(); return 1; }

template<class Op> static void run(const Op& op_in, size_t n)
{
  Op op = op_in;
  if(op.ld == 0) op.ld = op.n;
  if constexpr (Op::kHasBsdf)
  {
    BsdfDesc b = op.bsdf;                       // what the kernel prologue does per thread block
    if(b.n_tables) { bsdf_tables_phase1(b, op.component, 0, 1); for(int l=0; l < b.n_lobes; ++l) bsdf_tables_phase2(b, l); }
    for(size_t i=0; i < n; i += kVec) op.group(i, b);
  }
  else for(size_t i=0; i < n; i += kVec) op.group(i);
}

// host-compiled tests: the "device" table of a measured model is its host table
namespace bbmcu { const float* merl_device_table(const bbmcu_host::MerlData& m, int) { return m.rgb.data(); } }

// EvalOp::group_pre and the fused pass's group_pre of a single-lobe model with parameter-only factors (Student-t), the factors
// formed once before the loop as k_foreach4 forms them once per thread
template<class M> static void run_eval_pre(const BsdfDesc& d, int component, const float* in, const float* out, size_t n, float* rgb)
{
  using B = BsdfSingle<M>;
  static_assert(B::kHasPre, "model without parameter-only factors");
  EvalOp<B> op; op.bsdf = d; op.component = component; op.in = in; op.out = out; op.rgb = rgb; op.n = n; op.aligned = false; op.ld = n;
  const auto q = B::precompute(op.bsdf);
  for(size_t i=0; i < n; i += kVec) op.group_pre(i, op.bsdf, q);
}
template<class M> static void run_sep_pre(const BsdfDesc& d, int component, const float* out, const float* xi, size_t n, float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf)
{
  using B = BsdfSingle<M>;
  SampleEvalPdfOp<B> op; op.bsdf = d; op.component = component; op.out = out; op.xi = xi; op.dir = dir; op.spdf = spdf; op.flag = flag; op.rgb = rgb; op.pdf = pdf;
  op.n = n; op.aligned = false; op.ld = n;
  const auto q = B::precompute(op.bsdf);
  for(size_t i=0; i < n; i += kVec) op.group_pre(i, op.bsdf, q);
}
static BsdfDesc single_lobe_desc(const char* bsdf)
{
  auto parsed = bbmcu_host::parse_bsdf(bsdf);
  BsdfDesc d = make_desc(parsed);
  if(d.n_lobes != 1 || d.aggregate || (d.model[0] != M_Ribardiere && d.model[0] != M_RibardiereAnisotropic)) throw std::invalid_argument("hostsim: a single Student-t lobe expected");
  return d;
}
extern "C" {
const char* hostsim_last_error() { return g_err.c_str(); }
// the EPD G1 table (bbm_b200/data/epd_g1.f32), owned by the caller
void hostsim_set_epd_table(const float* table) { g_epd_g1_host = table; }
float hostsim_gamma_q_inv(float a, float q) { return epd_gamma_q_inv(a, q); }
float hostsim_hp_normalization_entry(int bi, int ci, int si) { return hp_normalization_entry(bi, ci, si); }

int hostsim_eval(const char* bsdf, int component, const float* in, const float* out, size_t n, float* rgb)
{ GUARD( EvalOp<BsdfGeneric> op; auto parsed = bbmcu_host::parse_bsdf(bsdf); op.bsdf = make_desc(parsed); op.component = component; op.in = in; op.out = out; op.rgb = rgb; op.n = n; op.aligned = false; run(op, n); ) }

int hostsim_eval_pre(const char* bsdf, int component, const float* in, const float* out, size_t n, float* rgb)
{
  GUARD(
    BsdfDesc d = single_lobe_desc(bsdf);
    if(d.model[0] == M_Ribardiere) run_eval_pre<ModelOf<M_Ribardiere>::type>(d, component, in, out, n, rgb);
    else run_eval_pre<ModelOf<M_RibardiereAnisotropic>::type>(d, component, in, out, n, rgb);
  )
}
int hostsim_sample_eval_pdf_pre(const char* bsdf, int component, const float* out, const float* xi, size_t n, float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf)
{
  GUARD(
    BsdfDesc d = single_lobe_desc(bsdf);
    if(d.model[0] == M_Ribardiere) run_sep_pre<ModelOf<M_Ribardiere>::type>(d, component, out, xi, n, dir, spdf, flag, rgb, pdf);
    else run_sep_pre<ModelOf<M_RibardiereAnisotropic>::type>(d, component, out, xi, n, dir, spdf, flag, rgb, pdf);
  )
}

int hostsim_pdf(const char* bsdf, int component, const float* in, const float* out, size_t n, float* pdf)
{ GUARD( PdfOp<BsdfGeneric> op; auto parsed = bbmcu_host::parse_bsdf(bsdf); op.bsdf = make_desc(parsed); op.component = component; op.in = in; op.out = out; op.pdf = pdf; op.n = n; op.aligned = false; run(op, n); ) }

int hostsim_reflectance(const char* bsdf, int component, const float* out, size_t n, float* rgb)
{ GUARD( ReflectanceOp<BsdfGeneric> op; auto parsed = bbmcu_host::parse_bsdf(bsdf); op.bsdf = make_desc(parsed); op.component = component; op.out = out; op.rgb = rgb; op.n = n; op.aligned = false; run(op, n); ) }

int hostsim_sample(const char* bsdf, int component, const float* out, const float* xi, size_t n, float* dir, float* pdf, int32_t* flag)
{ GUARD( SampleOp<BsdfGeneric> op; auto parsed = bbmcu_host::parse_bsdf(bsdf); op.bsdf = make_desc(parsed); op.component = component; op.out = out; op.xi = xi; op.dir = dir; op.pdf = pdf; op.flag = flag; op.n = n; op.aligned = false; run(op, n); ) }

int hostsim_merl_index(const float* in, const float* out, size_t n, uint32_t* index)
{ GUARD( MerlIndexOp op; op.in = in; op.out = out; op.index = index; op.n = n; op.aligned = false; run(op, n); ) }

int hostsim_merl_dirs(uint32_t first, size_t n, float* in, float* out)
{ GUARD( MerlDirsOp op; op.first = first; op.in = in; op.out = out; op.n = n; op.aligned = false; run(op, n); ) }

int hostsim_spherical_dirs(const uint32_t* samples, const float* ranges, uint64_t first, size_t n, float* in, float* out)
{
  GUARD(
    SphericalDirsOp op; SphericalGrid& g = op.grid;
    g.n_in_phi = samples[0]; g.n_in_theta = samples[1]; g.n_out_phi = samples[2]; g.n_out_theta = samples[3];
    g.start_in_phi = ranges[0]; g.start_in_theta = ranges[1]; g.size_in_phi = ranges[2] - ranges[0]; g.size_in_theta = ranges[3] - ranges[1];
    g.start_out_phi = ranges[4]; g.start_out_theta = ranges[5]; g.size_out_phi = ranges[6] - ranges[4]; g.size_out_theta = ranges[7] - ranges[5];
    op.first = first; op.in = in; op.out = out; op.n = n; op.aligned = false; run(op, n);
  )
}

// merl_dirs_tab (the fused-linearizer path) against merl_dirs over bins [first, first + n): number of bins whose six floats differ in any bit
size_t hostsim_merl_dirs_tab_mismatches(uint32_t first, size_t n)
{
  std::vector<float> tab(kMerlLinTabFloats);
  for(int j=0; j < 360; ++j) merl_lin_tab_fill(tab.data(), j);
  size_t bad = 0;
  for(size_t i=0; i < n; ++i)
  {
    f3 a, b, c, d;
    merl_dirs(first + (uint32_t)i, a, b);
    merl_dirs_tab(tab.data(), first + (uint32_t)i, c, d);
    bad += (std::memcmp(&a, &c, sizeof(f3)) != 0) || (std::memcmp(&b, &d, sizeof(f3)) != 0);
  }
  return bad;
}
// the counter-based input generator of the kernels (generate_inputs): element i of (seed, first)
void hostsim_generate_inputs(uint64_t seed, uint64_t first, size_t n, float* out_xyz, float* xi_uv)
{
  for(size_t i=0; i < n; ++i)
  {
    GenInputs g = generate_inputs(seed, first + i);
    out_xyz[i] = g.out.x; out_xyz[n + i] = g.out.y; out_xyz[2*n + i] = g.out.z; xi_uv[i] = g.xi.x; xi_uv[n + i] = g.xi.y;
  }
}

// glibc-port checks: the restated atan2f / sinf / cosf against this host's libm
float hostsim_atan2f(float y, float x) { return glibc_atan2f(y, x); }
float hostsim_sinf(float x) { return glibc_sinf(x); }
float hostsim_cosf(float x) { return glibc_cosf(x); }
size_t hostsim_libm_mismatches(int which, const float* a, const float* b, size_t n)
{
  size_t bad = 0;
  for(size_t i=0; i < n; ++i)
  {
    float both_s, both_c; glibc_sincosf_both(a[i], both_s, both_c);     // which 3 / 4: the shared-reduction variant
    float mine = which == 0 ? glibc_atan2f(a[i], b[i]) : which == 1 ? glibc_sinf(a[i]) : which == 2 ? glibc_cosf(a[i]) : which == 3 ? both_s : which == 4 ? both_c
               : which == 5 ? glibc_expf(a[i]) : which == 6 ? glibc_logf(a[i]) : which == 7 ? glibc_erff(a[i])
               : which == 8 ? glibc_tanf(a[i]) : which == 9 ? glibc_erfcf(a[i]) : glibc_atanf(a[i]);
    float ref = which == 0 ? atan2f(a[i], b[i]) : (which == 1 || which == 3) ? sinf(a[i]) : (which == 2 || which == 4) ? cosf(a[i])
              : which == 5 ? expf(a[i]) : which == 6 ? logf(a[i]) : which == 7 ? erff(a[i])
              : which == 8 ? tanf(a[i]) : which == 9 ? erfcf(a[i]) : atanf(a[i]);
    if(mine != mine && ref != ref) continue;
    uint32_t u, v; std::memcpy(&u, &mine, 4); std::memcpy(&v, &ref, 4);
    bad += (u != v);
  }
  return bad;
}

// loss / gradient of ONE parameter set over explicit samples (directions + reference values), host loop
int hostsim_loss(const char* bsdf, int metric, int component, const float* in, const float* out, const float* ref, size_t n, double inv_n,
                 double* loss, double* grad, float* terms)
{
  GUARD(
    auto b = bbmcu_host::parse_bsdf(bsdf);
    BsdfDesc d = make_desc(b);
    int P = b.param_count(15);
    std::vector<double> acc(1 + P, 0.0);
    for(size_t i=0; i < n; ++i)
    {
      f3 a = make_f3(in[i], in[n+i], in[2*n+i]), o = make_f3(out[i], out[n+i], out[2*n+i]);
      Spec<float> r(ref[i], ref[n+i], ref[2*n+i]);
      float g[kMaxParams];
      float e = loss_sample_generic(d, metric, component, a, o, r, grad ? g : nullptr);
      if(terms) terms[i] = e;
      acc[0] += (double)e;
      if(grad) for(int j=0; j < P; ++j) acc[1+j] += (double)g[j];
    }
    *loss = acc[0] * inv_n;
    if(grad) for(int j=0; j < P; ++j) grad[j] = acc[1+j] * inv_n;
  )
}
} // extern "C"

// the compact pair loss (bbmcu_losscompact.cuh) of ONE parameter set over explicit samples: per-set constants, per-sample
// invariants and the accumulation exactly as k_loss_tile_compact runs them (8 samples per float accumulator, the
// below-horizon constant, column factors), summed in double.  Returns 2 if the BSDF has no compact kernel.
template<class CL, bool LOG> static void loss_compact_host(const BsdfDesc& d, int metric, const float* in, const float* out, const float* ref, size_t n, double inv_n, double* loss, double* grad)
{
  float set[CL::NSET];
  CL::set(d.attrs, set);
  std::vector<double> tot(CL::C, 0.0);
  constexpr int kCSPT = CL::kSPT;
  for(size_t i0=0; i0 < n; i0 += kCSPT)
  {
    float acc[CL::C]; float e_const = 0.0f;
    typename CL::Sample smp[kCSPT];
    for(int s=0; s < kCSPT; ++s)
    {
      const size_t i = i0 + s; const bool live = i < n; const size_t ii = live ? i : 0;
      int state;
      smp[s] = CL::make_geom(metric, make_f3(in[ii], in[n+ii], in[2*n+ii]), make_f3(out[ii], out[n+ii], out[2*n+ii]), live, state);
      const Spec<float> r(ref[ii], ref[n+ii], ref[2*n+ii]);
      if(state == 2) e_const += CL::below_const(metric, smp[s], r);
      CL::template set_ref<LOG>(smp[s], r, state == 0);
    }
    acc[0] = e_const; for(int j=1; j < CL::C; ++j) acc[j] = 0.0f;
    for(int s=0; s < kCSPT; ++s) { if(grad) CL::template accumulate<true, LOG>(set, smp[s], acc); else CL::template accumulate<false, LOG>(set, smp[s], acc); }
    for(int j=0; j < CL::C; ++j) tot[j] += (double)acc[j];
  }
  *loss = tot[0] * inv_n;
  if(grad) for(int j=0; j < CL::P; ++j) grad[j] = tot[1+j] * CL::col_scale(1+j) * inv_n;
}
extern "C" int hostsim_loss_compact(const char* bsdf, int metric, const float* in, const float* out, const float* ref, size_t n, double inv_n, double* loss, double* grad)
{
  try {
    auto b = bbmcu_host::parse_bsdf(bsdf);
    BsdfDesc d = make_desc(b);
    const bool pair = d.n_lobes == 2 && d.aggregate && d.model[0] == M_Lambertian, single = d.n_lobes == 1 && !d.aggregate;
    if(!pair && !single) return 2;
    bool done = false;
    dispatch_model(d.model[pair ? 1 : 0], [&](auto* tag) {
      using M = typename std::remove_pointer<decltype(tag)>::type;
      if constexpr (CompactOf<M>::value)
      {
        using CP = typename CompactOf<M>::type;
        using CS = typename CompactSingleOf<M>::type;
        const bool lg = metric > METRIC_BIERON_L2;
        if(pair) { if(lg) loss_compact_host<CP, true>(d, metric, in, out, ref, n, inv_n, loss, grad); else loss_compact_host<CP, false>(d, metric, in, out, ref, n, inv_n, loss, grad); }
        else     { if(lg) loss_compact_host<CS, true>(d, metric, in, out, ref, n, inv_n, loss, grad); else loss_compact_host<CS, false>(d, metric, in, out, ref, n, inv_n, loss, grad); }
        done = true;
      }
    });
    return done ? 0 : 2;
  } catch(const std::exception& e) { g_err = e.what(); return 1; }
}
