// Host side of the device checkBsdf (include/bbmcu.h, "checkBsdf on the device"): draws the random numbers of the
// reference's own stream when asked to, chunks the sample range over launches of k_check (bbmcu_check.cuh) and adds the
// block partial rows in fixed order.
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <vector>

#include "bbmcu_launch.cuh"
#include "bbmcu_check.cuh"

namespace bbmcu {
namespace {

struct DevBuf
{
  void* p = nullptr;
  explicit DevBuf(size_t bytes) { BBMCU_CUDA(cudaMalloc(&p, bytes ? bytes : 1)); }
  ~DevBuf() { if(p) cudaFree(p); }
  DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
  template<class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// the reference's generator (bin/checkBsdf.cpp:18-25): `Vec2d(U(rnd), U(rnd))` - g++ evaluates the second argument first,
// so the FIRST number of the stream becomes xi[1] (pinned by tests/test_check.py against the compiled reference tool)
struct RefStream
{
  std::mt19937 rnd;
  void vec2(float& x0, float& x1) { std::uniform_real_distribution<float> U(0, 1); x1 = U(rnd); x0 = U(rnd); }
};

// sampleSphere / sampleHemisphere / spherical::convert on the host, with the libm the reference itself calls
void host_to_vec(float phi, float theta, float* d) { float st = std::sin(theta), ct = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi); d[0] = cp*st; d[1] = sp*st; d[2] = ct; }
void host_sample_sphere(float x0, float x1, float* d) { double c = std::fmin(1.0, std::fmax(-1.0, 1.0 - 2.0*(double)x0)); host_to_vec(x1 * kTwoPi, (float)std::acos(c), d); }
void host_sample_hemisphere(float x0, float x1, float* d) { host_to_vec(x1 * kTwoPi, std::acos(std::fmin(1.0f, std::fmax(-1.0f, x0))), d); }

// regularised upper incomplete gamma function Q(a, x) (Numerical Recipes 6.2: series for x < a + 1, Lentz continued
// fraction otherwise) in double; the reference evaluates bbm::gamma_q in float (util/gamma.h:563-587)
double host_gamma_q(double a, double x)
{
  if(!(x >= 0.0) || !(a > 0.0)) return std::numeric_limits<double>::quiet_NaN();
  if(x == 0.0) return 1.0;
  const double lg = std::lgamma(a);
  if(x < a + 1.0)
  {
    double ap = a, del = 1.0/a, sum = del;
    for(int n=0; n < 100000; ++n) { ap += 1.0; del *= x/ap; sum += del; if(std::fabs(del) < std::fabs(sum)*1e-16) break; }
    return 1.0 - sum * std::exp(-x + a*std::log(x) - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0/tiny, d = 1.0/b, h = d;
  for(int i=1; i < 100000; ++i)
  {
    double an = -i*(i - a);
    b += 2.0;
    d = an*d + b; if(std::fabs(d) < tiny) d = tiny;
    c = b + an/c; if(std::fabs(c) < tiny) c = tiny;
    d = 1.0/d;
    double del = d*c; h *= del;
    if(std::fabs(del - 1.0) < 1e-16) break;
  }
  return std::exp(-x + a*std::log(x) - lg) * h;
}

constexpr uint64_t kChunkItems = 1ull << 24;        // items per launch (the index of the maximum is kept in 32 bits)

struct Runner
{
  bbmcu_ctx* ctx; BsdfDesc desc; int rng; uint64_t seed; uint64_t next_counter = 0; RefStream ref;
  int max_blocks;
  std::unique_ptr<DevBuf> d_partial, d_key, d_xi;
  std::vector<double> h_partial; std::vector<unsigned long long> h_key; std::vector<float> h_xi;
  Runner(bbmcu_ctx* c, const bbmcu_bsdf* b, int rng_, uint64_t seed_) : ctx(c), rng(rng_), seed(seed_)
  {
    if(!ctx) throw std::invalid_argument("BBM: null context");
    if(!b) throw std::invalid_argument("BBM: null bsdf");
    if(rng != BBMCU_RNG_PHILOX && rng != BBMCU_RNG_MT19937) throw std::invalid_argument("BBM: unknown random number source");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    desc = make_desc(b->b, ctx->device);
    max_blocks = ctx->sm_count * 8;
    d_partial.reset(new DevBuf(sizeof(double) * (size_t)max_blocks * kCheckCols));
    d_key.reset(new DevBuf(sizeof(unsigned long long) * (size_t)max_blocks));
    h_partial.resize((size_t)max_blocks * kCheckCols); h_key.resize(max_blocks);
  }
  // a direction of the trial loop (drawn before the trial's samples, as the reference does)
  void trial_direction(bool sphere, float* d)
  {
    float x0, x1;
    if(rng == BBMCU_RNG_MT19937) ref.vec2(x0, x1);
    else { uint32_t w[4]; philox4x32_10(0xFFFFFFFF00000000ull + next_counter++, seed ^ 0x9E3779B97F4A7C15ull, w); x0 = u01(w[0]); x1 = u01(w[1]); }
    if(sphere) host_sample_sphere(x0, x1, d); else host_sample_hemisphere(x0, x1, d);
  }
  void fill_stream(CheckArgs& a, uint64_t items, int per_item)
  {
    a.seed = seed; a.first = next_counter; a.xi = nullptr; a.xi_per_item = per_item;
    if(rng == BBMCU_RNG_MT19937)
    {
      h_xi.resize(items * (size_t)per_item);
      for(uint64_t i=0; i < items; ++i) for(int j=0; j < per_item; j += 2) ref.vec2(h_xi[i*per_item + j], h_xi[i*per_item + j + 1]);
      d_xi.reset(new DevBuf(h_xi.size() * sizeof(float)));
      BBMCU_CUDA(cudaMemcpyAsync(d_xi->p, h_xi.data(), h_xi.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      a.xi = d_xi->as<float>();
    }
    else next_counter += items;
  }
  void launch(int kernel, const CheckArgs& a, dim3 grid)
  {
    switch(kernel) {
      case CHECK_KERNEL_E:  launch_check_e(ctx, ctx->stream, desc, a, grid); break;
      case CHECK_KERNEL_SE: launch_check_se(ctx, ctx->stream, desc, a, grid); break;
      case CHECK_KERNEL_S:  launch_check_s(ctx, ctx->stream, desc, a, grid); break;
      default:              launch_check_p(ctx, ctx->stream, desc, a, grid); break;
    }
  }
  // `total` items of one test in chunks; sums[kCheckCols] accumulate; on_chunk(base, items, args) after every chunk (synchronised)
  template<class Extra> void run(int kernel, CheckArgs a, uint64_t total, int per_item, double* sums, Extra&& on_chunk)
  {
    for(uint64_t base = 0; base < total; base += kChunkItems)
    {
      const uint64_t items = std::min<uint64_t>(kChunkItems, total - base);
      fill_stream(a, items, per_item);
      a.n = items; a.partial = d_partial->as<double>(); a.maxkey = d_key->as<unsigned long long>();
      unsigned blocks = (unsigned)std::min<uint64_t>((items + kCheckThreads - 1) / kCheckThreads, (uint64_t)max_blocks);
      launch(kernel, a, dim3(blocks, 1, 1));
      BBMCU_CUDA(cudaMemcpyAsync(h_partial.data(), d_partial->p, sizeof(double) * (size_t)blocks * kCheckCols, cudaMemcpyDeviceToHost, ctx->stream));
      BBMCU_CUDA(cudaMemcpyAsync(h_key.data(), d_key->p, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost, ctx->stream));
      BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
      for(unsigned b=0; b < blocks; ++b) for(int j=0; j < kCheckCols; ++j) sums[j] += h_partial[(size_t)b*kCheckCols + j];
      on_chunk(base, items, a, blocks);
    }
  }
};

CheckArgs base_args(int mode)
{
  CheckArgs a; std::memset(&a, 0, sizeof(a));
  a.mode = mode; a.component = BBMCU_ALL;
  return a;
}

} // anonymous namespace
} // namespace bbmcu

using namespace bbmcu;

extern "C" {

int bbmcu_check_reflectance(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int n_theta, int importance, int rng, uint64_t seed,
                            double* estimate, float* reflectance, float* out_dirs)
{
  return guarded(ctx, [&] {
    if(n_theta < 1 || !estimate) throw std::invalid_argument("BBM: invalid argument");
    Runner R(ctx, bsdf, rng, seed);
    DevBuf d_dir(3*sizeof(float)), d_rgb(3*sizeof(float));
    for(int t=0; t < n_theta; ++t)
    {
      // spherical::theta(out_sp) = theta_idx * Pi(0.5) / numtheta; phi = 0 (bin/checkBsdf.cpp:70-75)
      float theta = (float)t * (float)(0.5f * 3.14159265358979323846) / (float)n_theta, out[3];
      host_to_vec(0.0f, theta, out);
      CheckArgs a = base_args(CHECK_REFLECTANCE);
      a.ox = out[0]; a.oy = out[1]; a.oz = out[2]; a.importance = importance ? 1 : 0;
      double sums[kCheckCols] = {};
      R.run(importance ? CHECK_KERNEL_SE : CHECK_KERNEL_E, a, samples, 2, sums, [](uint64_t, uint64_t, const CheckArgs&, unsigned) {});
      for(int c=0; c < 3; ++c) estimate[3*t + c] = samples ? sums[c] / (double)samples : 0.0;
      if(out_dirs) std::memcpy(out_dirs + 3*t, out, sizeof(out));
      if(reflectance)
      {
        BBMCU_CUDA(cudaMemcpyAsync(d_dir.p, out, sizeof(out), cudaMemcpyHostToDevice, ctx->stream));
        const size_t saved = ctx->ld; ctx->ld = 0;
        launch_reflectance(ctx, ctx->stream, R.desc, BBMCU_ALL, d_dir.as<float>(), d_rgb.as<float>(), 1);
        ctx->ld = saved;
        BBMCU_CUDA(cudaMemcpyAsync(reflectance + 3*t, d_rgb.p, 3*sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
      }
    }
  });
}

int bbmcu_check_reciprocity(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int rng, uint64_t seed, double* mean, float* max_diff, float* max_pair)
{
  return guarded(ctx, [&] {
    if(!mean) throw std::invalid_argument("BBM: invalid argument");
    Runner R(ctx, bsdf, rng, seed);
    CheckArgs a = base_args(CHECK_RECIPROCITY);
    double sums[kCheckCols] = {};
    uint32_t best_h = 0; bool have = false; uint64_t best_counter = 0; float best_xi[4] = {};
    R.run(CHECK_KERNEL_E, a, samples, 4, sums, [&](uint64_t, uint64_t, const CheckArgs& ca, unsigned blocks) {
      unsigned long long k = 0;
      for(unsigned b=0; b < blocks; ++b) if(R.h_key[b] > k) k = R.h_key[b];
      const uint32_t h = (uint32_t)(k >> 32), local = 0xFFFFFFFFu - (uint32_t)k;
      if(k && h > best_h)                    // strict: an earlier chunk keeps a tie, as the reference's earlier sample does
      {
        best_h = h; have = true; best_counter = ca.first + local;
        if(ca.xi) std::memcpy(best_xi, R.h_xi.data() + (size_t)local*4, sizeof(best_xi));
      }
    });
    for(int c=0; c < 3; ++c) mean[c] = samples ? sums[c] / (double)samples : 0.0;
    float detail[9] = {};
    if(have)
    {
      // evaluate that one pair again and read its directions and differences
      DevBuf d_detail(sizeof(detail)), d_one(4*sizeof(float));
      CheckArgs one = base_args(CHECK_RECIPROCITY);
      one.seed = seed; one.first = best_counter; one.n = 1; one.xi_per_item = 4;
      if(rng == BBMCU_RNG_MT19937) { BBMCU_CUDA(cudaMemcpyAsync(d_one.p, best_xi, sizeof(best_xi), cudaMemcpyHostToDevice, ctx->stream)); one.xi = d_one.as<float>(); }
      one.partial = R.d_partial->as<double>(); one.maxkey = R.d_key->as<unsigned long long>(); one.detail = d_detail.as<float>();
      R.launch(CHECK_KERNEL_E, one, dim3(1, 1, 1));
      BBMCU_CUDA(cudaMemcpyAsync(detail, d_detail.p, sizeof(detail), cudaMemcpyDeviceToHost, ctx->stream));
      BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if(max_pair) std::memcpy(max_pair, detail, 6*sizeof(float));
    if(max_diff) std::memcpy(max_diff, detail + 6, 3*sizeof(float));
  });
}

int bbmcu_check_pdf(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int sample_sphere, int check_below_horizon, int rng, uint64_t seed,
                    uint64_t* counts, double* mismatch, float* offenders, int max_offenders, int* n_offenders)
{
  return guarded(ctx, [&] {
    if(!counts || !mismatch) throw std::invalid_argument("BBM: invalid argument");
    Runner R(ctx, bsdf, rng, seed);
    if(max_offenders < 0 || !offenders) max_offenders = 0;
    DevBuf d_off(sizeof(float) * kCheckOffenderFloats * (size_t)std::max(1, max_offenders)), d_noff(sizeof(unsigned int));
    BBMCU_CUDA(cudaMemsetAsync(d_noff.p, 0, sizeof(unsigned int), ctx->stream));
    CheckArgs a = base_args(CHECK_PDF);
    a.sphere = sample_sphere ? 1 : 0; a.below_horizon = check_below_horizon ? 1 : 0;
    if(max_offenders) { a.offenders = d_off.as<float>(); a.n_offenders = d_noff.as<unsigned int>(); a.max_offenders = max_offenders; }
    double sums[kCheckCols] = {};
    R.run(CHECK_KERNEL_S, a, samples, 6, sums, [](uint64_t, uint64_t, const CheckArgs&, unsigned) {});
    mismatch[0] = samples ? sums[0] / (double)samples : 0.0; mismatch[1] = samples ? sums[1] / (double)samples : 0.0;
    for(int j=0; j < 4; ++j) counts[j] = (uint64_t)std::llround(sums[2 + j]);
    unsigned int n = 0;
    if(max_offenders)
    {
      BBMCU_CUDA(cudaMemcpyAsync(&n, d_noff.p, sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
      BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
      if(n > (unsigned int)max_offenders) n = (unsigned int)max_offenders;
      BBMCU_CUDA(cudaMemcpy(offenders, d_off.p, sizeof(float) * kCheckOffenderFloats * (size_t)n, cudaMemcpyDeviceToHost));
    }
    if(n_offenders) *n_offenders = (int)n;
  });
}

int bbmcu_check_pdf_integral(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int trials, int sample_sphere, int rng, uint64_t seed,
                             double* integral, float* dirs)
{
  return guarded(ctx, [&] {
    if(trials < 0 || !integral) throw std::invalid_argument("BBM: invalid argument");
    Runner R(ctx, bsdf, rng, seed);
    for(int t=0; t < trials; ++t)
    {
      float d[3];
      R.trial_direction(sample_sphere != 0, d);
      CheckArgs a = base_args(CHECK_PDF_INTEGRAL);
      a.ox = d[0]; a.oy = d[1]; a.oz = d[2];
      double sums[kCheckCols] = {};
      R.run(CHECK_KERNEL_P, a, samples, 2, sums, [](uint64_t, uint64_t, const CheckArgs&, unsigned) {});
      integral[t] = samples ? sums[0] / (double)samples : 0.0;
      if(dirs) std::memcpy(dirs + 3*t, d, sizeof(d));
    }
  });
}

int bbmcu_check_sample(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t pdf_samples, uint64_t samples, int n_theta, int n_phi, int trials,
                       int sample_sphere, int include_zero, int rng, uint64_t seed, double* chi2, double* df, double* P, float* dirs,
                       double* bin_pdf, uint64_t* bin_count)
{
  return guarded(ctx, [&] {
    if(trials < 0 || n_theta < 1 || n_phi < 1 || !chi2 || !df) throw std::invalid_argument("BBM: invalid argument");
    const uint64_t bins = (uint64_t)n_theta * (uint64_t)n_phi;
    if(bins > (1u << 20) || pdf_samples > kChunkItems) throw std::invalid_argument("BBM: too many bins or pdf samples per bin");
    Runner R(ctx, bsdf, rng, seed);
    const unsigned bx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((pdf_samples + kCheckThreads - 1) / kCheckThreads, std::max<uint64_t>(1, (uint64_t)R.max_blocks / bins)));
    DevBuf d_binpart(sizeof(double) * bins * bx), d_counts(sizeof(unsigned int) * bins);
    std::vector<double> h_binpart(bins * bx), pdf(bins); std::vector<unsigned int> h_counts(bins); std::vector<uint64_t> count(bins);
    for(int t=0; t < trials; ++t)
    {
      float d[3];
      R.trial_direction(sample_sphere != 0, d);
      // (1) the pdf integrated over every bin: one launch, block row = bin
      CheckArgs a = base_args(CHECK_BIN_PDF);
      a.ox = d[0]; a.oy = d[1]; a.oz = d[2]; a.n_theta = n_theta; a.n_phi = n_phi;
      R.fill_stream(a, bins * pdf_samples, 2);
      a.n = pdf_samples; a.partial = d_binpart.as<double>();
      R.launch(CHECK_KERNEL_P, a, dim3(bx, (unsigned)bins, 1));
      BBMCU_CUDA(cudaMemcpyAsync(h_binpart.data(), d_binpart.p, sizeof(double) * bins * bx, cudaMemcpyDeviceToHost, ctx->stream));
      BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
      for(uint64_t b=0; b < bins; ++b) { double s = 0.0; for(unsigned k=0; k < bx; ++k) s += h_binpart[b*bx + k]; pdf[b] = pdf_samples ? s / (double)pdf_samples : 0.0; }
      // (2) histogram of sampled directions
      BBMCU_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(unsigned int) * bins, ctx->stream));
      std::fill(count.begin(), count.end(), 0);
      CheckArgs c = base_args(CHECK_BIN_COUNT);
      c.ox = d[0]; c.oy = d[1]; c.oz = d[2]; c.n_theta = n_theta; c.n_phi = n_phi; c.include_zero = include_zero ? 1 : 0; c.counts = d_counts.as<unsigned int>();
      double sums[kCheckCols] = {};
      R.run(CHECK_KERNEL_S, c, samples, 2, sums, [&](uint64_t, uint64_t, const CheckArgs&, unsigned) {
        // 32-bit device counters are drained after every chunk (2^24 items)
        BBMCU_CUDA(cudaMemcpyAsync(h_counts.data(), d_counts.p, sizeof(unsigned int) * bins, cudaMemcpyDeviceToHost, ctx->stream));
        BBMCU_CUDA(cudaMemsetAsync(d_counts.p, 0, sizeof(unsigned int) * bins, ctx->stream));
        BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
        for(uint64_t b=0; b < bins; ++b) count[b] += h_counts[b];
      });
      // (3) chi-square as the reference forms it (bin/checkBsdf.cpp:391-404): bins with m > eps and more than 5 samples
      double x2 = 0.0, dof = -1.0;
      for(uint64_t b=0; b < bins; ++b)
      {
        const double m = pdf[b] * (double)samples;
        if(m > (double)kEps && count[b] > 5) { x2 += ((double)count[b] - m)*((double)count[b] - m) / m; dof += 1.0; }
      }
      chi2[t] = x2; df[t] = dof;
      if(P) P[t] = dof > 1.0 ? host_gamma_q((dof - 1.0) / 2.0, x2 / 2.0) : std::numeric_limits<double>::quiet_NaN();
      if(dirs) std::memcpy(dirs + 3*t, d, sizeof(d));
      if(bin_pdf) std::memcpy(bin_pdf + (size_t)t*bins, pdf.data(), sizeof(double)*bins);
      if(bin_count) std::memcpy(bin_count + (size_t)t*bins, count.data(), sizeof(uint64_t)*bins);
    }
  });
}

} // extern "C"
