// The boundary, compiled against the REFERENCE's own headers (-I/root/reference/include, native backbone types through
// backbone/cuda/include/backbone.h): the fitting example of docs/source/fitting.rst:16-60 with the loss class replaced by
// bbm::cuda::loss - the model, parameter_values / bounds (include/bbm/bsdf_enumerate.h:102-237) and bbm::compass
// (include/optimizer/compass.h:39-185) are the reference's, UNMODIFIED.  The same search then runs on the reference's own
// standardLog loss on the CPU; both traces are printed as JSON for tests/test_reference_boundary.py.
// Built HERE (where /root/reference exists) by __graft_entry__.build() into tests/_build/; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <iostream>
#include "bbm.h"
using namespace bbm;

#include "bbm/bsdf_import.h"
#include "optimizer/compass.h"
#include "loss/cosine_weighted_log.h"
#include "loss/cosine_weighted_l2.h"
#include "bbm_cuda/loss.h"

// ---- what the boundary promises, checked by the compiler against the reference's concepts --------------------------------
static_assert(concepts::config<floatRGB_cuda>, "floatRGB_cuda is a bbm configuration");
using fitted_t = decltype(aggregate(lambertian<floatRGB_cuda>(), cooktorrance<floatRGB_cuda>()));
using cuda_loss_t = bbm::cuda::loss<fitted_t>;
static_assert(concepts::has_config<cuda_loss_t>);
static_assert(concepts::lossfunction<cuda_loss_t>, "concepts::lossfunction (include/concepts/lossfunction.h:28-35)");
static_assert(concepts::sampledlossfunction<cuda_loss_t>, "concepts::sampledlossfunction (include/concepts/sampledlossfunction.h:26-36)");
using param_t = decltype(parameter_values(std::declval<fitted_t&>()));
using box_t = decltype(parameter_lower_bound(std::declval<fitted_t&>()));
static_assert(concepts::optimization_algorithm<compass<cuda_loss_t, param_t, box_t>>, "bbm::compass on the CUDA loss (include/concepts/optimization_algorithm.h:24-37)");

template<typename P> static void print_params(const char* name, const P& param)
{
  std::printf("\"%s\": [", name);
  size_t j = 0;
  for(const auto& p : param) std::printf("%s%.9g", j++ ? ", " : "", double(float(p)));
  std::printf("]");
}

int main(int argc, char** argv)
{
  BBM_IMPORT_CONFIG( floatRGB_cuda );
  const size_t maxItr = argc > 1 ? std::atoi(argv[1]) : 30;
  try
  {
    const std::string truth_s = "Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))";
    auto reference = bsdf_import<Config>(truth_s);

    // ---- the documented example on the CUDA backbone ------------------------------------------------------------------
    auto fitted = aggregate( lambertian<Config>(), cooktorrance<Config>() );
    ::bbmcu::context ctx(0);
    bbm::cuda::loss loss(ctx, bbm::cuda::metric::standardLog, fitted, reference, vec2d<Size_t>{90, 30}, vec2d<Size_t>{1, 9});
    auto param = parameter_values(fitted);
    auto low = parameter_lower_bound(fitted);
    auto up = parameter_upper_bound(fitted);
    std::printf("{\n\"samples\": %zu,\n\"initial_loss\": %.9g,\n", size_t(loss.samples()), double(loss()));
    std::printf("\"term_12345\": %.9g,\n", double(loss(Size_t(12345))));
    std::vector<Value> grad;
    Value lg = loss.gradient(grad);
    std::printf("\"gradient_loss\": %.9g, ", double(lg)); print_params("gradient", grad); std::printf(",\n");

    compass opt(loss, param, low, up);
    std::printf("\"cuda_trace\": [");
    auto t0 = std::chrono::steady_clock::now();
    size_t t = 0;
    for(; t < maxItr && !bbm::all(opt.is_converged()); ++t) std::printf("%s%.9g", t ? ", " : "", double(opt.step()));
    double cuda_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("],\n\"cuda_steps\": %zu, \"cuda_seconds\": %.6f, ", t, cuda_s); print_params("cuda_params", param);
    std::printf(",\n\"cuda_fitted\": \"%s\",\n", bbm::toString(fitted).c_str());

    // ---- the same search on the reference's own CPU loss --------------------------------------------------------------------
    auto fitted_cpu = aggregate( lambertian<Config>(), cooktorrance<Config>() );
    standardLog loss_cpu(fitted_cpu, reference, {90, 30}, {1, 9});
    auto param_cpu = parameter_values(fitted_cpu);
    std::printf("\"cpu_initial_loss\": %.9g, \"cpu_term_12345\": %.9g,\n", double(loss_cpu()), double(loss_cpu(Size_t(12345))));
    compass opt_cpu(loss_cpu, param_cpu, low, up);
    std::printf("\"cpu_trace\": [");
    t0 = std::chrono::steady_clock::now();
    for(t = 0; t < maxItr && !bbm::all(opt_cpu.is_converged()); ++t) std::printf("%s%.9g", t ? ", " : "", double(opt_cpu.step()));
    double cpu_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("],\n\"cpu_steps\": %zu, \"cpu_seconds\": %.6f, ", t, cpu_s); print_params("cpu_params", param_cpu);
    std::printf("\n}\n");
  }
  catch(const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
