// Exercises the C++ adapter headers (include/bbmcu/*.hpp) the way docs/source/fitting.rst:16-60 of the reference uses
// bbm: import two BSDFs, build a loss over a spherical grid, run compass search, print what happened as JSON.
// Built and run by tests/test_cpp_adapters.py on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <string>
#include "bbmcu/bsdf.hpp"
#include "bbmcu/loss.hpp"
#include "bbmcu/optimizer.hpp"
#include "bbmcu/fit.hpp"
#include <cuda_runtime_api.h>

using namespace bbmcu;

static void print_vec(const char* name, const std::vector<double>& v, bool comma = true)
{
  std::printf("\"%s\": [", name);
  for(size_t i = 0; i < v.size(); ++i) std::printf("%s%.9g", i ? ", " : "", v[i]);
  std::printf("]%s\n", comma ? "," : "");
}

int main(int argc, char** argv)
{
  const int steps = argc > 1 ? std::atoi(argv[1]) : 12;
  try
  {
    context ctx(0);
    const std::string truth_s = "Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))";
    cuda_bsdf truth = bsdf_import(ctx, truth_s);
    cuda_bsdf fitted = bsdf_import(ctx, "Aggregate(Lambertian(), CookTorrance())");
    std::printf("{\n\"toString\": \"%s\",\n", truth.toString().c_str());

    // scalar concept calls
    vec3 in{0.3f, 0.2f, 0.9327379f}, out{0.5f, -0.1f, 0.8602325f};
    spectrum e = truth.eval(in, out);
    float p = truth.pdf(in, out);
    bsdf_sample s = truth.sample(out, vec2{0.3f, 0.6f});
    spectrum r = truth.reflectance(out);
    std::printf("\"eval\": [%.9g, %.9g, %.9g], \"pdf\": %.9g, \"sample\": [%.9g, %.9g, %.9g, %.9g, %d], \"reflectance\": [%.9g, %.9g, %.9g],\n",
                e[0], e[1], e[2], p, s.direction[0], s.direction[1], s.direction[2], s.pdf, (int)s.flag, r[0], r[1], r[2]);

    // loss over the (13,8) x (5,6) spherical grid of tests/golden/losses.json
    std::vector<double> params = fitted.parameter_values();
    cuda_loss loss(metric::nganL2, fitted, truth, params, spherical_grid({13, 8}, {5, 6}));
    std::vector<double> grad;
    double l0 = loss.gradient(grad);
    std::printf("\"samples\": %zu, \"loss0\": %.12g,\n", loss.samples(), l0);
    print_vec("gradient", grad);

    // the reference's compass search, probe by probe.  A single model on purpose: with a run-time Aggregate the
    // reference checks per-lobe REVERSED parameters against forward-order bounds, every probe
    // falls "outside the box" and its compass never moves.
    cuda_bsdf truth1 = bsdf_import(ctx, "CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5)");
    cuda_bsdf fitted1 = bsdf_import(ctx, "CookTorrance()");
    std::vector<double> params1 = fitted1.parameter_values();
    cuda_loss loss1(metric::nganL2, fitted1, truth1, params1, spherical_grid({13, 8}, {5, 6}));
    std::vector<double> lo = fitted1.parameter_lower_bound(), hi = fitted1.parameter_upper_bound();
    compass<cuda_loss> opt(loss1, params1, lo, hi);
    std::vector<double> trace;
    for(int i = 0; i < steps && !opt.is_converged(); ++i) trace.push_back(opt.step());
    print_vec("compass_trace", trace);
    print_vec("compass_params", params1);

    // all 2P probes per launch
    std::vector<double> params_b = fitted1.parameter_values();
    cuda_loss loss_b(metric::nganL2, fitted1, truth1, params_b, spherical_grid({13, 8}, {5, 6}));
    compass_batched opt_b(loss_b, lo, hi);
    std::vector<double> trace_b;
    uint64_t l_before = ctx.launches();
    for(int i = 0; i < steps && !opt_b.is_converged(); ++i) trace_b.push_back(opt_b.step());
    std::printf("\"batched_launches\": %llu,\n", (unsigned long long)(ctx.launches() - l_before));
    print_vec("batched_trace", trace_b);

    // gradient optimiser (on the aggregate)
    lo = fitted.parameter_lower_bound(); hi = fitted.parameter_upper_bound();
    std::vector<double> params_g = fitted.parameter_values();
    cuda_loss loss_g(metric::nganL2, fitted, truth, params_g, spherical_grid({13, 8}, {5, 6}));
    gradient_descent gd(loss_g, lo, hi, 3e-2);
    std::vector<double> trace_g;
    for(int i = 0; i < 200 && !gd.is_converged(); ++i) trace_g.push_back(gd.step());
    std::printf("\"gd_first\": %.9g, \"gd_last\": %.9g,\n", trace_g.front(), trace_g.back());
    loss_g.commit();
    std::printf("\"gd_fitted\": \"%s\",\n", fitted.toString().c_str());

    // two shards of one loss on two contexts (streams) of this device, combined over each other's exchange windows
    {
      context ctx2(0);
      cuda_bsdf fitted_b = bsdf_import(ctx2, fitted.toString());
      std::vector<double> pa = fitted.parameter_values(), pb = pa;
      const auto grid = spherical_grid({13, 8}, {5, 6});
      cuda_loss whole(metric::nganL2, fitted, truth, pa, grid);
      const uint64_t n = whole.samples(), cut = n / 2 + 7;
      cuda_loss sa(metric::nganL2, fitted, truth, pa, grid, bsdf_flag::All, 0, cut);
      cuda_loss sb(metric::nganL2, fitted_b, truth, pb, grid, bsdf_flag::All, cut, n - cut);
      const size_t cols = 1 + pa.size();
      void *wa = nullptr, *wb = nullptr;
      sa.peer_init(0, 2, 4*cols, &wa);
      sb.peer_init(1, 2, 4*cols, &wb);
      sa.peer_connect(std::vector<void*>{wa, wb});
      sb.peer_connect(std::vector<void*>{wa, wb});
      double *da = nullptr, *db = nullptr;
      if(cudaMalloc((void**)&da, cols*sizeof(double)) != cudaSuccess || cudaMalloc((void**)&db, cols*sizeof(double)) != cudaSuccess) throw std::runtime_error("cudaMalloc");
      sa.eval_device(pa, da);                      // asynchronous: its gather kernel waits on the device for the other shard's rows
      sb.eval_device(pb, db);
      ctx.synchronize(); ctx2.synchronize();
      std::vector<double> ha(cols), hb(cols);
      cudaMemcpy(ha.data(), da, cols*sizeof(double), cudaMemcpyDeviceToHost);
      cudaMemcpy(hb.data(), db, cols*sizeof(double), cudaMemcpyDeviceToHost);
      cudaFree(da); cudaFree(db);
      std::vector<double> g;
      const double l = whole.gradient(g);
      std::printf("\"peer_identical\": %s, \"peer_loss\": %.12g, \"whole_loss\": %.12g, \"peer_grad0\": %.12g, \"whole_grad0\": %.12g,\n",
                  ha == hb ? "true" : "false", ha[0], l, ha[1], g[0]);
    }

    // error behaviour: the reference throws std::invalid_argument from its parser
    bool threw = false;
    try { bsdf_import(ctx, "NoSuchModel(1, 2)"); } catch(const std::invalid_argument&) { threw = true; }
    std::printf("\"invalid_argument\": %s\n}\n", threw ? "true" : "false");
    return 0;
  }
  catch(const std::exception& ex) { std::fprintf(stderr, "FAILED: %s\n", ex.what()); return 1; }
}
