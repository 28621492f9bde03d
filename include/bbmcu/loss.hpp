// bbmcu/loss.hpp - the reference's loss functions on the CUDA backbone.
//
// cuda_loss has the METHODS of concepts::lossfunction (include/concepts/lossfunction.h:28-35: update() and
// Value operator()(Mask = true) const) and samples(), but it is a stand-alone class over the C ABI that knows nothing of
// bbm's Config types, so it does not itself satisfy the concept (which requires a Config typedef,
// include/concepts/config.h:53).  The class that does - and that the UNMODIFIED bbm::compass (include/optimizer/compass.h:39-185)
// drives - is bbm::cuda::loss in backbone/cuda/include/bbm_cuda/loss.h, a thin shell over this one compiled against the
// reference's headers (tests/cpp/test_reference_boundary.cpp checks the concepts with static_assert).  The optimizers of
// optimizer.hpp run against this class directly.  New on this backbone: evaluation of K parameter vectors in ONE launch
// and the analytic gradient.
//
// The six metrics are the error functors of include/loss/cosine_weighted_l2.h:25-34,96-105,166-176 and
// include/loss/cosine_weighted_log.h:32-43,101-112,170-181 (nganL2, lowL2, bieronL2, lowLog, bieronLog, standardLog).
#pragma once
#include <array>
#include <algorithm>
#include <optional>
#include "bsdf.hpp"

namespace bbmcu {

enum class metric : int { nganL2 = BBMCU_NGAN_L2, lowL2 = BBMCU_LOW_L2, bieronL2 = BBMCU_BIERON_L2,
                          lowLog = BBMCU_LOW_LOG, bieronLog = BBMCU_BIERON_LOG, standardLog = BBMCU_STANDARD_LOG };

// spherical_linearizer constructor arguments (include/linearizer/spherical_linearizer.h:37-44); samples = (phi, theta)
inline bbmcu_spherical_grid spherical_grid(std::array<uint32_t,2> samples_in, std::array<uint32_t,2> samples_out)
{ bbmcu_spherical_grid g; bbmcu_spherical_grid_default(&g, samples_in[0], samples_in[1], samples_out[0], samples_out[1]); return g; }

class cuda_loss
{
public:
  // `params` is the live parameter vector the optimizer mutates (the reference binds bbm::vector<Value&> into the
  // model, include/bbm/bsdf_enumerate.h:102-120; here the binding is explicit).  grid == nullopt: merl_linearizer.
  // [first, first + count) is the shard of the sample axis this process owns (count == 0: everything).
  cuda_loss(metric m, cuda_bsdf& fitted, const cuda_bsdf& reference, std::vector<double>& params,
            std::optional<bbmcu_spherical_grid> grid = std::nullopt, bsdf_flag component = bsdf_flag::All, uint64_t first = 0, uint64_t count = 0)
    : _fitted(fitted), _params(&params)
  { create((int)m, grid, component, reference.get(), nullptr, first, count); }
  // measured reference: a MERL table of 3 x BBMCU_MERL_BINS floats (host or device), looked up as merl_data::eval does
  cuda_loss(metric m, cuda_bsdf& fitted, const float* merl_rgb, std::vector<double>& params,
            std::optional<bbmcu_spherical_grid> grid = std::nullopt, bsdf_flag component = bsdf_flag::All, uint64_t first = 0, uint64_t count = 0)
    : _fitted(fitted), _params(&params)
  { create((int)m, grid, component, nullptr, merl_rgb, first, count); }

  // ---- concepts::lossfunction ------------------------------------------------------------------------------------
  void update() {}                                   // the reference tabulation is fixed at construction
  float operator()(bool mask = true) const
  {
    if(!mask) return 0.0f;
    double l = 0.0;
    check(bbmcu_loss_eval(_loss.get(), _fitted.get(), _params->data(), 1, &l, nullptr, nullptr), _fitted.ctx().get());
    return (float)l;
  }
  size_t samples() const { return (size_t)bbmcu_loss_samples(_loss.get()); }
  bbmcu_loss* handle() const { return _loss.get(); }

  // ---- batched / gradient (new) -------------------------------------------------------------------------------------
  size_t parameters() const { return _params->size(); }
  std::vector<double>& live_parameters() const { return *_params; }
  // loss at K parameter vectors (row-major K x P) in one launch
  std::vector<double> operator()(const std::vector<double>& params_KxP) const
  {
    const size_t P = parameters(), K = P ? params_KxP.size() / P : 0;
    std::vector<double> l(K);
    if(K) check(bbmcu_loss_eval(_loss.get(), _fitted.get(), params_KxP.data(), K, l.data(), nullptr, nullptr), _fitted.ctx().get());
    return l;
  }
  // loss and d loss / d parameter at the live parameters
  double gradient(std::vector<double>& grad) const
  {
    double l = 0.0;
    grad.resize(parameters());
    check(bbmcu_loss_eval(_loss.get(), _fitted.get(), _params->data(), 1, &l, grad.data(), nullptr), _fitted.ctx().get());
    return l;
  }
  // K x (1 + P) doubles [loss_k, grad_k...] left in DEVICE memory on the context's stream: what a rank hands to ncclAllReduce
  void eval_device(const std::vector<double>& params_KxP, double* device_out) const
  {
    const size_t P = parameters(), K = P ? params_KxP.size() / P : 0;
    if(K) check(bbmcu_loss_eval(_loss.get(), _fitted.get(), params_KxP.data(), K, nullptr, nullptr, device_out), _fitted.ctx().get());
  }
  // ---- shards of one loss on the GPUs of a node: combine inside the library's kernels over NVLink peer memory ------------
  // (bbmcu_loss_peer_*; after connecting, every evaluation is a collective over the shards and returns their sum)
  // allocate this shard's exchange window for batches of up to max_values = K*(1+P) doubles: the 64-byte handle goes to
  // the other processes (MPI_Allgather, a pipe, ...), *window to contexts of this process
  std::array<unsigned char, 64> peer_init(int rank, int world, size_t max_values, void** window = nullptr)
  {
    std::array<unsigned char, 64> h{};
    check(bbmcu_loss_peer_init(_loss.get(), rank, world, max_values, h.data(), window), _fitted.ctx().get());
    return h;
  }
  void peer_connect(const std::vector<std::array<unsigned char, 64>>& handles_by_rank)
  {
    std::vector<unsigned char> blob(64*handles_by_rank.size());
    for(size_t r=0; r < handles_by_rank.size(); ++r) std::copy(handles_by_rank[r].begin(), handles_by_rank[r].end(), blob.begin() + 64*r);
    check(bbmcu_loss_peer_connect(_loss.get(), blob.data()), _fitted.ctx().get());
  }
  void peer_connect(const std::vector<void*>& windows_by_rank) { check(bbmcu_loss_peer_connect_ptrs(_loss.get(), windows_by_rank.data()), _fitted.ctx().get()); }
  // write the live parameters back into the fitted BSDF (toString / export)
  void commit() { _fitted.set_parameter_values(*_params); }
  cuda_bsdf& fitted() const { return _fitted; }

private:
  void create(int m, const std::optional<bbmcu_spherical_grid>& grid, bsdf_flag component, const bbmcu_bsdf* ref_b, const float* ref_t, uint64_t first, uint64_t count)
  {
    bbmcu_loss* L = nullptr;
    check(bbmcu_loss_create(_fitted.ctx().get(), m, grid ? &*grid : nullptr, (int)component, BBMCU_RADIANCE, ref_b, ref_t, first, count, &L), _fitted.ctx().get());
    _loss.reset(L, bbmcu_loss_free);
  }
  cuda_bsdf& _fitted;
  std::vector<double>* _params;
  std::shared_ptr<bbmcu_loss> _loss;
};

} // namespace bbmcu
