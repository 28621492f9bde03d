// checkBsdf kernels (bbmcu_check.cuh), kind "s": eval = false, sample = true - one instance per model + the run-time lobe list
#include "bbmcu_launch.cuh"
#include "bbmcu_check.cuh"
namespace bbmcu {
void launch_check_s(bbmcu_ctx* ctx, cudaStream_t s, const BsdfDesc& d, const CheckArgs& a, dim3 grid)
{
  bind_device_tables();
  auto go = [&](auto* tag) {
    using B = typename std::remove_pointer<decltype(tag)>::type;
    k_check<CheckView<B, false, true>><<<grid, kCheckThreads, 0, s>>>(a, d);
    BBMCU_CUDA(cudaGetLastError());
    ++ctx->launches;
  };
  if(!d.aggregate && d.n_lobes == 1)
  {
    bool ok = dispatch_model_host(d.model[0], [&](auto* m) { using M = typename std::remove_pointer<decltype(m)>::type; go((BsdfSingle<M>*)nullptr); });
    if(!ok) throw std::invalid_argument("BBM: model id " + std::to_string(d.model[0]) + " has no CUDA kernel in this build");
  }
  else go((BsdfGeneric*)nullptr);
}
}
