// Fused loss(+gradient) pass: per sample read the direction pair and the tabulated reference value,
// evaluate the fitted BSDF (and its parameter jacobian through dual numbers), apply the metric and
// reduce - replaces the scalar loop of include/bbm/sampledlossfunction.h:78-87.
//
// Launch shape: grid (blocks_x, K) - block (bx, k) handles parameter set k over a strided slice of
// the shard's samples.  Reduction: FP64 per-thread accumulators -> warp shuffle -> shared ->
// one partial row per block; a second tiny kernel adds the partial rows in fixed order, so the
// result is deterministic and independent of scheduling (totals are judged
// against a double accumulation of the per-sample terms).
#pragma once
#include "bbmcu_ctx.hpp"
#include "bbmcu_lossop.cuh"
#include "bbmcu_tables.cuh"

namespace bbmcu {

struct LossArgs
{
  const float* in;        // 3 planes of n
  const float* out;       // 3 planes of n
  const float* ref;       // 3 planes of n (reference BSDF tabulated at the samples)
  size_t n;               // samples in this shard
  const float* attrs;     // K x attr_stride attribute blocks
  int attr_stride;
  int n_attrs;
  int metric, component;
  int want_grad;
  double* partial;        // K x blocks_x x (1 + P)
  int P;
};

constexpr int kLossThreads = 256;

// LossT::sample(attrs, metric, component, in, out, ref, grad[P], want_grad) -> e   (LossSingle / LossPair)
template<class LossT>
__global__ void __launch_bounds__(kLossThreads) k_loss_static(const LossArgs a)
{
  constexpr int P = LossT::P;
  __shared__ float s_attr[kMaxAttrs];
  __shared__ double s_red[kLossThreads/32][1 + P];
  const int k = blockIdx.y;
  for(int i = threadIdx.x; i < a.n_attrs; i += blockDim.x) s_attr[i] = a.attrs[(size_t)k*a.attr_stride + i];
  __syncthreads();
  double acc[1 + P];
#pragma unroll
  for(int j=0; j <= P; ++j) acc[j] = 0.0;
  const bool wg = a.want_grad != 0;
  for(size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x*blockDim.x)
  {
    f3 in = make_f3(__ldg(a.in + i), __ldg(a.in + a.n + i), __ldg(a.in + 2*a.n + i));
    f3 out = make_f3(__ldg(a.out + i), __ldg(a.out + a.n + i), __ldg(a.out + 2*a.n + i));
    Spec<float> ref(__ldg(a.ref + i), __ldg(a.ref + a.n + i), __ldg(a.ref + 2*a.n + i));
    float g[P];
    float e = LossT::sample(s_attr, a.metric, a.component, in, out, ref, g, wg);
    acc[0] += (double)e;
    if(wg) {
#pragma unroll
      for(int j=0; j < P; ++j) acc[1 + j] += (double)g[j];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for(int j=0; j <= P; ++j)
  {
    double v = acc[j];
#pragma unroll
    for(int o=16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if(lane == 0) s_red[warp][j] = v;
  }
  __syncthreads();
  if(threadIdx.x <= P)
  {
    double v = 0.0;
#pragma unroll
    for(int w=0; w < kLossThreads/32; ++w) v += s_red[w][threadIdx.x];
    a.partial[((size_t)k*gridDim.x + blockIdx.x)*(1 + a.P) + threadIdx.x] = v;
  }
}

template<class LossT> static void launch_loss_static(cudaStream_t s, const LossArgs& a, unsigned blocks_x, unsigned K)
{
  bind_device_tables();
  k_loss_static<LossT><<<dim3(blocks_x, K), kLossThreads, 0, s>>>(a);
}

// one translation unit per group of models (compile time); returns false if `model` is not in that group
bool launch_loss_single_g0(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);
bool launch_loss_single_g1(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);
bool launch_loss_single_g2(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);
bool launch_loss_single_g3(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);
bool launch_loss_pair_g0(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);   // Aggregate(Lambertian, model)
bool launch_loss_pair_g1(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);
bool launch_loss_pair_g2(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);
bool launch_loss_pair_g3(int model, cudaStream_t, const LossArgs&, unsigned, unsigned);

#define BBMCU_LOSS_CASE_SINGLE(m) case m: launch_loss_static<LossSingle<typename ModelOf<m>::type>>(s, a, bx, K); return true;
#define BBMCU_LOSS_CASE_PAIR(m)   case m: launch_loss_static<LossPair<Lambertian, typename ModelOf<m>::type>>(s, a, bx, K); return true;

} // namespace bbmcu
