// Holzschuch-Pacanowski G1 table on the GPU: what the reference's generator precompute/HolzschuchPacanowski/G1.cpp
// computes in hours on one core (100 values of p x 1000 values of tan(theta) x a 10 000-point quadrature of P2, Eq. 36),
// restated as one kernel.
//
//   P2(r, p) = 2 N sum_k dq_k exp(-(r^2 + q_k^2)^p),  N = p / (pi Gamma(1/p))                     G1.cpp:88-108
//   Delta_j  = (Delta_{j-1} + P_{j-1}) tan_j / tan_{j-1} - P_{j-1} + (r_j tan_j - 1)+ P2(r_j) dr_j  G1.cpp:166-213
//   G1_j     = 1 / (1 + Delta_j)
// with the exponentiated-log abscissae conv(x) = log(1/x)^20.  The abscissae and weights (q_k, dq_k, r_j, dr_j, tan_j)
// are differences of nearly equal float powers - they are what they are because of the host libm - so they are computed
// on the host exactly as the generator does (same float / double mix, same decrementing float loop variable).  The 1e9
// smooth, positive terms of the quadrature are the GPU's part: one thread per (p, j), sequential float accumulation in the
// generator's order.  The 1000-step recurrence per p is sequential and cheap; it runs on the host in float like the
// generator's.  Result: the 100 x 1000 table of include/precomputed/holzschuchpacanowski/G1.h to its printed 6 digits.
#include <cmath>
#include <vector>

#include "bbmcu_ctx.hpp"
#include "bbmcu_hpnorm.cuh"

using namespace bbmcu;

namespace {

constexpr int kP = 100, kJ = 1000;

__global__ void __launch_bounds__(128) k_hp_p2(const float* __restrict__ q, const float* __restrict__ dq, int nk,
                                               const float* __restrict__ r, const float* __restrict__ norm, float* __restrict__ p2)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;       // 1 .. kJ-1 used
  const int pi = blockIdx.y;
  if(j < 1 || j >= kJ) return;
  const float p = 5.0f / (float)(pi + 1);
  const float rj = r[j], r2 = rj*rj;
  float integral = 0.0f;
  for(int k=0; k < nk; ++k)
  {
    const float qk = q[k];
    integral += dq[k] * expf(-powf(r2 + qk*qk, p));
  }
  p2[pi*kJ + j] = (float)(2.0 * (double)norm[pi] * (double)integral);
}

// ---- the renormalisation table sigma_rel^2 / sigma_s^2 (precompute/HolzschuchPacanowski/normalization.cpp) ---------------
// 100 x 100 x 100 entries over (b, c, sin theta_i); each is the generator's integralSH (normalization.cpp:133-155): a
// sequential sum over f = 1 - sin .. 1 + sin in steps of 0.01 degree of  alpha(f) f df S_HS(f)  - up to 11 459 terms of a
// double pow and a double acos - plus the closed form of the inner disc.  The generator runs the 5.7e9 terms on one core;
// here one thread owns one entry and walks it in the generator's order with the generator's float / double mix (float f
// and float running sum, double pow / acos / products), so an entry differs from the generator's only where the device's
// double pow / acos round differently from the host libm's.  Threads of a warp share sin theta_i (equal trip counts).
__global__ void __launch_bounds__(128) k_hp_normalization(float* __restrict__ table)
{
  const int si = blockIdx.y;                                   // sin(theta) index: uniform per block
  const int bc = blockIdx.x * blockDim.x + threadIdx.x;        // b index * 100 + c index
  if(bc >= kHpNormN*kHpNormN) return;
  const int bi = bc / kHpNormN, ci = bc % kHpNormN;
  table[((size_t)bi*kHpNormN + ci)*kHpNormN + si] = hp_normalization_entry(bi, ci, si);
}

float conv_f(float x) { return std::pow(std::log(1.0f / x), 20.0f); }
float conv_d(double x) { return (float)std::pow(std::log(1.0 / x), (double)20.0f); }

} // anonymous namespace

extern "C" int bbmcu_hp_precompute_g1(bbmcu_ctx* ctx, float* table)
{
  return guarded(ctx, [&] {
    if(!ctx || !table) throw std::invalid_argument("BBM: null argument");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    // ---- host: quadrature nodes of P2 (G1.cpp:92-104; the loop variable is a float that is decremented) -----------
    std::vector<float> q, dq;
    const float deltax = 0.0001f;
    for(float x = 1.0f; x > deltax; x -= deltax)
    {
      float d = conv_f(x - deltax) - conv_f(x);
      float qq = conv_d((double)x - 0.5*(double)deltax);
      if(std::isnan(d)) continue;
      q.push_back(qq); dq.push_back(d);
    }
    // ---- host: the tan(theta) grid (G1.cpp:176-201) ---------------------------------------------------------------
    const float delta_x = 1.0f / (float)kJ;
    std::vector<float> tanT(kJ, 0.0f), dr(kJ, 0.0f), r(kJ, 1.0f);
    for(int j=1; j < kJ; ++j)
    {
      float x = (float)(j + 1) / (float)kJ;
      tanT[j] = 1.0f / conv_f(x);
      dr[j] = conv_f(x - delta_x) - conv_f(x);
      r[j] = conv_d((double)x - 0.5*(double)delta_x);
    }
    std::vector<float> norm(kP);
    for(int pi=0; pi < kP; ++pi) { float p = 5.0f / (float)(pi + 1); norm[pi] = (float)((double)p / ((double)(float)M_PI * std::tgamma(1.0 / (double)p))); }
    // ---- device: 100 x 999 quadratures of nk terms ----------------------------------------------------------------------
    float *d_q, *d_dq, *d_r, *d_norm, *d_p2;
    const int nk = (int)q.size();
    BBMCU_CUDA(cudaMalloc(&d_q, nk*sizeof(float))); BBMCU_CUDA(cudaMalloc(&d_dq, nk*sizeof(float)));
    BBMCU_CUDA(cudaMalloc(&d_r, kJ*sizeof(float))); BBMCU_CUDA(cudaMalloc(&d_norm, kP*sizeof(float))); BBMCU_CUDA(cudaMalloc(&d_p2, kP*kJ*sizeof(float)));
    BBMCU_CUDA(cudaMemcpyAsync(d_q, q.data(), nk*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BBMCU_CUDA(cudaMemcpyAsync(d_dq, dq.data(), nk*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BBMCU_CUDA(cudaMemcpyAsync(d_r, r.data(), kJ*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BBMCU_CUDA(cudaMemcpyAsync(d_norm, norm.data(), kP*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    BBMCU_CUDA(cudaMemsetAsync(d_p2, 0, kP*kJ*sizeof(float), ctx->stream));
    k_hp_p2<<<dim3((kJ + 127)/128, kP), 128, 0, ctx->stream>>>(d_q, d_dq, nk, d_r, d_norm, d_p2);
    BBMCU_CUDA(cudaGetLastError());
    ++ctx->launches;
    std::vector<float> P2(kP*kJ);
    BBMCU_CUDA(cudaMemcpyAsync(P2.data(), d_p2, kP*kJ*sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_q); cudaFree(d_dq); cudaFree(d_r); cudaFree(d_norm); cudaFree(d_p2);
    // ---- host: the incremental Delta recurrence, float, in the generator's order (G1.cpp:203-218) ---------------------------
    for(int pi=0; pi < kP; ++pi)
    {
      float* integral = table + (size_t)pi*kJ;
      integral[0] = 0.0f;
      float prevTan = 0.0f, Pj = 0.0f;
      for(int j=1; j < kJ; ++j)
      {
        const float t = tanT[j];
        if(!std::isinf(t))
        {
          const float p2 = P2[pi*kJ + j] * dr[j];
          integral[j] = 0.0f;
          if(prevTan > 0.0f) integral[j] = (integral[j-1] + Pj) * t / prevTan - Pj;
          if(r[j]*t > 1.0f) integral[j] += (r[j]*t - 1.0f)*p2;
          prevTan = t;
          Pj += p2;
        }
        else integral[j] = t;
      }
      for(int j=0; j < kJ; ++j) integral[j] = (float)(1.0 / (1.0 + (double)integral[j]));
    }
  });
}

extern "C" int bbmcu_hp_precompute_normalization(bbmcu_ctx* ctx, float* table)
{
  return guarded(ctx, [&] {
    if(!ctx || !table) throw std::invalid_argument("BBM: null argument");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    const size_t n = (size_t)kHpNormN*kHpNormN*kHpNormN;
    float* d_table = nullptr;
    BBMCU_CUDA(cudaMalloc(&d_table, n*sizeof(float)));
    k_hp_normalization<<<dim3((kHpNormN*kHpNormN + 127)/128, kHpNormN), 128, 0, ctx->stream>>>(d_table);
    cudaError_t e = cudaGetLastError();
    if(e == cudaSuccess) e = cudaMemcpyAsync(table, d_table, n*sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if(e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_table);
    ++ctx->launches;
    if(e != cudaSuccess) throw CudaError(std::string("bbmcu_hp_precompute_normalization: ") + cudaGetErrorString(e));
  });
}
