// Fused loss(+gradient) pass: per sample read the direction pair and the tabulated reference value,
// evaluate the fitted BSDF (and its parameter jacobian through dual numbers), apply the metric and
// reduce - replaces the scalar loop of include/bbm/sampledlossfunction.h:78-87.
//
// Two launch shapes.  Compile-time BSDF shapes (one model, or Aggregate(Lambertian, model) - every entry of the
// reference's fits/*.fit): the sample-stationary tile kernel below.  Arbitrary run-time aggregates: k_loss_generic
// (bbmcu_loss.cu), grid (blocks_x, K), block (bx, k) = parameter set k over a strided slice of the samples.
// Either way one partial row per block, and a second tiny kernel adds the rows in fixed order, so the result is
// deterministic and independent of scheduling (totals are judged against a double accumulation
// of the per-sample terms).
#pragma once
#include <cstdlib>
#include "bbmcu_ctx.hpp"
#include "bbmcu_lossop.cuh"
#include "bbmcu_tables.cuh"

namespace bbmcu {

enum : int { LIN_MATERIALISED = 0, LIN_MERL_TABLES = 1, LIN_SPHERICAL = 2 };
struct LossArgs
{
  // where a sample's direction pair comes from: LIN_MATERIALISED reads the planes `in` / `out` (24 B per sample);
  // LIN_MERL_TABLES forms merl_linearizer(first + i) from the 900-float separable table `lin_tab` staged in shared
  // memory; LIN_SPHERICAL forms spherical_linearizer(first + i) from the per-loss table `lin_tab` of its phi / theta
  // samples (global memory, L1-resident).  The fused modes return the very bits
  // the materialised planes hold (same functions), so a pass reads only the 12 B per sample of measured data.
  int lin_mode;
  const float* lin_tab;
  uint64_t first;         // linearizer index of the shard's sample 0 (contiguous shards)
  int il_world, il_rank;  // il_world > 1: the shard is blocks il_rank, il_rank + il_world, ... of kTileSamples consecutive samples
  SphericalGrid grid;
  const float* in;        // 3 planes of n (LIN_MATERIALISED)
  const float* out;       // 3 planes of n
  const float* ref;       // M x 3 planes of n (reference operand tabulated at the samples, one set of planes per material)
  size_t ref_stride;      // floats between the planes of consecutive materials (3 n)
  int k_per_material;     // parameter sets per material (block z = material; sets m*K .. m*K + K - 1 belong to material m)
  size_t n;               // samples in this shard
  const float* attrs;     // (M x K) x attr_stride attribute blocks
  int attr_stride;
  int n_attrs;
  int metric, component;
  int want_grad;
  double* partial;        // K x (1 + P) x blocks_x
  int P;
  int sm_count;
  int n_materials;        // grid z of the tile kernel
  // small batches (one compass step: K = 2P parameter sets) travel inside the kernel arguments: no staging buffer, no
  // host-to-device copy, no event on the path.  inline_count = K * n_attrs floats (0: read `attrs`)
  int inline_count;
  float inline_attrs[256];
};
constexpr int kInlineAttrFloats = 256;
BBMCU_D float loss_attr(const LossArgs& a, size_t idx) { return a.inline_count ? a.inline_attrs[idx] : a.attrs[idx]; }

constexpr int kLossThreads = 256;
constexpr int kTileSPT = 4;
constexpr int kTileKChunk = 8;
constexpr int kTileSamples = kTileSPT * kLossThreads;

// linearizer index of sample i of the shard
BBMCU_D uint64_t loss_lin_index(const LossArgs& a, size_t i)
{
  if(a.il_world > 1) { const size_t t = i / kTileSamples; return ((uint64_t)t * (uint64_t)a.il_world + (uint64_t)a.il_rank) * kTileSamples + (uint64_t)(i - t*kTileSamples); }
  return a.first + i;
}
// sample i of the shard: direction pair (generated or loaded) - s_lin is the shared-memory copy of a.lin_tab
BBMCU_D void loss_dirs(const LossArgs& a, const float* s_lin, size_t i, f3& in, f3& out)
{
  if(a.lin_mode == LIN_MERL_TABLES) merl_dirs_tab(s_lin, (uint32_t)loss_lin_index(a, i), in, out);
  else if(a.lin_mode == LIN_SPHERICAL) spherical_dirs_tab(a.grid, a.lin_tab, loss_lin_index(a, i), in, out);
  else
  {
#ifdef __CUDA_ARCH__
    in = make_f3(__ldg(a.in + i), __ldg(a.in + a.n + i), __ldg(a.in + 2*a.n + i));
    out = make_f3(__ldg(a.out + i), __ldg(a.out + a.n + i), __ldg(a.out + 2*a.n + i));
#else
    in = make_f3(a.in[i], a.in[a.n + i], a.in[2*a.n + i]);
    out = make_f3(a.out[i], a.out[a.n + i], a.out[2*a.n + i]);
#endif
  }
}
#ifdef __CUDACC__
// stage the linearizer table (3.6 KB) in shared memory; no-op for the other modes.  Callers __syncthreads() afterwards.
__device__ __forceinline__ void loss_stage_lin(const LossArgs& a, float* s_lin)
{
  if(a.lin_mode == LIN_MERL_TABLES) for(int i = threadIdx.x; i < kMerlLinTabFloats; i += blockDim.x) s_lin[i] = __ldg(a.lin_tab + i);
}
#endif

// LossT::sample(attrs, metric, component, in, out, ref, grad[P], want_grad) -> e   (LossSingle / LossPair)
// ---- sample-stationary variant: the launch shape for batched passes --------------------------------
// A block owns a tile of kTileSPT * 256 samples, loads them ONCE into registers and loops over its range of parameter
// sets (attribute blocks staged in shared memory).  Everything that depends on the directions only - half vector,
// dots, geometric terms, metric weights - is loop invariant and hoisted out of the parameter loop by the compiler, and
// the 36 B/sample of L2 traffic are paid once per tile instead of once per parameter set.  Per-thread sums are float
// over the thread's kTileSPT samples, warp-shuffled in float (128 terms), and enter FP64 at the cross-warp step; block
// partials are written per (parameter set, tile) and added in fixed order by k_loss_finish: deterministic, and a
// parameter set's result does not depend on which other sets share the launch.
// Warp reduction of C values per lane with a halving butterfly: at each offset a lane keeps one half of its values
// and ships the other half to its partner, so C values cost CP/2 + CP/4 + ... + 1 (+ the remaining plain steps)
// shuffles instead of 5 C.  Afterwards the lane with (lane & rest) == 0 holds the full sum of value `idx`.
template<int CNT, int OFF> struct WarpButterfly
{
  static __device__ __forceinline__ void run(float* w, int lane, int& idx)
  {
    if constexpr (OFF >= 1)
    {
      if constexpr (CNT > 1)
      {
        constexpr int H = CNT / 2;
        const bool bit = (lane & OFF) != 0;
#pragma unroll
        for(int i=0; i < H; ++i)
        {
          const float send = bit ? w[i] : w[i + H];
          const float keep = bit ? w[i + H] : w[i];
          w[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        if(bit) idx += H;
        WarpButterfly<H, OFF/2>::run(w, lane, idx);
      }
      else
      {
        w[0] += __shfl_xor_sync(0xffffffffu, w[0], OFF);
        WarpButterfly<1, OFF/2>::run(w, lane, idx);
      }
    }
  }
};
// sums v[0..C) over the warp; lane_is_writer lanes get (idx, value) with idx < CP (idx >= C: padding, skip)
template<int C> __device__ __forceinline__ void warp_reduce_multi(const float (&v)[C], int lane, int& idx, float& value, bool& writer)
{
  constexpr int CP = C <= 1 ? 1 : C <= 2 ? 2 : C <= 4 ? 4 : C <= 8 ? 8 : C <= 16 ? 16 : 32;
  float w[CP];
#pragma unroll
  for(int i=0; i < CP; ++i) w[i] = (i < C) ? v[i < C ? i : 0] : 0.0f;
  idx = 0;
  WarpButterfly<CP, 16>::run(w, lane, idx);
  value = w[0];
  // offsets 16, 8, ... were used for halving while the count was > 1: log2(CP) of them; the rest were plain sums
  constexpr int used = CP == 1 ? 0 : CP == 2 ? 16 : CP == 4 ? 24 : CP == 8 ? 28 : CP == 16 ? 30 : 31;
  writer = (lane & ~used & 31) == 0;
}

// (launch bounds: block size only, which lets the compiler settle at 128 registers = 2 resident blocks: 2.65 ms for K = 256;
// measured alternatives: (256, 1) -> 255 registers allowed, 3.81 ms; (256, 2) 2.71 ms; (256, 3) -> 80 registers, spills, 2.80 ms)
template<class LossT, bool WG>
__global__ void __launch_bounds__(kLossThreads) k_loss_tile(const LossArgs a, int K, int k_per_block, int n_tiles)
{
  constexpr int P = LossT::P;
  constexpr int C = WG ? 1 + P : 1;
  extern __shared__ float s_attr_tile[];                       // (k1 - k0) x n_attrs
  __shared__ float s_red[kTileKChunk][kLossThreads/32][C];
  __shared__ float s_lin[kMerlLinTabFloats];
  const int mat = blockIdx.z;
  const int k0 = blockIdx.y * k_per_block, k1 = min(K, k0 + k_per_block);
  const size_t kbase = (size_t)mat * K;                        // first parameter set of this material
  for(int i = threadIdx.x; i < (k1 - k0)*a.n_attrs; i += blockDim.x)
    s_attr_tile[i] = loss_attr(a, (kbase + k0 + i / a.n_attrs)*a.attr_stride + (i % a.n_attrs));
  loss_stage_lin(a, s_lin);
  __syncthreads();
  const float* refp = a.ref + (size_t)mat * a.ref_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // a block walks over tiles blockIdx.x, blockIdx.x + gridDim.x, ...: with few parameter sets per block (small K, or many
  // materials) the launch uses fewer, longer-lived blocks so that the staging above is paid once per block, not per tile;
  // partial rows stay per (parameter set, tile), so the result does not depend on the launch shape
  for(int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
  {
    f3 in[kTileSPT], out[kTileSPT]; Spec<float> ref[kTileSPT]; bool valid[kTileSPT];
#pragma unroll
    for(int s=0; s < kTileSPT; ++s)
    {
      const size_t i = (size_t)tile*kTileSamples + (size_t)s*kLossThreads + threadIdx.x;
      valid[s] = i < a.n;
      const size_t ii = valid[s] ? i : 0;
      loss_dirs(a, s_lin, ii, in[s], out[s]);
      ref[s] = Spec<float>(__ldg(refp + ii), __ldg(refp + a.n + ii), __ldg(refp + 2*a.n + ii));
    }
    typename LossT::Geom geo[kTileSPT];                          // direction-only work, once per sample
#pragma unroll
    for(int s=0; s < kTileSPT; ++s) geo[s] = LossT::geom(a.metric, in[s], out[s]);
    for(int kc = k0; kc < k1; kc += kTileKChunk)
    {
      const int nkk = min(kTileKChunk, k1 - kc);
      for(int kk=0; kk < nkk; ++kk)
      {
        const float* at = s_attr_tile + (size_t)(kc + kk - k0)*a.n_attrs;
        float acc[C];
#pragma unroll
        for(int j=0; j < C; ++j) acc[j] = 0.0f;
#pragma unroll
        for(int s=0; s < kTileSPT; ++s)
        {
          if(!valid[s]) continue;
          float g[P];
          const float e = LossT::sample(at, a.metric, a.component, geo[s], in[s], out[s], ref[s], g, WG);
          acc[0] += e;
          if(WG) {
#pragma unroll
            for(int j=0; j < P; ++j) acc[1 + j] += g[j];
          }
        }
        int ridx; float rval; bool rwriter;
        warp_reduce_multi<C>(acc, lane, ridx, rval, rwriter);
        if(rwriter && ridx < C) s_red[kk][warp][ridx] = rval;
      }
      __syncthreads();
      // one value per (parameter set of the chunk, column); a loop, so any P up to kMaxParams is covered
      for(int t = threadIdx.x; t < nkk*(1 + a.P); t += blockDim.x)
      {
        const int kk = t / (1 + a.P), j = t % (1 + a.P);
        double v = 0.0;
        if(j < C) {
#pragma unroll
          for(int w=0; w < kLossThreads/32; ++w) v += (double)s_red[kk][w][j];
        }
        a.partial[((kbase + kc + kk)*(1 + a.P) + j)*n_tiles + tile] = v;          // [material, k][column][tile]: the finish kernel reads tiles coalesced
      }
      __syncthreads();
    }
  }
}

// launch shape of the tile kernel for n samples, K parameter sets of n_attrs floats: tiles x k-splits
// (Ktot = parameter sets of the whole launch, over all materials; K = per material: the k-split happens inside a material).
// A block costs its tile's direction / geometry work once plus one evaluation per parameter set of its range, blocks are
// equal, and `slots` of them are resident at a time: the launch takes ceil(blocks / slots) block durations.  The split is
// the one that minimises that product - e.g. a 1/8 shard (178 tiles) at K = 256 takes 8 ranges of 32 sets (1424 blocks, 5
// rounds of 296) rather than 14 of 19 (9 rounds): 347 -> 322 us (tools/loss_shape_sweep.py).  BBMCU_LOSS_BLOCKS_PER_SM (tuning)
// restores the fixed blocks-per-SM target.
inline void loss_tile_shape(size_t n, size_t Ktot, size_t K, int n_attrs, int sm_count, int blocks_per_sm, unsigned& tiles, unsigned& ksplit, int& k_per_block,
                            size_t smem_budget = 40*1024, double w_set = 185.0, double w_tile = 215.0)
{
  tiles = (unsigned)((n + kTileSamples - 1) / kTileSamples);
  if(tiles < 1) tiles = 1;
  const size_t smem_k = smem_budget / ((size_t)n_attrs*sizeof(float));          // parameter sets that fit the shared-memory budget
  size_t split_fit = (K + smem_k - 1) / (smem_k ? smem_k : 1);
  if(split_fit < 1) split_fit = 1;
  static const int fill = [] { const char* e = std::getenv("BBMCU_LOSS_BLOCKS_PER_SM"); int v = e ? std::atoi(e) : 0; return v > 0 ? v : 0; }();
  const size_t mats = K ? (Ktot + K - 1) / K : 1;
  size_t sp = split_fit;
  if(fill > 0)
  {
    size_t split_fill = ((size_t)sm_count*fill + tiles*mats - 1) / (tiles*mats);
    sp = split_fit > split_fill ? split_fit : split_fill;
  }
  else
  {
    // relative cost of one parameter set and of the per-tile work of a sample (instructions, loss + gradient of the
    // Cook-Torrance aggregate: profiles/r02_s3_ncu_loss_tile_ct_fused_linearizer.txt) are the defaults of w_set, w_tile
    const size_t slots = (size_t)sm_count * (size_t)(blocks_per_sm > 0 ? blocks_per_sm : 2);
    // two passes: the cheapest shape, then the finest split within 2 % of it (equal on paper, but short blocks even out
    // the run-time differences between blocks: full grid at K = 256, 1 range 2276 us, 2 ranges 2218, 4 ranges 2200)
    auto cost_of = [&](size_t ks, bool& valid) {
      const size_t kpb = (K + ks - 1) / ks, ks_eff = (K + kpb - 1) / kpb;
      valid = ks_eff == ks;                                      // otherwise the same shape as a smaller ks
      const size_t blocks = (size_t)tiles * ks * mats, rounds = (blocks + slots - 1) / slots;
      return (double)rounds * ((double)kpb * w_set + w_tile);
    };
    double best = 0.0;
    for(size_t ks = split_fit; ks <= K && ks <= 256; ++ks) { bool v; const double c = cost_of(ks, v); if(v && (best == 0.0 || c < best)) best = c; }
    for(size_t ks = split_fit; ks <= K && ks <= 256; ++ks) { bool v; const double c = cost_of(ks, v); if(v && c <= best * 1.02) sp = ks; }
  }
  if(sp > K) sp = K;
  if(sp < 1) sp = 1;
  k_per_block = (int)((K + sp - 1) / sp);
  ksplit = (unsigned)((K + k_per_block - 1) / k_per_block);
}

// internal linkage (see bbmcu_tables.cuh).  blocks_x is the tile count the caller sized `partial` with.
// K = parameter sets per material, a.n_materials materials (grid z)
template<class LossT> static void launch_loss_static(cudaStream_t s, const LossArgs& a, unsigned blocks_x, unsigned K)
{
  bind_device_tables();
  // resident blocks per SM of the two kernels (registers decide: 2 at 122 registers); shared memory stays below the budget
  static int per_sm[2] = {0, 0};
  const int wg = a.want_grad ? 1 : 0;
  if(per_sm[wg] == 0)
  {
    int v = 0;
    cudaError_t e = wg ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_loss_tile<LossT, true>, kLossThreads, 40*1024)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_loss_tile<LossT, false>, kLossThreads, 40*1024);
    per_sm[wg] = (e == cudaSuccess && v > 0) ? v : 2;
  }
  unsigned tiles, ksplit; int kpb;
  loss_tile_shape(a.n, (size_t)K*a.n_materials, K, a.n_attrs, a.sm_count, per_sm[wg], tiles, ksplit, kpb);
  const size_t smem = (size_t)kpb*a.n_attrs*sizeof(float);
  (void)blocks_x;
  // blocks along the tile axis: all tiles when the launch is small, else a whole number of rounds of resident blocks, each
  // block walking over several tiles (K = 1 passes and many-material launches: the table / attribute staging is then paid
  // once per block)
  const size_t slots = (size_t)a.sm_count * (size_t)per_sm[wg];
  static const int small_rounds = [] { const char* e = std::getenv("BBMCU_LOSS_SMALLK_ROUNDS"); int v = e ? std::atoi(e) : 0; return v > 0 ? v : 2; }();   // tuning
  const size_t target = slots * (kpb >= 8 ? 8 : (size_t)small_rounds), other = (size_t)ksplit * a.n_materials;
  unsigned gx = tiles;
  if((size_t)tiles * other > target) gx = (unsigned)std::max<size_t>(1, (target + other - 1) / other);
  if(gx > tiles) gx = tiles;
  const dim3 grid(gx, ksplit, (unsigned)a.n_materials);
  if(a.want_grad) k_loss_tile<LossT, true><<<grid, kLossThreads, smem, s>>>(a, (int)K, kpb, (int)tiles);
  else            k_loss_tile<LossT, false><<<grid, kLossThreads, smem, s>>>(a, (int)K, kpb, (int)tiles);
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
