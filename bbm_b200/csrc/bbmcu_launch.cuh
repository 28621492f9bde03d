// Host-side launch helper: pick the compile-time single-model kernel when the BSDF is one model,
// the run-time lobe-list kernel otherwise.
#pragma once
#include <cstdlib>
#include "bbmcu_ctx.hpp"
#include "bbmcu_kernels.cuh"
#include "bbmcu_tables.cuh"

namespace bbmcu {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// internal linkage: every translation unit keeps its own copy next to its own device symbols (bbmcu_tables.cuh)
template<class Op> static void launch_foreach4(bbmcu_ctx* ctx, cudaStream_t stream, const Op& op_in, size_t n)
{
  if(n == 0) return;
  Op op = op_in;
  // planes of an SoA argument are n floats apart unless the context says otherwise (bbmcu_set_plane_stride; the host-pointer
  // path stages with a padded stride).  16-byte vector access needs every plane start aligned: callers check the base
  // pointers, the stride is checked here - so a batch of 2^26 + 1 elements keeps its vector loads when the planes are padded.
  op.ld = ctx->ld ? ctx->ld : op.n;
  if(op.ld % 4 != 0) op.aligned = false;
  size_t groups = (n + kVec - 1) / kVec;
  bind_device_tables();
  unsigned grid = grid_for(ctx, groups, Op::kBlock, 8*256/Op::kBlock);
  if constexpr (UsesLinTab<Op>::value)
  {
    // every block stages the 3.6 KB linearizer table first: one resident wave striding over the grid pays that once per SM slot
    // (the whole MERL grid is 1424 blocks' worth of work - a prologue per block would be a fifth of the kernel)
    static int per_sm = 0;
    if(per_sm == 0)
    {
      BBMCU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_foreach4<Op>, Op::kBlock, 0));
      if(per_sm < 1) per_sm = 1;
    }
    const unsigned wave = (unsigned)ctx->sm_count * (unsigned)per_sm;
    if(grid > wave) grid = wave;
  }
  if constexpr (Op::kHasBsdf)
  {
    if constexpr (Op::kTables)
    {
      if(Op::kOneWaveWithTables && op.bsdf.n_tables)
      {
        // every block rebuilds the 90-bin sampling CDF in its prologue (90 back-scatter evaluations of the model): launch
        // one resident wave and let it stride over the batch, so the prologue is paid once per SM slot
        static int per_sm = 0;                                     // per operator instantiation
        if(per_sm == 0)
        {
          BBMCU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_foreach4<Op>, Op::kBlock, 0));
          if(per_sm < 1) per_sm = 1;
        }
        const unsigned wave = (unsigned)ctx->sm_count * (unsigned)per_sm;
        if(grid > wave) grid = wave;
      }
    }
  }
  k_foreach4<Op><<<grid, Op::kBlock, 0, stream>>>(op, groups);
  BBMCU_CUDA(cudaGetLastError());
  ++ctx->launches;
}

// OpT<B> is one of the BSDF operators of bbmcu_kernels.cuh; fill(op) sets everything but op.bsdf
template<template<class> class OpT, class Fill>
static void launch_bsdf_op(bbmcu_ctx* ctx, cudaStream_t stream, const BsdfDesc& d, size_t n, Fill&& fill)
{
  auto go = [&](auto* tag) {
    using B = typename std::remove_pointer<decltype(tag)>::type;
    OpT<B> op; op.bsdf = d; fill(op);
    launch_foreach4(ctx, stream, op, n);
  };
  if(!d.aggregate && d.n_lobes == 1)
  {
    bool ok = dispatch_model_host(d.model[0], [&](auto* m) { using M = typename std::remove_pointer<decltype(m)>::type; go((BsdfSingle<M>*)nullptr); });
    if(!ok) throw std::invalid_argument("BBM: model id " + std::to_string(d.model[0]) + " has no CUDA kernel in this build");
  }
  else go((BsdfGeneric*)nullptr);
}

// Aggregate(Lambertian, M), M without device tables: the two-model kernels (BsdfPair).  Returns false for any other shape.
template<template<class> class OpT, class Fill>
static bool launch_pair_op(bbmcu_ctx* ctx, cudaStream_t stream, const BsdfDesc& d, size_t n, Fill&& fill)
{
  if(!(d.aggregate && d.n_lobes == 2 && d.model[0] == M_Lambertian && d.n_tables == 0)) return false;
  // BBMCU_DISABLE_PAIR_KERNELS=1 keeps such aggregates on the run-time lobe list (tests compare the two paths bit for bit)
  static const bool disabled = [] { const char* e = std::getenv("BBMCU_DISABLE_PAIR_KERNELS"); return e && e[0] == '1'; }();
  if(disabled) return false;
  bool launched = false;
  dispatch_model_host(d.model[1], [&](auto* m) {
    using M = typename std::remove_pointer<decltype(m)>::type;
    if constexpr (TableFloats<M>::N == 0)
    {
      OpT<BsdfPair<Lambertian, M>> op; op.bsdf = d; fill(op);
      launch_foreach4(ctx, stream, op, n);
      launched = true;
    }
  });
  return launched;
}
bool launch_pair_eval(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* in, const float* out, float* rgb, size_t n, bool aligned);
bool launch_pair_pdf(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* in, const float* out, float* pdf, size_t n, bool aligned);
bool launch_pair_sample(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* out, const float* xi, float* dir, float* pdf, int32_t* flag, size_t n, bool aligned);
// generated inputs of the fused pass (SampleEvalPdfOp): gen != 0 draws (out, xi) of element i from (seed, first + i)
struct GenArgs { int gen = 0; uint64_t seed = 0, first = 0; float* out = nullptr; float* xi = nullptr; };
bool launch_pair_sample_eval_pdf(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* out, const float* xi,
                                 float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf, size_t n, bool aligned);
void launch_sample_eval_pdf_gen(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component,
                                float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf, size_t n, bool aligned, const GenArgs& g);
bool launch_pair_eval_grid(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* lin_tab, uint32_t first, float* rgb, float* in, float* out, size_t n, bool aligned);

// entry points implemented one per translation unit so the model instantiations compile in parallel
void launch_eval(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* in, const float* out, float* rgb, size_t n);
void launch_pdf(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* in, const float* out, float* pdf, size_t n);
void launch_reflectance(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* out, float* rgb, size_t n);
void launch_sample(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* out, const float* xi, float* dir, float* pdf, int32_t* flag, size_t n);
void launch_sample_eval_pdf(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* out, const float* xi,
                            float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf, size_t n, const GenArgs& g = GenArgs());
// the fused pass writing the value before the leading RGB scale (one plane) instead of rgb; false unless the BSDF is a single
// hand-merged lobe (launch_sample_eval_pdf_gray_capable says so without launching)
bool launch_sample_eval_pdf_gray_capable(const BsdfDesc&);
bool launch_sample_eval_pdf_gray(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, const float* out, const float* xi,
                                 float* dir, float* spdf, int32_t* flag, float* gray, float* pdf, size_t n);
// eval over bins [first, first + n) of the MERL grid, directions generated in the kernel; in / out may be null
void launch_eval_grid(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, int component, uint32_t first, float* rgb, float* in, float* out, size_t n);
// the device's separable merl_linearizer table (900 floats, built on first use, lives for the process)
const float* merl_lin_table_device(int device);

} // namespace bbmcu
