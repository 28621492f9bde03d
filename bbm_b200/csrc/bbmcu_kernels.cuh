// Element-wise operators over struct-of-arrays batches and the kernel that drives them.
//
// Every operator is a small functor with a __host__ __device__ `item(i)` so the same code can be
// compiled for the host-side unit tests (tests/hostsim, test infrastructure only - the product
// always launches the kernels below).  The kernel gives each thread 4 consecutive elements:
// six (eval) 16-byte coalesced loads in flight per thread before any arithmetic, 16-byte stores,
// grid sized in whole waves of the 148 SMs.
//
// These replace the scalar loops of the reference's tools and losses
// (bin/checkBsdf.cpp:79-87,123-141,206-267; include/bbm/sampledlossfunction.h:78-87).
#pragma once
#include "bbmcu_bsdf.cuh"
#include "bbmcu_linearizer.cuh"
#include <cstddef>

namespace bbmcu {

constexpr int kVec = 4;

// ---- SoA access: 4 consecutive elements of one plane ---------------------------------------------
struct Lanes { float v[kVec]; };

BBMCU_D Lanes load4(const float* p, size_t i, size_t n, bool aligned)
{
  Lanes r;
#ifdef __CUDA_ARCH__
  if(aligned && i + kVec <= n) { float4 t = __ldg(reinterpret_cast<const float4*>(p + i)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
#endif
#pragma unroll
  for(int k=0; k < kVec; ++k) r.v[k] = (i + k < n) ? p[i + k] : 0.0f;
  return r;
}
BBMCU_D void store4(float* p, size_t i, size_t n, bool aligned, const Lanes& r)
{
#ifdef __CUDA_ARCH__
  if(aligned && i + kVec <= n) { *reinterpret_cast<float4*>(p + i) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]); return; }
#endif
#pragma unroll
  for(int k=0; k < kVec; ++k) if(i + k < n) p[i + k] = r.v[k];
}
BBMCU_D void store4i(int32_t* p, size_t i, size_t n, bool aligned, const int (&r)[kVec])
{
#ifdef __CUDA_ARCH__
  if(aligned && i + kVec <= n) { *reinterpret_cast<int4*>(p + i) = make_int4(r[0], r[1], r[2], r[3]); return; }
#endif
#pragma unroll
  for(int k=0; k < kVec; ++k) if(i + k < n) p[i + k] = r[k];
}
struct Lanes3 { Lanes x, y, z; BBMCU_D f3 get(int k) const { return make_f3(x.v[k], y.v[k], z.v[k]); } BBMCU_D void set(int k, f3 a) { x.v[k] = a.x; y.v[k] = a.y; z.v[k] = a.z; } };
BBMCU_D Lanes3 load4x3(const float* p, size_t i, size_t n, bool aligned) { Lanes3 r; r.x = load4(p, i, n, aligned); r.y = load4(p + n, i, n, aligned); r.z = load4(p + 2*n, i, n, aligned); return r; }
BBMCU_D void store4x3(float* p, size_t i, size_t n, bool aligned, const Lanes3& r) { store4(p, i, n, aligned, r.x); store4(p + n, i, n, aligned, r.y); store4(p + 2*n, i, n, aligned, r.z); }

// ---- operators --------------------------------------------------------------------------------------
template<class B> struct EvalOp
{
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = false;     // eval never reads the sampling tables
  BsdfDesc bsdf; int component; const float* in; const float* out; float* rgb; size_t n; bool aligned;
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 a = load4x3(in, i, n, aligned), b = load4x3(out, i, n, aligned), r;
#pragma unroll
    for(int k=0; k < kVec; ++k) { Spec<float> s = B::eval(bsdf, a.get(k), b.get(k), component); r.set(k, make_f3(s.r, s.g, s.b)); }
    store4x3(rgb, i, n, aligned, r);
  }
};

template<class B> struct PdfOp
{
  static constexpr bool kOneWaveWithTables = true;            // cheap body: pay the CDF prologue once per SM slot (bbmcu_launch.cuh)
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = B::kTables;
  BsdfDesc bsdf; int component; const float* in; const float* out; float* pdf; size_t n; bool aligned;
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 a = load4x3(in, i, n, aligned), b = load4x3(out, i, n, aligned); Lanes r;
#pragma unroll
    for(int k=0; k < kVec; ++k) r.v[k] = B::pdf(bsdf, a.get(k), b.get(k), component);
    store4(pdf, i, n, aligned, r);
  }
};

template<class B> struct ReflectanceOp
{
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = false;
  BsdfDesc bsdf; int component; const float* out; float* rgb; size_t n; bool aligned;
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 b = load4x3(out, i, n, aligned), r;
#pragma unroll
    for(int k=0; k < kVec; ++k) { Spec<float> s = B::reflectance(bsdf, b.get(k), component); r.set(k, make_f3(s.r, s.g, s.b)); }
    store4x3(rgb, i, n, aligned, r);
  }
};

template<class B> struct SampleOp
{
  static constexpr bool kOneWaveWithTables = true;
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = B::kTables;
  BsdfDesc bsdf; int component; const float* out; const float* xi; float* dir; float* pdf; int32_t* flag; size_t n; bool aligned;
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 b = load4x3(out, i, n, aligned), d; Lanes u = load4(xi, i, n, aligned), v = load4(xi + n, i, n, aligned), p; int f[kVec];
#pragma unroll
    for(int k=0; k < kVec; ++k) { f3 dd; B::sample(bsdf, b.get(k), make_f2(u.v[k], v.v[k]), component, dd, p.v[k], f[k]); d.set(k, dd); }
    store4x3(dir, i, n, aligned, d); store4(pdf, i, n, aligned, p); store4i(flag, i, n, aligned, f);
  }
};

// s = sample(out, xi); rgb = eval(s.direction, out); pdf = pdf(s.direction, out)   (20 B in, 36 B out per element)
template<class B> struct SampleEvalPdfOp
{
  static constexpr bool kOneWaveWithTables = false;           // dominated by the model's eval, whose cost varies per element: keep many blocks for balance
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocksFused;     // (the hand-merged GGX kernel: 256 x 3 with 4736 blocks 94.6 G pairs/s, 256 x 4 94.1, 512 x 2 92.9, 128 x 8 93.1, 1024 x 1 91.3)
  static constexpr bool kHasBsdf = true, kTables = B::kTables;
  BsdfDesc bsdf; int component; const float* out; const float* xi; float* dir; float* spdf; int32_t* flag; float* rgb; float* pdf; size_t n; bool aligned;
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 b = load4x3(out, i, n, aligned), d, c; Lanes u = load4(xi, i, n, aligned), v = load4(xi + n, i, n, aligned), sp, p; int f[kVec];
#pragma unroll
    for(int k=0; k < kVec; ++k)
    {
      f3 o = b.get(k), dd;
      Spec<float> s;
      if constexpr (B::kHandFused)
      {
        B::sample_eval_pdf_merged(bsdf, o, make_f2(u.v[k], v.v[k]), component, dd, f[k], s, p.v[k]);
        sp.v[k] = p.v[k];                                     // p is already 0 when the sample is invalid
      }
      else if constexpr (B::kFusedSample)
      {
        B::sample_dir(bsdf, o, make_f2(u.v[k], v.v[k]), component, dd, f[k]);
        B::eval_pdf(bsdf, dd, o, component, s, p.v[k]);
        sp.v[k] = (f[k] != FLAG_NONE) ? p.v[k] : 0.0f;
      }
      else
      {
        B::sample(bsdf, o, make_f2(u.v[k], v.v[k]), component, dd, sp.v[k], f[k]);
        // an aggregate's sample.pdf IS pdf(sample.direction, out) - the same weights and lobe pdfs in the same order
        // (aggregatebsdf.h:119-126 and :183) - so the fused pass does not evaluate it twice
        if(B::kAggregatePdfFromSample && bsdf.aggregate) p.v[k] = sp.v[k];
        else p.v[k] = B::pdf(bsdf, dd, o, component);
        s = B::eval(bsdf, dd, o, component);
      }
      d.set(k, dd); c.set(k, make_f3(s.r, s.g, s.b));
    }
    store4x3(dir, i, n, aligned, d); store4(spdf, i, n, aligned, sp); store4i(flag, i, n, aligned, f);
    store4x3(rgb, i, n, aligned, c); store4(pdf, i, n, aligned, p);
  }
};

struct MerlIndexOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  const float* in; const float* out; uint32_t* index; size_t n; bool aligned;
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a = load4x3(in, i, n, aligned), b = load4x3(out, i, n, aligned); int r[kVec];
#pragma unroll
    for(int k=0; k < kVec; ++k) r[k] = (int)merl_index(a.get(k), b.get(k));
    store4i(reinterpret_cast<int32_t*>(index), i, n, aligned, r);
  }
};

struct MerlDirsOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  uint32_t first; float* in; float* out; size_t n; bool aligned;
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a, b;
#pragma unroll
    for(int k=0; k < kVec; ++k) { f3 x, y; merl_dirs(first + (uint32_t)(i + k), x, y); a.set(k, x); b.set(k, y); }
    store4x3(in, i, n, aligned, a); store4x3(out, i, n, aligned, b);
  }
};

struct SphericalDirsOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  SphericalGrid grid; uint64_t first; float* in; float* out; size_t n; bool aligned;
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a, b;
#pragma unroll
    for(int k=0; k < kVec; ++k) { f3 x, y; spherical_dirs(grid, first + i + k, x, y); a.set(k, x); b.set(k, y); }
    store4x3(in, i, n, aligned, a); store4x3(out, i, n, aligned, b);
  }
};

// gather the measured grid at the bin of each direction pair: merl_data::eval (staticmodel/merl.h:78-96)
struct MerlLookupOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  const float* table; const float* in; const float* out; float* rgb; uint32_t* bad; size_t n; bool aligned;
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a = load4x3(in, i, n, aligned), b = load4x3(out, i, n, aligned), r;
#pragma unroll
    for(int k=0; k < kVec; ++k)
    {
      f3 x = a.get(k), y = b.get(k), s = make_f3(0, 0, 0);
      if((x.z >= 0.0f) && (y.z >= 0.0f) && (i + k < n))
      {
        uint32_t idx = merl_index(x, y);
        if(idx < kMerlBins) s = make_f3(table[idx], table[kMerlBins + idx], table[2*kMerlBins + idx]);
        else if(bad) {                       // the reference throws here (backbone/native control.h:75)
#ifdef __CUDA_ARCH__
          atomicAdd(bad, 1u);
#else
          ++*bad;
#endif
        }
      }
      r.set(k, s);
    }
    store4x3(rgb, i, n, aligned, r);
  }
};

#ifdef __CUDACC__
template<class Op> __global__ void __launch_bounds__(Op::kBlock, Op::kMinBlocks) k_foreach4(const Op op, size_t groups)
{
  const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if constexpr (!Op::kHasBsdf) { for(size_t g = first; g < groups; g += stride) op.group(g * kVec); }
  else
  {
    if constexpr (Op::kTables)
    {
      if(op.bsdf.n_tables)                       // uniform: a He-family lobe - build its CDF in shared memory first
      {
        __shared__ BsdfDesc sb;
        const int words = (int)((offsetof(BsdfDesc, attrs) + sizeof(float)*op.bsdf.n_floats + 3) / 4);
        for(int i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(&sb)[i] = reinterpret_cast<const uint32_t*>(&op.bsdf)[i];
        __syncthreads();
        bsdf_tables_phase1(sb, op.component, threadIdx.x, blockDim.x);
        __syncthreads();
        bsdf_tables_phase2(sb, threadIdx.x);
        __syncthreads();
        for(size_t g = first; g < groups; g += stride) op.group(g * kVec, sb);
        return;
      }
    }
    for(size_t g = first; g < groups; g += stride) op.group(g * kVec, op.bsdf);
  }
}
#endif

} // namespace bbmcu
