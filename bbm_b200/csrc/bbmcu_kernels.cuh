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
template<class...> struct VoidOf { using type = void; };

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
// unconditional 16-byte accesses (the caller has established alignment and a full group)
BBMCU_D Lanes load4_fast(const float* p)
{
  Lanes r;
#ifdef __CUDA_ARCH__
  const float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
#else
  for(int k=0; k < kVec; ++k) r.v[k] = p[k];
#endif
  return r;
}
BBMCU_D void store4_fast(float* p, const Lanes& r)
{
#ifdef __CUDA_ARCH__
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
#else
  for(int k=0; k < kVec; ++k) p[k] = r.v[k];
#endif
}
struct Lanes3 { Lanes x, y, z; BBMCU_D f3 get(int k) const { return make_f3(x.v[k], y.v[k], z.v[k]); } BBMCU_D void set(int k, f3 a) { x.v[k] = a.x; y.v[k] = a.y; z.v[k] = a.z; } };
BBMCU_D Lanes3 load4x3(const float* p, size_t i, size_t n, size_t ld, bool aligned) { Lanes3 r; r.x = load4(p, i, n, aligned); r.y = load4(p + ld, i, n, aligned); r.z = load4(p + 2*ld, i, n, aligned); return r; }
BBMCU_D void store4x3(float* p, size_t i, size_t n, size_t ld, bool aligned, const Lanes3& r) { store4(p, i, n, aligned, r.x); store4(p + ld, i, n, aligned, r.y); store4(p + 2*ld, i, n, aligned, r.z); }

// ---- counter-based inputs: Philox4x32-10 (Salmon et al. 2011) ---------------------------------------------------------
// One call per element index gives four 32-bit words: the outgoing direction (uniform on the upper hemisphere, as
// bin/checkBsdf.cpp:38-45 sampleHemisphere draws it: cos theta = u0, phi = 2 pi u1) and the two random numbers xi of
// sample().  Every kernel that takes "generated" inputs calls this one function, so a batch is a pure function of
// (seed, first index): the oracle side of the parity tests asks the library to write the inputs out and reads those.
struct GenInputs { f3 out; f2 xi; };
BBMCU_D uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
BBMCU_D void philox4x32_10(uint64_t counter, uint64_t seed, uint32_t (&r)[4])
{
  uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = 0x9E3779B9u, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for(int round=0; round < 10; ++round)
  {
    const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0, h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}
BBMCU_D float u01(uint32_t u) { return (float)(u >> 8) * 0x1p-24f; }          // [0, 1), exact in float
BBMCU_D GenInputs generate_inputs(uint64_t seed, uint64_t index)
{
  uint32_t r[4];
  philox4x32_10(index, seed, r);
  GenInputs g;
  const float z = u01(r[0]), ph = u01(r[1]) * kTwoPi;
  const float st = sqrtf(fmaxf(1.0f - z*z, 0.0f));
  float sp, cp; glibc_sincosf_both(ph, sp, cp);
  g.out = make_f3(st*cp, st*sp, z);
  g.xi = make_f2(u01(r[2]), u01(r[3]));
  return g;
}

// ---- operators --------------------------------------------------------------------------------------
template<class B> struct EvalOp
{
  using BsdfT = B;
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = false;     // eval never reads the sampling tables
  BsdfDesc bsdf; int component; const float* in; const float* out; float* rgb; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 a = load4x3(in, i, n, ld, aligned), b = load4x3(out, i, n, ld, aligned), r;
#pragma unroll
    for(int k=0; k < kVec; ++k) { Spec<float> s = B::eval(bsdf, a.get(k), b.get(k), component); r.set(k, make_f3(s.r, s.g, s.b)); }
    store4x3(rgb, i, n, ld, aligned, r);
  }
  // the same with the model's parameter-only factors formed once per thread by the kernel (BsdfSingle<M>::precompute)
  template<class PRE> BBMCU_D void group_pre(size_t i, const BsdfDesc& bsdf, const PRE& q) const
  {
    Lanes3 a = load4x3(in, i, n, ld, aligned), b = load4x3(out, i, n, ld, aligned), r;
#pragma unroll
    for(int k=0; k < kVec; ++k) { Spec<float> s = B::eval_pre(bsdf, q, a.get(k), b.get(k), component); r.set(k, make_f3(s.r, s.g, s.b)); }
    store4x3(rgb, i, n, ld, aligned, r);
  }
};

template<class B> struct PdfOp
{
  using BsdfT = B;
  static constexpr bool kOneWaveWithTables = true;            // cheap body: pay the CDF prologue once per SM slot (bbmcu_launch.cuh)
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = B::kTables;
  BsdfDesc bsdf; int component; const float* in; const float* out; float* pdf; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 a = load4x3(in, i, n, ld, aligned), b = load4x3(out, i, n, ld, aligned); Lanes r;
#pragma unroll
    for(int k=0; k < kVec; ++k) r.v[k] = B::pdf(bsdf, a.get(k), b.get(k), component);
    store4(pdf, i, n, aligned, r);
  }
};

template<class B> struct ReflectanceOp
{
  using BsdfT = B;
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = false;
  BsdfDesc bsdf; int component; const float* out; float* rgb; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 b = load4x3(out, i, n, ld, aligned), r;
#pragma unroll
    for(int k=0; k < kVec; ++k) { Spec<float> s = B::reflectance(bsdf, b.get(k), component); r.set(k, make_f3(s.r, s.g, s.b)); }
    store4x3(rgb, i, n, ld, aligned, r);
  }
};

template<class B> struct SampleOp
{
  using BsdfT = B;
  static constexpr bool kOneWaveWithTables = true;
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = B::kTables;
  BsdfDesc bsdf; int component; const float* out; const float* xi; float* dir; float* pdf; int32_t* flag; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 b = load4x3(out, i, n, ld, aligned), d; Lanes u = load4(xi, i, n, aligned), v = load4(xi + ld, i, n, aligned), p; int f[kVec];
#pragma unroll
    for(int k=0; k < kVec; ++k) { f3 dd; B::sample(bsdf, b.get(k), make_f2(u.v[k], v.v[k]), component, dd, p.v[k], f[k]); d.set(k, dd); }
    store4x3(dir, i, n, ld, aligned, d); store4(pdf, i, n, aligned, p); store4i(flag, i, n, aligned, f);
  }
};

// s = sample(out, xi); rgb = eval(s.direction, out); pdf = pdf(s.direction, out)   (20 B in, 36 B out per element)
// Any OUTPUT pointer may be null: that plane is neither stored nor (host path) copied back.  Inputs are read from
// `out` / `xi`, or - gen != 0 - drawn per element from generate_inputs(gen_seed, gen_first + i) (0 B in); the generated
// inputs are written to `gen_out` / `gen_xi` when those are given.
template<class B, bool GEN> struct SampleEvalPdfOpT
{
  using BsdfT = B;
  static constexpr bool kOneWaveWithTables = false;           // dominated by the model's eval, whose cost varies per element: keep many blocks for balance
  static constexpr int kBlock = 256, kMinBlocks = (!GEN && B::kHandFused) ? 4 : B::kMinBlocksFused;     // (the hand-merged kernel with its fast path: 4 resident blocks = 64 registers, session 33; before it: 256 x 3 with 4736 blocks 94.6 G pairs/s at 62 registers, 256 x 3 with 4736 blocks 94.6 G pairs/s, 256 x 4 94.1, 512 x 2 92.9, 128 x 8 93.1, 1024 x 1 91.3)
  static constexpr bool kHasBsdf = true, kTables = B::kTables;
  BsdfDesc bsdf; int component; const float* out; const float* xi; float* dir; float* spdf; int32_t* flag; float* rgb; float* pdf; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  uint64_t gen_seed = 0, gen_first = 0; float* gen_out = nullptr; float* gen_xi = nullptr;       // GEN only
  // FAST: every plane of the call is 16-byte aligned, every output is wanted and the group is a full one - the kernel tests
  // that once per launch (fast_ok) and runs the groups below n / 4 without per-plane tests and fall-back paths (the
  // headline kernel spent ~25 of its 341 instructions per pair on them)
  static constexpr bool kFastPath = !GEN && B::kHandFused;
  BBMCU_D bool fast_ok() const { return aligned && dir && spdf && flag && rgb && pdf; }
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const { group_t<false>(i, bsdf); }
  template<bool FAST> BBMCU_D void group_t(size_t i, const BsdfDesc& bsdf) const { group_q<FAST, int>(i, bsdf, nullptr); }
  // q != nullptr: the model's parameter-only factors, formed once per thread by the kernel (BsdfSingle<M>::precompute)
  template<class PRE> BBMCU_D void group_pre(size_t i, const BsdfDesc& bsdf, const PRE& q) const { group_q<false, PRE>(i, bsdf, &q); }
  template<bool FAST, class PRE> BBMCU_D void group_q(size_t i, const BsdfDesc& bsdf, const PRE* q) const
  {
    Lanes3 b, d, c; Lanes u, v, sp, p; int f[kVec];
    if constexpr (GEN)
    {
#pragma unroll
      for(int k=0; k < kVec; ++k) { GenInputs g = generate_inputs(gen_seed, gen_first + i + k); b.set(k, g.out); u.v[k] = g.xi.x; v.v[k] = g.xi.y; }
      if(gen_out) store4x3(gen_out, i, n, ld, aligned, b);
      if(gen_xi) { store4(gen_xi, i, n, aligned, u); store4(gen_xi + ld, i, n, aligned, v); }
    }
    else if constexpr (FAST) { b.x = load4_fast(out + i); b.y = load4_fast(out + ld + i); b.z = load4_fast(out + 2*ld + i); u = load4_fast(xi + i); v = load4_fast(xi + ld + i); }
    else { b = load4x3(out, i, n, ld, aligned); u = load4(xi, i, n, aligned); v = load4(xi + ld, i, n, aligned); }
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
        if constexpr (!std::is_same<PRE, int>::value) B::eval_pdf_pre(bsdf, *q, dd, o, component, s, p.v[k]);
        else B::eval_pdf(bsdf, dd, o, component, s, p.v[k]);
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
    if constexpr (FAST)
    {
      store4_fast(dir + i, d.x); store4_fast(dir + ld + i, d.y); store4_fast(dir + 2*ld + i, d.z);
      store4_fast(spdf + i, sp);
      *reinterpret_cast<int4*>(flag + i) = make_int4(f[0], f[1], f[2], f[3]);
      store4_fast(rgb + i, c.x); store4_fast(rgb + ld + i, c.y); store4_fast(rgb + 2*ld + i, c.z);
      store4_fast(pdf + i, p);
      return;
    }
    if(dir) store4x3(dir, i, n, ld, aligned, d);
    if(spdf) store4(spdf, i, n, aligned, sp);
    if(flag) store4i(flag, i, n, aligned, f);
    if(rgb) store4x3(rgb, i, n, ld, aligned, c);
    if(pdf) store4(pdf, i, n, aligned, p);
  }
};

// The fused pass of a hand-merged single-lobe model with the value BEFORE its leading RGB scale: `gray` receives u (one
// plane), eval = u * scale.  Used by the host-pointer path only (bbmcu_api.cu): 4 bytes per element cross the link instead
// of 12 and host threads form the three products (same IEEE single multiplications as the device: same bits).
template<class B> struct SampleEvalPdfGrayOp
{
  using BsdfT = B;
  static constexpr bool kOneWaveWithTables = false;
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocksFused;
  static constexpr bool kHasBsdf = true, kTables = B::kTables;
  BsdfDesc bsdf; int component; const float* out; const float* xi; float* dir; float* spdf; int32_t* flag; float* gray; float* pdf; size_t n; bool aligned; size_t ld = 0;
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf) const
  {
    Lanes3 b = load4x3(out, i, n, ld, aligned), d; Lanes u = load4(xi, i, n, aligned), v = load4(xi + ld, i, n, aligned), g, p; int f[kVec];
#pragma unroll
    for(int k=0; k < kVec; ++k) { f3 dd; B::sample_eval_pdf_merged_u(bsdf, b.get(k), make_f2(u.v[k], v.v[k]), component, dd, f[k], g.v[k], p.v[k]); d.set(k, dd); }
    if(dir) store4x3(dir, i, n, ld, aligned, d);
    if(spdf) store4(spdf, i, n, aligned, p);                  // p is already 0 when the sample is invalid
    if(flag) store4i(flag, i, n, aligned, f);
    if(gray) store4(gray, i, n, aligned, g);
    if(pdf) store4(pdf, i, n, aligned, p);
  }
};

template<class B> using SampleEvalPdfOp = SampleEvalPdfOpT<B, false>;
template<class B> using SampleEvalPdfGenOp = SampleEvalPdfOpT<B, true>;      // inputs drawn in the kernel (0 B in)

// eval over the MERL grid with the linearizer fused: element i is bin first + i, its direction pair comes from the
// separable table in shared memory (bit-identical to MerlDirsOp), 0 B in + 12 B out per eval.
// `in` / `out` (optional) receive the generated directions - the parity protocol of SURVEY.md section 7.
template<class B> struct EvalGridOp
{
  using BsdfT = B;
  static constexpr int kBlock = 256, kMinBlocks = B::kMinBlocks;
  static constexpr bool kHasBsdf = true, kTables = false, kLinTab = true;
  BsdfDesc bsdf; int component; const float* lin_tab; uint32_t first; float* rgb; float* in; float* out; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i, const BsdfDesc& bsdf, const float* s_lin) const
  {
    Lanes3 a, b, r;
#pragma unroll
    for(int k=0; k < kVec; ++k)
    {
      f3 x, y;
      merl_dirs_tab(s_lin, (i + k < n) ? first + (uint32_t)(i + k) : first, x, y);
      Spec<float> s = B::eval(bsdf, x, y, component);
      a.set(k, x); b.set(k, y); r.set(k, make_f3(s.r, s.g, s.b));
    }
    store4x3(rgb, i, n, ld, aligned, r);
    if(in) store4x3(in, i, n, ld, aligned, a);
    if(out) store4x3(out, i, n, ld, aligned, b);
  }
  template<class PRE> BBMCU_D void group_pre(size_t i, const BsdfDesc& bsdf, const float* s_lin, const PRE& q) const
  {
    Lanes3 a, b, r;
#pragma unroll
    for(int k=0; k < kVec; ++k)
    {
      f3 x, y;
      merl_dirs_tab(s_lin, (i + k < n) ? first + (uint32_t)(i + k) : first, x, y);
      Spec<float> s = B::eval_pre(bsdf, q, x, y, component);
      a.set(k, x); b.set(k, y); r.set(k, make_f3(s.r, s.g, s.b));
    }
    store4x3(rgb, i, n, ld, aligned, r);
    if(in) store4x3(in, i, n, ld, aligned, a);
    if(out) store4x3(out, i, n, ld, aligned, b);
  }
};

struct MerlIndexOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  const float* in; const float* out; uint32_t* index; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a = load4x3(in, i, n, ld, aligned), b = load4x3(out, i, n, ld, aligned); int r[kVec];
#pragma unroll
    for(int k=0; k < kVec; ++k) r[k] = (int)merl_index(a.get(k), b.get(k));
    store4i(reinterpret_cast<int32_t*>(index), i, n, aligned, r);
  }
};

struct MerlDirsOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  uint32_t first; float* in; float* out; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a, b;
#pragma unroll
    for(int k=0; k < kVec; ++k) { f3 x, y; merl_dirs(first + (uint32_t)(i + k), x, y); a.set(k, x); b.set(k, y); }
    store4x3(in, i, n, ld, aligned, a); store4x3(out, i, n, ld, aligned, b);
  }
};

struct SphericalDirsOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  SphericalGrid grid; uint64_t first; float* in; float* out; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a, b;
#pragma unroll
    for(int k=0; k < kVec; ++k) { f3 x, y; spherical_dirs(grid, first + i + k, x, y); a.set(k, x); b.set(k, y); }
    store4x3(in, i, n, ld, aligned, a); store4x3(out, i, n, ld, aligned, b);
  }
};

// gather the measured grid at the bin of each direction pair: merl_data::eval (staticmodel/merl.h:78-96)
struct MerlLookupOp
{
  static constexpr int kBlock = 256, kMinBlocks = 1;
  static constexpr bool kHasBsdf = false, kTables = false;
  const float* table; const float* in; const float* out; float* rgb; uint32_t* bad; size_t n; bool aligned; size_t ld = 0;   // ld (0 = n): floats between the planes of one SoA argument (>= n)
  BBMCU_D void group(size_t i) const
  {
    Lanes3 a = load4x3(in, i, n, ld, aligned), b = load4x3(out, i, n, ld, aligned), r;
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
    store4x3(rgb, i, n, ld, aligned, r);
  }
};

template<class Op, class = void> struct OpUsesEpd { static constexpr bool value = false; };
template<class Op> struct OpUsesEpd<Op, typename VoidOf<typename Op::BsdfT>::type> { static constexpr bool value = UsesEpd<typename Op::BsdfT>::value; };
template<class Op, class = void> struct OpFastPath { static constexpr bool value = false; };
template<class Op> struct OpFastPath<Op, typename std::enable_if<Op::kFastPath>::type> { static constexpr bool value = true; };
template<class Op, class = void> struct OpHasPre { static constexpr bool value = false; };
template<class B> struct OpHasPre<EvalOp<B>, typename std::enable_if<B::kHasPre>::type> { static constexpr bool value = true; };
template<class B> struct OpHasPre<EvalGridOp<B>, typename std::enable_if<B::kHasPre>::type> { static constexpr bool value = true; };
template<class B, bool GEN> struct OpHasPre<SampleEvalPdfOpT<B, GEN>, typename std::enable_if<B::kHasPre && B::kFusedSample && !B::kHandFused>::type> { static constexpr bool value = true; };
template<class Op, class = void> struct UsesLinTab { static constexpr bool value = false; };
template<class Op> struct UsesLinTab<Op, typename std::enable_if<Op::kLinTab>::type> { static constexpr bool value = true; };

#ifdef __CUDACC__
// the separable merl_linearizer table of this device (bbmcu_linearizer.cuh), built once by k_merl_lin_tab
__global__ void k_merl_lin_tab(float* tab);

template<class Op> __global__ void __launch_bounds__(Op::kBlock, Op::kMinBlocks) k_foreach4(const Op op, size_t groups)
{
  const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  if constexpr (OpUsesEpd<Op>::value) { epd_stage_rows(op.bsdf, threadIdx.x, blockDim.x); __syncthreads(); }       // G1 rows of this launch's p -> shared memory
  if constexpr (!Op::kHasBsdf) { for(size_t g = first; g < groups; g += stride) op.group(g * kVec); }
  else if constexpr (UsesLinTab<Op>::value)
  {
    __shared__ float s_lin[kMerlLinTabFloats];
    for(int i = threadIdx.x; i < kMerlLinTabFloats; i += blockDim.x) s_lin[i] = __ldg(op.lin_tab + i);
    __syncthreads();
    if constexpr (OpHasPre<Op>::value)
    {
      const auto q = Op::BsdfT::precompute(op.bsdf);
      for(size_t g = first; g < groups; g += stride) op.group_pre(g * kVec, op.bsdf, s_lin, q);
    }
    else
    for(size_t g = first; g < groups; g += stride) op.group(g * kVec, op.bsdf, s_lin);
  }
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
    if constexpr (OpFastPath<Op>::value)
    {
      if(op.fast_ok())                           // uniform: full groups without per-plane tests, the ragged last group (if any) on the general path
      {
        const size_t full = op.n / kVec;
        for(size_t g = first; g < full; g += stride) op.template group_t<true>(g * kVec, op.bsdf);
        for(size_t g = full + first; g < groups; g += stride) op.group(g * kVec, op.bsdf);
        return;
      }
    }
    if constexpr (OpHasPre<Op>::value)
    {
      const auto q = Op::BsdfT::precompute(op.bsdf);     // parameter-only factors of the model, once per thread
      for(size_t g = first; g < groups; g += stride) op.group_pre(g * kVec, op.bsdf, q);
      return;
    }
    for(size_t g = first; g < groups; g += stride) op.group(g * kVec, op.bsdf);
  }
}
#endif

} // namespace bbmcu
