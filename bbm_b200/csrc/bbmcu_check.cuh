// checkBsdf on the device: the six consistency tests of the reference's bin/checkBsdf.cpp as fused
// generate -> evaluate -> reduce kernels.  The reference loops over samples on one thread, calling
// bsdf_ptr::eval / sample / pdf per sample and accumulating in float; here one item = one sample of that
// loop, its random numbers come from a counter-based generator (Philox, 0 bytes in) or - to reproduce the
// reference's printed numbers - from the reference's own std::mt19937 stream drawn on the host, and the
// per-sample terms are reduced on the device (float terms, double sums, one partial row per block added in
// fixed order by the host).
//
//   CHECK_REFLECTANCE   bin/checkBsdf.cpp:51-97    sum of eval(dir, out) * cos / pdf over sphere or BSDF samples
//   CHECK_RECIPROCITY   bin/checkBsdf.cpp:102-152  |eval(a, b) - eval(b, a)|: mean, max and where (adjoint, :157-201, is
//                                                   the same numbers: no model of the reference looks at unit_t)
//   CHECK_PDF           bin/checkBsdf.cpp:206-267  negative pdfs, directions below the horizon, |sample.pdf - pdf|
//   CHECK_PDF_INTEGRAL  bin/checkBsdf.cpp:272-316  MC integral of pdf(., out) over the sphere
//   CHECK_BIN_PDF       bin/checkBsdf.cpp:352-371  MC integral of pdf(., out) over each (theta, phi) bin
//   CHECK_BIN_COUNT     bin/checkBsdf.cpp:373-388  histogram of sampled directions over the same bins
#pragma once
#include <cstring>
#include "bbmcu_kernels.cuh"

namespace bbmcu {

enum : int { CHECK_REFLECTANCE = 0, CHECK_RECIPROCITY, CHECK_PDF, CHECK_PDF_INTEGRAL, CHECK_BIN_PDF, CHECK_BIN_COUNT };
constexpr int kCheckCols = 6;
constexpr int kCheckThreads = 256;
constexpr int kCheckSharedBins = 4096;       // histograms up to this many bins are privatised per block in shared memory
constexpr int kCheckOffenderFloats = 8;      // kind, pdf, sampled direction (3), view direction (3)

struct CheckArgs
{
  int mode, component;
  uint64_t seed, first;                // Philox: item i draws from counters 2 (first + i) and 2 (first + i) + 1 of `seed`
  unsigned long long n;                // items of this launch (CHECK_BIN_PDF: per bin)
  const float* xi; int xi_per_item;    // a given stream instead: item i reads xi[i * xi_per_item + j]
  float ox, oy, oz;                    // the fixed direction of the test (out / trial direction)
  int importance, sphere, include_zero, below_horizon;
  int n_theta, n_phi;
  double* partial;                     // [blocks][kCheckCols]   (CHECK_BIN_PDF: [bins][gridDim.x])
  unsigned long long* maxkey;          // [blocks]                (CHECK_RECIPROCITY)
  unsigned int* counts;                // [n_theta * n_phi]       (CHECK_BIN_COUNT)
  float* detail;                       // 9 floats: in, out, diff of item 0 (CHECK_RECIPROCITY with n == 1)
  unsigned int* n_offenders; float* offenders; int max_offenders;     // CHECK_PDF: the first few offending samples
};

struct CheckU { float u[8]; };
BBMCU_D CheckU check_draw(const CheckArgs& a, unsigned long long i, int need)
{
  CheckU r;
  if(a.xi)
  {
#pragma unroll
    for(int j=0; j < 8; ++j) r.u[j] = (j < a.xi_per_item) ? a.xi[i * (unsigned long long)a.xi_per_item + j] : 0.0f;
    return r;
  }
  uint32_t w[4];
  philox4x32_10(2*(a.first + i), a.seed, w);
#pragma unroll
  for(int j=0; j < 4; ++j) r.u[j] = u01(w[j]);
  if(need > 4)
  {
    philox4x32_10(2*(a.first + i) + 1, a.seed, w);
#pragma unroll
    for(int j=0; j < 4; ++j) r.u[4 + j] = u01(w[j]);
  }
  else { r.u[4] = r.u[5] = r.u[6] = r.u[7] = 0.0f; }
  return r;
}

// spherical::convert(vec2d) (core/spherical.h:58-65) with the host libm's sinf / cosf
BBMCU_D f3 check_to_vec(float phi, float theta)
{
  float st, ct, sp, cp;
  glibc_sincosf_both(theta, st, ct); glibc_sincosf_both(phi, sp, cp);
  return make_f3(cp*st, sp*st, ct);
}
// sampleSphere / sampleHemisphere (bin/checkBsdf.cpp:27-45): x0 -> theta, x1 -> phi; pdf 1 / (4 pi) and 1 / (2 pi)
BBMCU_D f3 check_sample_sphere(float x0, float x1)
{
  double c = 1.0 - 2.0*(double)x0;
  c = fmin(1.0, fmax(-1.0, c));
  return check_to_vec(x1 * kTwoPi, (float)acos(c));
}
BBMCU_D f3 check_sample_hemisphere(float x0, float x1) { return check_to_vec(x1 * kTwoPi, acosf(fminf(1.0f, fmaxf(-1.0f, x0)))); }
BBMCU_D float check_sphere_pdf() { return (float)(1.0 / (double)(float)(4.0f * 3.14159265358979323846)); }

// one item of a test: adds its terms to acc[], may update the running maximum (key) and the histogram
template<class B>
BBMCU_D void check_item(const CheckArgs& a, const BsdfDesc& bsdf, unsigned long long i, unsigned long long bin, double (&acc)[kCheckCols],
                        unsigned long long& key, unsigned int* hist)
{
  const f3 fixed = make_f3(a.ox, a.oy, a.oz);
  if(a.mode == CHECK_REFLECTANCE)
  {
    if constexpr (B::kCheckEval)
    {
      CheckU r = check_draw(a, i, 2);
      f3 dir; float pdf; int flag = 0;
      if(a.importance)
      {
        if constexpr (B::kCheckSample) B::sample(bsdf, fixed, make_f2(r.u[0], r.u[1]), a.component, dir, pdf, flag);
        else return;
      }
      else { dir = check_sample_sphere(r.u[0], r.u[1]); pdf = check_sphere_pdf(); }
      if(pdf > kEps)
      {
        Spec<float> e = B::eval(bsdf, dir, fixed, a.component);
        acc[0] += (double)(e.r * dir.z / pdf); acc[1] += (double)(e.g * dir.z / pdf); acc[2] += (double)(e.b * dir.z / pdf);
      }
    }
  }
  else if(a.mode == CHECK_RECIPROCITY)
  {
    if constexpr (B::kCheckEval && !B::kCheckSample)
    {
      CheckU r = check_draw(a, i, 4);
      f3 in = check_sample_sphere(r.u[0], r.u[1]), out = check_sample_sphere(r.u[2], r.u[3]);
      Spec<float> e0 = B::eval(bsdf, in, out, a.component), e1 = B::eval(bsdf, out, in, a.component);
      float dr = fabsf(e0.r - e1.r), dg = fabsf(e0.g - e1.g), db = fabsf(e0.b - e1.b);
      acc[0] += (double)dr; acc[1] += (double)dg; acc[2] += (double)db;
      float h = (dr + dg) + db;
      if(h > 0.0f)      // strict '>' in the reference: the first sample that reaches the maximum keeps it (NaN never does)
      {
#ifdef __CUDA_ARCH__
        unsigned long long k = ((unsigned long long)__float_as_uint(h) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
#else
        uint32_t hb; memcpy(&hb, &h, 4);
        unsigned long long k = ((unsigned long long)hb << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
#endif
        if(k > key) key = k;
      }
      if(a.detail && i == 0)
      {
        a.detail[0] = in.x; a.detail[1] = in.y; a.detail[2] = in.z; a.detail[3] = out.x; a.detail[4] = out.y; a.detail[5] = out.z;
        a.detail[6] = dr; a.detail[7] = dg; a.detail[8] = db;
      }
    }
  }
  else if(a.mode == CHECK_PDF)
  {
    if constexpr (B::kCheckSample && !B::kCheckEval)
    {
      CheckU r = check_draw(a, i, 6);
      f3 view = a.sphere ? check_sample_sphere(r.u[0], r.u[1]) : check_sample_hemisphere(r.u[0], r.u[1]);
#pragma unroll
      for(int pass=0; pass < 2; ++pass)      // radiance, importance: separate random numbers, the same functions
      {
        f3 dir; float sp; int flag;
        B::sample(bsdf, view, make_f2(r.u[2 + 2*pass], r.u[3 + 2*pass]), a.component, dir, sp, flag);
        float p = B::pdf(bsdf, dir, view, a.component);
        const bool below = a.below_horizon && (dir.z < 0.0f), negative = p < 0.0f;
        if(below) acc[4 + pass] += 1.0;
        if(negative) acc[2 + pass] += 1.0;
        acc[pass] += (double)fabsf(sp - p);
        if((below || negative) && a.offenders)
        {
#ifdef __CUDA_ARCH__
          unsigned int slot = atomicAdd(a.n_offenders, 1u);
#else
          unsigned int slot = (*a.n_offenders)++;
#endif
          if(slot < (unsigned int)a.max_offenders)
          {
            float* o = a.offenders + (size_t)slot*kCheckOffenderFloats;
            o[0] = (float)((negative ? 2 : 0) + pass); o[1] = p; o[2] = dir.x; o[3] = dir.y; o[4] = dir.z; o[5] = view.x; o[6] = view.y; o[7] = view.z;
          }
        }
      }
    }
  }
  else if(a.mode == CHECK_PDF_INTEGRAL)
  {
    if constexpr (!B::kCheckSample && !B::kCheckEval)
    {
      CheckU r = check_draw(a, i, 2);
      f3 dir = check_sample_sphere(r.u[0], r.u[1]);
      float ps = check_sphere_pdf();
      if(ps > kEps) acc[0] += (double)(B::pdf(bsdf, dir, fixed, a.component) / ps);
    }
  }
  else if(a.mode == CHECK_BIN_PDF)
  {
    if constexpr (!B::kCheckSample && !B::kCheckEval)
    {
      // item i of bin (t, p): phi = 2 pi (p + rnd0) / n_phi, theta = pi (t + rnd1) / n_theta, weighted by the bin's solid angle
      CheckU r = check_draw(a, bin * a.n + i, 2);
      const unsigned long long t = bin / (unsigned long long)a.n_phi, p = bin % (unsigned long long)a.n_phi;
      float phi = kTwoPi * ((float)p + r.u[0]) / (float)a.n_phi;
      float theta = kPi * ((float)t + r.u[1]) / (float)a.n_theta;
      f3 dir = check_to_vec(phi, theta);
      float w = ((2.0f * kPi) * kPi) * fabsf(glibc_sinf(theta)) / (float)((unsigned long long)a.n_phi * (unsigned long long)a.n_theta);
      acc[0] += (double)(B::pdf(bsdf, dir, fixed, a.component) * w);
    }
  }
  else if(a.mode == CHECK_BIN_COUNT)
  {
    if constexpr (B::kCheckSample && !B::kCheckEval)
    {
      CheckU r = check_draw(a, i, 2);
      f3 dir; float sp; int flag;
      B::sample(bsdf, fixed, make_f2(r.u[0], r.u[1]), a.component, dir, sp, flag);
      if(a.include_zero || sp > kEps)
      {
        float ft = fminf(sph_theta(dir) / kPi * (float)a.n_theta, (float)(a.n_theta - 1));
        float fp = fminf(sph_phi(dir) / kTwoPi * (float)a.n_phi, (float)(a.n_phi - 1));
        // a NaN direction casts to an arbitrary size_t in the reference (undefined); counted in bin 0 here
        int t = (ft >= 0.0f) ? (int)ft : 0, p = (fp >= 0.0f) ? (int)fp : 0;
        unsigned int idx = (unsigned int)(t * a.n_phi + p);
#ifdef __CUDA_ARCH__
        atomicAdd(hist + idx, 1u);
#else
        ++hist[idx];
#endif
      }
    }
  }
}

// which calls of the model a kernel instance holds (keeps each of the three kernels per model small):
//   CheckE: eval only (reflectance over the sphere, reciprocity)      CheckS: sample + pdf (pdf test, bin counts)
//   CheckSE: sample + eval (reflectance with importance sampling)     CheckP: pdf only (pdf integral, bin pdf)
template<class B, bool EVAL, bool SAMPLE> struct CheckView : B { using Base = B; static constexpr bool kCheckEval = EVAL, kCheckSample = SAMPLE; };

#ifdef __CUDACC__
template<class V>
__global__ void __launch_bounds__(kCheckThreads) k_check(const CheckArgs a, const BsdfDesc desc)
{
  __shared__ double s_sum[kCheckThreads/32][kCheckCols];
  __shared__ unsigned long long s_key[kCheckThreads/32];
  __shared__ unsigned int s_hist[kCheckSharedBins];
  __shared__ BsdfDesc sb;
  const BsdfDesc* bsdf = &desc;
  if constexpr (UsesEpd<typename V::Base>::value) { epd_stage_rows(desc, threadIdx.x, blockDim.x); __syncthreads(); }
  if constexpr (V::kTables)
  {
    if(desc.n_tables)                          // He-family / measured lobe: build the sampling CDF first (as k_foreach4 does)
    {
      const int words = (int)((offsetof(BsdfDesc, attrs) + sizeof(float)*desc.n_floats + 3) / 4);
      for(int i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(&sb)[i] = reinterpret_cast<const uint32_t*>(&desc)[i];
      __syncthreads();
      bsdf_tables_phase1(sb, a.component, threadIdx.x, blockDim.x);
      __syncthreads();
      bsdf_tables_phase2(sb, threadIdx.x);
      __syncthreads();
      bsdf = &sb;
    }
  }
  const int bins = a.n_theta * a.n_phi;
  unsigned int* hist = a.counts;
  const bool shared_hist = (a.mode == CHECK_BIN_COUNT) && bins <= kCheckSharedBins;
  if(shared_hist)
  {
    for(int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    hist = s_hist;
  }
  double acc[kCheckCols];
#pragma unroll
  for(int j=0; j < kCheckCols; ++j) acc[j] = 0.0;
  unsigned long long key = 0ull;
  const unsigned long long bin = blockIdx.y;
  for(unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (unsigned long long)gridDim.x * blockDim.x)
    check_item<V>(a, *bsdf, i, bin, acc, key, hist);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for(int j=0; j < kCheckCols; ++j)
  {
    double v = acc[j];
#pragma unroll
    for(int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if(lane == 0) s_sum[warp][j] = v;
  }
#pragma unroll
  for(int off = 16; off > 0; off >>= 1) { unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off); if(o > key) key = o; }
  if(lane == 0) s_key[warp] = key;
  __syncthreads();
  if(threadIdx.x < kCheckCols)
  {
    double v = 0.0;
#pragma unroll
    for(int w=0; w < kCheckThreads/32; ++w) v += s_sum[w][threadIdx.x];
    if(a.mode == CHECK_BIN_PDF) { if(threadIdx.x == 0) a.partial[bin * gridDim.x + blockIdx.x] = v; }
    else a.partial[(size_t)blockIdx.x * kCheckCols + threadIdx.x] = v;
  }
  if(threadIdx.x == 0 && a.maxkey)
  {
    unsigned long long k = 0ull;
#pragma unroll
    for(int w=0; w < kCheckThreads/32; ++w) if(s_key[w] > k) k = s_key[w];
    a.maxkey[blockIdx.x] = k;
  }
  if(shared_hist)
  {
    __syncthreads();
    for(int i = threadIdx.x; i < bins; i += blockDim.x) if(s_hist[i]) atomicAdd(a.counts + i, s_hist[i]);
  }
}
#endif

enum : int { CHECK_KERNEL_E = 0, CHECK_KERNEL_SE, CHECK_KERNEL_S, CHECK_KERNEL_P };
// one translation unit per kernel kind (the model instantiations compile in parallel)
void launch_check_e(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, const CheckArgs&, dim3 grid);
void launch_check_se(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, const CheckArgs&, dim3 grid);
void launch_check_s(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, const CheckArgs&, dim3 grid);
void launch_check_p(bbmcu_ctx*, cudaStream_t, const BsdfDesc&, const CheckArgs&, dim3 grid);

} // namespace bbmcu
