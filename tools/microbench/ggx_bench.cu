// Stand-alone timing of the headline kernel (k_foreach4<SampleEvalPdfOp<BsdfSingle<GGX>>>) for quick iteration on
// launch shape and code changes without rebuilding libbbmcu.so:  nvcc ... -DBLOCK=256 -DMINB=3 ggx_bench.cu -o ggx_bench
// Development tool only; bench.py measures the shipped library.
#include <cstdio>
#include <vector>
#include "bbmcu_launch.cuh"
namespace bbmcu { const float* epd_table_device(int) { return nullptr; } }
using namespace bbmcu;
#ifndef BLOCK
#define BLOCK 256
#endif
#ifndef MINB
#define MINB 1
#endif
#ifndef MODEL
#define MODEL M_GGX
#endif
#ifndef OP
#define OP 2          // 0 eval, 1 sample, 2 sample + eval + pdf
#endif
using GGXM = ModelOf<MODEL>::type;
#ifdef PAIR            // Aggregate(Lambertian, MODEL) on the compile-time pair kernels
using BsdfT = BsdfPair<Lambertian, GGXM>;
#else
using BsdfT = BsdfSingle<GGXM>;
#endif
#if OP == 0
using Op = EvalOp<BsdfT>;
#elif OP == 1
using Op = SampleOp<BsdfT>;
#else
using Op = SampleEvalPdfOp<BsdfT>;
#endif
__global__ void __launch_bounds__(BLOCK, MINB) k_var(const Op op, size_t groups)
{
  for(size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) op.group(g * kVec, op.bsdf);
}
__global__ void k_init(float* out, float* xi, size_t n, uint32_t salt = 0)
{
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    uint32_t s = ((uint32_t)i + salt) * 747796405u + 2891336453u; auto rnd = [&]() { s ^= s >> 17; s *= 0xed5ad4bbu; s ^= s >> 11; s *= 0xac4c1b51u; s ^= s >> 15; return (s >> 8) * (1.0f / 16777216.0f); };
    float z = rnd(), ph = rnd() * 6.2831853f, r = sqrtf(fmaxf(1 - z*z, 0.f));
    out[i] = r*cosf(ph); out[n + i] = r*sinf(ph); out[2*n + i] = z; xi[i] = rnd(); xi[n + i] = rnd();
  }
}
int main(int argc, char** argv)
{
  const size_t n = size_t(1) << (argc > 1 ? atoi(argv[1]) : 26);
  const int blocks_per_sm = argc > 2 ? atoi(argv[2]) : 32;
  float *out, *xi, *dir, *sp, *rgb, *pdf; int32_t* flag;
  cudaMalloc(&out, 3*n*4); cudaMalloc(&xi, 2*n*4); cudaMalloc(&dir, 3*n*4); cudaMalloc(&sp, n*4); cudaMalloc(&rgb, 3*n*4); cudaMalloc(&pdf, n*4); cudaMalloc(&flag, n*4);
  k_init<<<1184, 256>>>(out, xi, n);
  Op op; memset(&op.bsdf, 0, sizeof(op.bsdf));
  op.bsdf.n_lobes = 1; op.bsdf.model[0] = MODEL;
  // default parameters of the model come from a file written by the host library (tools/microbench/defaults.py): "n v0 v1 ..."
  { float a[64] = {0.5f, 0.5f, 0.5f, 0.1f, 1.3f}; int na = 5;
    if(argc > 3) { FILE* f = fopen(argv[3], "r"); if(f) { if(fscanf(f, "%d", &na) == 1) for(int i=0; i < na && i < 64; ++i) if(fscanf(f, "%f", &a[i]) != 1) break; fclose(f); } }
#ifdef PAIR
    const float lam[3] = {0.2f, 0.1f, 0.05f};
    memcpy(op.bsdf.attrs, lam, sizeof(lam)); memcpy(op.bsdf.attrs + 3, a, sizeof(float)*na); op.bsdf.n_floats = 3 + na;
    op.bsdf.n_lobes = 2; op.bsdf.aggregate = 1; op.bsdf.model[0] = M_Lambertian; op.bsdf.model[1] = MODEL; op.bsdf.offset[0] = 0; op.bsdf.offset[1] = 3;
#else
    memcpy(op.bsdf.attrs, a, sizeof(float)*na); op.bsdf.n_floats = na;
#endif
  }
  op.component = 3; op.out = out; op.n = n; op.aligned = true;
#if OP == 0
  op.in = dir; op.rgb = rgb; k_init<<<1184, 256>>>(dir, xi, n, 0x9e3779b9u);            // independent incident directions
#elif OP == 1
  op.xi = xi; op.dir = dir; op.pdf = sp; op.flag = flag;
#else
  op.xi = xi; op.dir = dir; op.spdf = sp; op.flag = flag; op.rgb = rgb; op.pdf = pdf;
#endif
  size_t groups = n / kVec;
  unsigned grid = (unsigned)std::min<size_t>((groups + BLOCK - 1) / BLOCK, (size_t)148 * blocks_per_sm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int i=0; i < 3; ++i) k_var<<<grid, BLOCK>>>(op, groups);
  cudaEventRecord(e0);
  const int reps = 10;
  for(int i=0; i < reps; ++i) k_var<<<grid, BLOCK>>>(op, groups);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_var);
  std::vector<float> h(8); cudaMemcpy(h.data(), pdf, 32, cudaMemcpyDeviceToHost);
  printf("MODEL=%d OP=%d BLOCK=%d MINB=%d KVEC=%d grid=%u regs=%d  %.3f ms  %.2f G/s  pdf[1]=%g err=%s\n", (int)MODEL, OP, BLOCK, MINB, kVec, grid, fa.numRegs, ms,
         n / ms / 1e6, h[1], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
