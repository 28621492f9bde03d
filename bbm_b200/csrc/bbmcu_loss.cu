// (batched losses evaluate many parameter sets per launch: the EPD G1 rows are not launch-uniform here - read the table, not a staged copy)
#define BBMCU_EPD_NO_STAGE
// Loss objects of the C ABI (include/bbmcu.h): the six metrics over the MERL or spherical linearizer,
// batched over K parameter sets, with the analytic parameter gradient.
// Mirrors include/bbm/sampledlossfunction.h:26-95 and the thin metric classes of include/loss/*.h.
#include <cstring>
#include <memory>

#include "bbmcu_launch.cuh"
#include "bbmcu_losskernel.cuh"
#include "bbmcu_losscompact.cuh"

using namespace bbmcu;

// ---- exchange window of one shard (device memory, cudaIpc-exported) ------------------------------------------------------
//   PeerWord rows[2][world][cap]: rows[parity][r][i] = value i of rank r's K x (1+P) totals of the batch with that parity,
//   as two 8-byte words {low 32 bits of the double | seq << 32, high 32 bits | seq << 32} (seq = batch number mod 2^32).
// Every word carries its own arrival flag (aligned 8-byte stores are single-copy atomic), so there is no separate flag,
// fence or counter on the path: a value has arrived when both of its words show the batch's seq.
// Two parities are enough: rank r can only deliver batch s+2 after it has gathered batch s+1, which needs every peer's
// rows of s+1, and a peer sends those only after its own gather of batch s has finished (kernels of one shard run in
// stream order).
constexpr int kMaxPeers = 16;
struct PeerWord { unsigned long long lo, hi; };
struct PeerArgs
{
  int rank, world;
  unsigned int seq;
  size_t cap;
  unsigned char* win[kMaxPeers];
  unsigned int* status;
  unsigned long long timeout_ns;       // a peer that has not delivered after this long fails the batch (BBMCU_PEER_TIMEOUT_MS, default 10 s)
};

namespace bbmcu { SphericalGrid to_device_grid(const bbmcu_spherical_grid& g); }

struct bbmcu_loss
{
  bbmcu_ctx* ctx = nullptr;
  int device = 0;            // copy of ctx->device (bbmcu_loss_free must not touch a context that may be gone)
  int metric = 0, component = BBMCU_ALL, unit = 0;
  bool merl_grid = true;
  SphericalGrid grid{};
  uint64_t N = 0;            // samples of the whole linearizer
  uint64_t first = 0;        // this shard
  size_t count = 0;
  int il_world = 0, il_rank = 0;   // BBMCU_LOSS_SHARD_INTERLEAVED: blocks il_rank, il_rank + il_world, ... of kTileSamples samples
  int n_materials = 1;       // reference operands sharing this linearizer (bbmcu_loss_create_ex: a batch of measured tables)
  bool fused = true;         // directions generated inside the kernels (no d_in / d_out planes kept)
  const float* d_lin_tab = nullptr;   // the device's separable merl_linearizer table (owned by the library, per device), or d_sph_tab
  float* d_sph_tab = nullptr;         // this loss's spherical_linearizer table (sph_lin_entries x 2 floats)
  float* d_in = nullptr;     // 3 planes of count (materialised mode only)
  float* d_out = nullptr;
  float* d_ref = nullptr;    // n_materials x 3 planes of count
  // scratch, grown on demand
  float* d_attrs = nullptr;   size_t attrs_cap = 0;
  float* h_attrs[2] = {nullptr, nullptr}; size_t h_attrs_cap[2] = {0, 0};   // pinned, double buffered
  cudaEvent_t h_attrs_free[2] = {nullptr, nullptr};                          // recorded after the upload from each
  int h_flip = 0;
  double* d_partial = nullptr; size_t partial_cap = 0;
  double* d_result = nullptr;  size_t result_cap = 0;
  double* h_result = nullptr;  size_t h_result_cap = 0;  // pinned
  uint32_t* d_bad = nullptr;
  // exchange over peer memory (bbmcu_loss_peer_*): this shard's window, the peers' windows mapped into this process
  int peer_rank = 0, peer_world = 1;
  bool peer_connected = false;
  size_t peer_cap = 0;                                   // doubles per (parity, rank) row
  unsigned char* peer_win[kMaxPeers] = {};               // [rank] = that shard's window; own entry = own allocation
  bool peer_ipc[kMaxPeers] = {};                         // opened with cudaIpcOpenMemHandle (to be closed)
  unsigned int* h_peer_status = nullptr;                 // pinned, mapped: 1 = a peer did not arrive
  unsigned int peer_seq = 0;
  ~bbmcu_loss()
  {
    for(int r=0; r < kMaxPeers; ++r) if(peer_ipc[r] && peer_win[r]) cudaIpcCloseMemHandle(peer_win[r]);
    if(peer_win[peer_rank]) cudaFree(peer_win[peer_rank]);
    if(h_peer_status) cudaFreeHost(h_peer_status);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_ref); cudaFree(d_sph_tab); cudaFree(d_attrs); cudaFree(d_partial); cudaFree(d_result); cudaFree(d_bad);
    for(int i=0; i < 2; ++i) { if(h_attrs[i]) cudaFreeHost(h_attrs[i]); if(h_attrs_free[i]) cudaEventDestroy(h_attrs_free[i]); }
    if(h_result) cudaFreeHost(h_result);
  }
};

namespace {

// run-time lobe list (any aggregate): value pass + per-lobe jacobian pass, gradient through local memory
__global__ void __launch_bounds__(kLossThreads) k_loss_generic(const LossArgs a, const BsdfDesc shape)
{
  extern __shared__ double s_red[];                 // [warps][1 + P]
  __shared__ BsdfDesc b;
  __shared__ float s_lin[kMerlLinTabFloats];
  const int k = blockIdx.y, P = a.P;                // k runs over all materials' parameter sets
  const float* refp = a.ref + (size_t)(k / a.k_per_material) * a.ref_stride;
  if(threadIdx.x == 0) { b = shape; }
  loss_stage_lin(a, s_lin);
  __syncthreads();
  for(int i = threadIdx.x; i < a.n_attrs; i += blockDim.x) b.attrs[i] = loss_attr(a, (size_t)k*a.attr_stride + i);
  __syncthreads();
  double acc[1 + kMaxParams];
  for(int j=0; j <= P; ++j) acc[j] = 0.0;
  for(size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x*blockDim.x)
  {
    f3 in, out;
    loss_dirs(a, s_lin, i, in, out);
    Spec<float> ref(__ldg(refp + i), __ldg(refp + a.n + i), __ldg(refp + 2*a.n + i));
    float g[kMaxParams];
    float e = loss_sample_generic(b, a.metric, a.component, in, out, ref, a.want_grad ? g : nullptr);
    acc[0] += (double)e;
    if(a.want_grad) for(int j=0; j < P; ++j) acc[1 + j] += (double)g[j];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for(int j=0; j <= P; ++j)
  {
    double v = acc[j];
    for(int o=16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if(lane == 0) s_red[warp*(1 + P) + j] = v;
  }
  __syncthreads();
  if((int)threadIdx.x <= P)
  {
    double v = 0.0;
    for(int w=0; w < kLossThreads/32; ++w) v += s_red[w*(1 + P) + threadIdx.x];
    a.partial[((size_t)k*(1 + P) + threadIdx.x)*gridDim.x + blockIdx.x] = v;
  }
}

// second pass: block (k, column) adds that column's per-block partials - strided per thread, then a fixed shared-memory
// tree - and scales by 1/N.  The order depends only on the launch shape: deterministic.
constexpr int kFinishThreads = 128;
__global__ void __launch_bounds__(kFinishThreads) k_loss_finish(const double* partial, int blocks_x, int cols, double inv_n, double* result)
{
  __shared__ double s[kFinishThreads];
  const int k = blockIdx.x, j = blockIdx.y;
  const double* p = partial + ((size_t)k*cols + j)*blocks_x;
  double v = 0.0;
  for(int b = threadIdx.x; b < blocks_x; b += kFinishThreads) v += p[b];
  s[threadIdx.x] = v;
  __syncthreads();
  for(int o = kFinishThreads/2; o > 0; o >>= 1)
  {
    if((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if(threadIdx.x == 0) result[(size_t)k*cols + j] = s[0] * inv_n;
}

// ---- the same finish, fused with the exchange over peer memory ----------------------------------------------------------
__device__ __forceinline__ PeerWord* peer_row(unsigned char* win, int parity, int world, size_t cap, int r) { return reinterpret_cast<PeerWord*>(win) + ((size_t)parity*world + r)*cap; }
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// send: every block finishes one (k, column) total of this shard as k_loss_finish does and stores it - two flagged words -
// into the window of every peer (remote stores over NVLink; its own window included).  Nothing waits here.
__global__ void __launch_bounds__(kFinishThreads) k_loss_finish_send(const double* partial, int blocks_x, int cols, double inv_n, const PeerArgs pa)
{
  __shared__ double s[kFinishThreads];
  const int k = blockIdx.x, j = blockIdx.y, tid = threadIdx.x;
  const double* p = partial + ((size_t)k*cols + j)*blocks_x;
  double v = 0.0;
  for(int b = tid; b < blocks_x; b += kFinishThreads) v += p[b];
  s[tid] = v;
  __syncthreads();
  for(int o = kFinishThreads/2; o > 0; o >>= 1)
  {
    if(tid < o) s[tid] += s[tid + o];
    __syncthreads();
  }
  if(tid < pa.world)
  {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(s[0] * inv_n), flag = (unsigned long long)pa.seq << 32;
    PeerWord* dst = peer_row(pa.win[tid], (int)(pa.seq & 1u), pa.world, pa.cap, pa.rank) + ((size_t)k*cols + j);
    st_relaxed_sys(&dst->lo, (bits & 0xffffffffull) | flag);
    st_relaxed_sys(&dst->hi, (bits >> 32) | flag);
  }
}

// gather: thread i polls value i of every rank's row in its own window until both words carry this batch's seq, and adds
// them in rank order - the same order on every shard, so all of them return identical bits.  Bounded: a value that has not
// arrived after ~10 s sets status = 1 and yields NaN instead of hanging the device.
__global__ void __launch_bounds__(kFinishThreads) k_loss_gather(unsigned int nvals, double* result, const PeerArgs pa)
{
  const unsigned int i = blockIdx.x * kFinishThreads + threadIdx.x;
  if(i >= nvals) return;
  const PeerWord* rows = peer_row(pa.win[pa.rank], (int)(pa.seq & 1u), pa.world, pa.cap, 0) + i;
  double vals[kMaxPeers];
  unsigned int pending = (pa.world >= 32) ? 0xffffffffu : ((1u << pa.world) - 1u);
  const unsigned long long t0 = global_ns();
  bool bad = false;
  while(pending)
  {
#pragma unroll
    for(int r = 0; r < kMaxPeers; ++r)
    {
      if(!(pending & (1u << r))) continue;
      const PeerWord* w = rows + (size_t)r*pa.cap;
      const unsigned long long lo = ld_relaxed_sys(&w->lo), hi = ld_relaxed_sys(&w->hi);
      if((unsigned int)(lo >> 32) == pa.seq && (unsigned int)(hi >> 32) == pa.seq)
      {
        vals[r] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
        pending &= ~(1u << r);
      }
    }
    if(pending && (global_ns() - t0 > pa.timeout_ns)) { bad = true; break; }
  }
  double sum = 0.0;
#pragma unroll
  for(int r = 0; r < kMaxPeers; ++r) if(r < pa.world) sum += vals[r];
  if(bad) { *pa.status = 1u; sum = __longlong_as_double(0x7ff8000000000000ll); }
  result[i] = sum;
}

// the phi / theta sample table of a spherical_linearizer (bbmcu_linearizer.cuh); this translation unit is built -fmad=false
__global__ void k_sph_lin_tab(const SphericalGrid g, float* tab)
{
  const uint64_t n = sph_lin_entries(g);
  for(uint64_t j = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x*blockDim.x) sph_lin_entry(g, (uint32_t)j, tab[2*j], tab[2*j + 1]);
}

// per-sample terms l(idx) (sampledlossfunction::operator()(idx))
__global__ void __launch_bounds__(256) k_loss_terms(const LossArgs a, const BsdfDesc b, float* terms)
{
  __shared__ float s_lin[kMerlLinTabFloats];
  loss_stage_lin(a, s_lin);
  __syncthreads();
  for(size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x*blockDim.x)
  {
    f3 in, out;
    loss_dirs(a, s_lin, i, in, out);
    Spec<float> ref(a.ref[i], a.ref[a.n + i], a.ref[2*a.n + i]);
    terms[i] = loss_sample_generic(b, a.metric, a.component, in, out, ref, nullptr);
  }
}

template<class T> void grow(T*& p, size_t& cap, size_t need)
{
  if(need <= cap) return;
  if(p) BBMCU_CUDA(cudaFree(p));
  p = nullptr; cap = 0;
  BBMCU_CUDA(cudaMalloc(&p, need*sizeof(T)));
  cap = need;
}

// the part of LossArgs that says where samples come from (directions: generated or planes; reference planes)
void fill_sample_source(const bbmcu_loss* L, LossArgs& a)
{
  a.lin_mode = !L->fused ? LIN_MATERIALISED : (L->merl_grid ? LIN_MERL_TABLES : LIN_SPHERICAL);
  a.lin_tab = L->d_lin_tab; a.first = L->first; a.grid = L->grid; a.il_world = L->il_world; a.il_rank = L->il_rank;
  a.in = L->d_in; a.out = L->d_out; a.ref = L->d_ref; a.ref_stride = 3*L->count; a.n = L->count;
  a.k_per_material = 1; a.n_materials = 1;
}

void check_metric(int metric) { if(metric < 0 || metric > 5) throw std::invalid_argument("BBM: unknown loss metric " + std::to_string(metric)); }

} // anonymous namespace

extern "C" {

int bbmcu_loss_create_ex(bbmcu_ctx* ctx, int metric, const bbmcu_spherical_grid* grid, int component, int unit,
                         const bbmcu_bsdf* reference_bsdf, const float* const* reference_merl_rgb, int n_materials,
                         uint64_t first, uint64_t count, unsigned flags, bbmcu_loss** out)
{
  return guarded(ctx, [&] {
    if(!ctx || !out) throw std::invalid_argument("BBM: null argument");
    check_metric(metric);
    if((reference_bsdf != nullptr) == (reference_merl_rgb != nullptr)) throw std::invalid_argument("BBM: exactly one of reference_bsdf / reference_merl_rgb must be given");
    if(n_materials < 1 || (reference_bsdf && n_materials != 1)) throw std::invalid_argument("BBM: n_materials must be 1 for an analytic reference and >= 1 for measured tables");
    if(reference_merl_rgb) for(int m=0; m < n_materials; ++m) if(!reference_merl_rgb[m]) throw std::invalid_argument("BBM: null measured table");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    std::unique_ptr<bbmcu_loss> L(new bbmcu_loss);
    L->ctx = ctx; L->device = ctx->device; L->metric = metric; L->component = component; L->unit = unit;
    L->merl_grid = (grid == nullptr);
    L->n_materials = n_materials;
    L->fused = !(flags & BBMCU_LOSS_MATERIALISE_DIRECTIONS);
    if(grid) { L->grid = to_device_grid(*grid); L->N = L->grid.size(); } else L->N = kMerlBins;
    if(flags & BBMCU_LOSS_SHARD_INTERLEAVED)
    {
      // (first, count) = (rank, world): every world-th block of kTileSamples consecutive samples
      if(count < 1 || count > 65536 || first >= count) throw std::invalid_argument("BBM: interleaved loss shard needs first = rank < count = world");
      const uint64_t blocks = (L->N + kTileSamples - 1) / kTileSamples, world = count, rank = first;
      const uint64_t mine = rank < blocks ? (blocks - rank + world - 1) / world : 0;
      uint64_t n_il = mine * kTileSamples;
      if(mine && (rank + (mine - 1)*world) == blocks - 1) n_il -= blocks*kTileSamples - L->N;       // the grid's last, partial block is this shard's last
      if(world > 1) { L->il_world = (int)world; L->il_rank = (int)rank; }
      first = 0; count = n_il;
      if(world == 1) count = L->N;
    }
    else
    {
      if(first > L->N) throw std::out_of_range("BBM: loss shard starts beyond the linearizer size");
      if(count == 0) count = L->N - first;
      if(first + count > L->N) throw std::out_of_range("BBM: loss shard exceeds the linearizer size");
    }
    L->first = first; L->count = (size_t)count;
    const size_t n = L->count;
    if(L->merl_grid) L->d_lin_tab = merl_lin_table_device(ctx->device);
    else if(L->fused)
    {
      const uint64_t entries = sph_lin_entries(L->grid);
      if(entries > (uint64_t(1) << 22)) L->fused = false;              // a 32 MB table would not stay on chip: keep the planes instead
      else
      {
        BBMCU_CUDA(cudaMalloc(&L->d_sph_tab, 2*entries*sizeof(float)));
        k_sph_lin_tab<<<(unsigned)std::min<uint64_t>(1024, (entries + 255) / 256), 256, 0, ctx->stream>>>(L->grid, L->d_sph_tab);
        BBMCU_CUDA(cudaGetLastError());
        L->d_lin_tab = L->d_sph_tab;
      }
    }
    if(n == 0) { *out = L.release(); return; }
    BBMCU_CUDA(cudaMalloc(&L->d_in, 3*n*sizeof(float)));
    BBMCU_CUDA(cudaMalloc(&L->d_out, 3*n*sizeof(float)));
    BBMCU_CUDA(cudaMalloc(&L->d_ref, (size_t)n_materials*3*n*sizeof(float)));
    BBMCU_CUDA(cudaMalloc(&L->d_bad, sizeof(uint32_t)));
    BBMCU_CUDA(cudaMemsetAsync(L->d_bad, 0, sizeof(uint32_t), ctx->stream));
    // 1. the linearizer's directions for this shard: needed here to tabulate the reference operand; kept only in
    //    materialised mode (the fused kernels regenerate the same bits from the bin index)
    const bool al = true;                     // cudaMalloc'ed planes; launch_foreach4 checks the plane stride
    // (an interleaved shard is generated block by block into planes that are n floats apart)
    const size_t piece = L->il_world > 1 ? (size_t)kTileSamples : n;
    const size_t saved_ld = ctx->ld;
    ctx->ld = n;
    try
    {
      for(size_t off = 0; off < n; off += piece)
      {
        const size_t len = std::min(piece, n - off);
        const uint64_t lin = L->il_world > 1 ? ((uint64_t)(off / kTileSamples) * (uint64_t)L->il_world + (uint64_t)L->il_rank) * kTileSamples : first;
        if(L->merl_grid) { MerlDirsOp op; op.first = (uint32_t)lin; op.in = L->d_in + off; op.out = L->d_out + off; op.n = len; op.aligned = al; launch_foreach4(ctx, ctx->stream, op, len); }
        else { SphericalDirsOp op; op.grid = L->grid; op.first = lin; op.in = L->d_in + off; op.out = L->d_out + off; op.n = len; op.aligned = al; launch_foreach4(ctx, ctx->stream, op, len); }
      }
    }
    catch(...) { ctx->ld = saved_ld; throw; }
    ctx->ld = saved_ld;
    // 2. the reference operand tabulated at those directions (it never changes during a fit)
    if(reference_bsdf)
    {
      BsdfDesc d = make_desc(reference_bsdf->b, ctx->device);
      launch_eval(ctx, ctx->stream, d, component, L->d_in, L->d_out, L->d_ref, n);
    }
    else
    {
      // merl_data::eval: component must be exactly All (staticmodel/merl.h:83), nearest-bin lookup.  Host tables pass
      // through one device staging buffer, one material after the other.
      float* d_table = nullptr;
      for(int m=0; m < n_materials; ++m)
      {
        float* ref_m = L->d_ref + (size_t)m*3*n;
        if(component != BBMCU_ALL) { BBMCU_CUDA(cudaMemsetAsync(ref_m, 0, 3*n*sizeof(float), ctx->stream)); continue; }
        const float* table = reference_merl_rgb[m];
        if(!is_device_pointer(table))
        {
          if(!d_table) BBMCU_CUDA(cudaMalloc(&d_table, 3*(size_t)kMerlBins*sizeof(float)));
          BBMCU_CUDA(cudaMemcpyAsync(d_table, table, 3*(size_t)kMerlBins*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
          table = d_table;
        }
        MerlLookupOp op; op.table = table; op.in = L->d_in; op.out = L->d_out; op.rgb = ref_m; op.bad = L->d_bad; op.n = n; op.aligned = al;
        launch_foreach4(ctx, ctx->stream, op, n);
      }
      uint32_t bad = 0;
      BBMCU_CUDA(cudaMemcpyAsync(&bad, L->d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
      BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
      if(d_table) BBMCU_CUDA(cudaFree(d_table));
      // the reference throws from lookup() on such samples
      if(bad) throw std::out_of_range("BBM: " + std::to_string(bad) + " samples of the linearizer map outside the MERL table (NaN direction pairs); the reference throws 'lookup out of range' here");
    }
    BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
    if(L->fused) { BBMCU_CUDA(cudaFree(L->d_in)); BBMCU_CUDA(cudaFree(L->d_out)); L->d_in = L->d_out = nullptr; }
    *out = L.release();
  });
}

int bbmcu_loss_create(bbmcu_ctx* ctx, int metric, const bbmcu_spherical_grid* grid, int component, int unit,
                      const bbmcu_bsdf* reference_bsdf, const float* reference_merl_rgb,
                      uint64_t first, uint64_t count, bbmcu_loss** out)
{
  return bbmcu_loss_create_ex(ctx, metric, grid, component, unit, reference_bsdf, reference_merl_rgb ? &reference_merl_rgb : nullptr, 1, first, count, 0u, out);
}

int bbmcu_loss_set_metric(bbmcu_loss* L, int metric)
{
  return guarded(L ? L->ctx : nullptr, [&] { if(!L) throw std::invalid_argument("BBM: null argument"); check_metric(metric); L->metric = metric; });
}
int bbmcu_loss_materials(const bbmcu_loss* L) { return L ? L->n_materials : 0; }

// the loss keeps its own copy of the device index: the context may already be gone (Python shutdown order, a cuda_loss that
// outlives its context), so nothing of *ctx is touched here
void bbmcu_loss_free(bbmcu_loss* L) { if(L) { if(cudaSetDevice(L->device) == cudaSuccess) cudaDeviceSynchronize(); delete L; } }
uint64_t bbmcu_loss_samples(const bbmcu_loss* L) { return L ? L->N : 0; }
uint64_t bbmcu_loss_shard_count(const bbmcu_loss* L) { return L ? (uint64_t)L->count : 0; }

int bbmcu_loss_eval(bbmcu_loss* L, const bbmcu_bsdf* bsdf, const double* params, size_t K, double* loss_out, double* grad_out, double* device_out)
{
  if(L && L->n_materials != 1) { bbmcu_ctx* c = L->ctx; return guarded(c, [&] { throw std::invalid_argument("BBM: this loss holds " + std::to_string(L->n_materials) + " materials: use bbmcu_loss_eval_multi"); }); }
  return bbmcu_loss_eval_multi(L, bsdf, params, K, loss_out, grad_out, device_out);
}

int bbmcu_loss_eval_multi(bbmcu_loss* L, const bbmcu_bsdf* bsdf, const double* params, size_t Kper, double* loss_out, double* grad_out, double* device_out)
{ return bbmcu_loss_eval_multi_ex(L, bsdf, params, Kper, loss_out, grad_out, device_out, -1); }

// Kper parameter sets for EACH of the loss's M materials in one launch: params M x Kper x P (material-major), results M x Kper
int bbmcu_loss_eval_multi_ex(bbmcu_loss* L, const bbmcu_bsdf* bsdf, const double* params, size_t Kper, double* loss_out, double* grad_out, double* device_out, int gradient)
{
  bbmcu_ctx* ctx = L ? L->ctx : nullptr;
  return guarded(ctx, [&] {
    if(!L || !bsdf) throw std::invalid_argument("BBM: null argument");
    if(Kper == 0) return;
    const size_t M = (size_t)L->n_materials, K = M*Kper;                 // K: parameter sets of the whole launch
    if(!params && K != 1) throw std::invalid_argument("BBM: params == NULL requires one material and K == 1");
    if(L->h_peer_status && *(volatile unsigned int*)L->h_peer_status) throw std::runtime_error("BBM: a peer shard did not arrive at an earlier loss exchange within 10 s");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    const int P = bsdf->b.param_count(BBMCU_ATTR_ALL);
    if(P > kMaxParams) throw std::invalid_argument("BBM: more than " + std::to_string(kMaxParams) + " fit parameters");
    BsdfDesc shape = make_desc(bsdf->b, ctx->device);
    const int A = shape.n_floats;                 // device floats per parameter set (He lobes carry an unused table gap)
    const bool want_grad = gradient < 0 ? ((grad_out != nullptr) || (device_out != nullptr)) : (gradient != 0);
    if(!want_grad && grad_out) throw std::invalid_argument("BBM: grad_out given but gradient == 0");
    const int cols = 1 + P;
    // attribute blocks for the K parameter sets: inside the kernel arguments when they fit (one compass step), else
    // through two pinned staging buffers - packing the next batch never waits for the kernels of the previous one, only
    // (and in practice never) for the upload that last read the buffer it is about to overwrite
    LossArgs a;
    const bool inline_attrs = K*(size_t)A <= (size_t)kInlineAttrFloats;
    a.inline_count = inline_attrs ? (int)(K*(size_t)A) : 0;
    float* h_attrs = a.inline_attrs;
    int hb = 0;
    if(!inline_attrs)
    {
      grow(L->d_attrs, L->attrs_cap, K*(size_t)A);
      hb = L->h_flip; L->h_flip ^= 1;
      if(!L->h_attrs_free[hb]) BBMCU_CUDA(cudaEventCreateWithFlags(&L->h_attrs_free[hb], cudaEventDisableTiming));
      else BBMCU_CUDA(cudaEventSynchronize(L->h_attrs_free[hb]));
      if(K*(size_t)A > L->h_attrs_cap[hb])
      {
        if(L->h_attrs[hb]) { BBMCU_CUDA(cudaFreeHost(L->h_attrs[hb])); L->h_attrs[hb] = nullptr; L->h_attrs_cap[hb] = 0; }
        BBMCU_CUDA(cudaMallocHost(&L->h_attrs[hb], K*(size_t)A*sizeof(float)));
        L->h_attrs_cap[hb] = K*(size_t)A;
      }
      h_attrs = L->h_attrs[hb];
    }
    {
      bbmcu_host::Bsdf tmp = bsdf->b;
      for(size_t k=0; k < K; ++k)
      {
        if(params) tmp.set_params(BBMCU_ATTR_ALL, params + k*P, P);
        // start from the descriptor's block: it carries what is not a parameter value - the device address of a Merl(...)
        // lobe's table (two float slots, bbmcu_desc.hpp) and the zeroed table gaps of He lobes - then overlay the values
        std::memcpy(h_attrs + k*A, shape.attrs, (size_t)A*sizeof(float));
        for(size_t l=0; l < tmp.lobes.size(); ++l)
        {
          size_t off = (size_t)shape.offset[l];
          for(double v : tmp.lobes[l].values) h_attrs[k*A + off++] = (float)v;
        }
      }
    }
    if(!inline_attrs)
    {
      BBMCU_CUDA(cudaMemcpyAsync(L->d_attrs, h_attrs, K*(size_t)A*sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
      BBMCU_CUDA(cudaEventRecord(L->h_attrs_free[hb], ctx->stream));
    }
    // launch shape: about 8 resident blocks per SM over all K
    const size_t n = L->count;
    const bool static_shape = (shape.n_lobes == 1 && !shape.aggregate) || (shape.n_lobes == 2 && shape.aggregate && shape.model[0] == M_Lambertian);
    unsigned max_bx = (unsigned)std::max<size_t>(1, (n + kLossThreads - 1) / kLossThreads);
    unsigned bx = (unsigned)std::max<size_t>(1, ((size_t)ctx->sm_count*8 + Kper - 1) / Kper);     // from the per-material count: a material's sums do not depend on how many others share the launch
    if(bx > max_bx) bx = max_bx;
    if(static_shape) bx = (unsigned)std::max<size_t>(1, (n + kTileSamples - 1) / kTileSamples);     // one partial row per sample tile
    grow(L->d_partial, L->partial_cap, K*(size_t)bx*cols);
    grow(L->d_result, L->result_cap, K*(size_t)cols);
    fill_sample_source(L, a);
    a.attrs = L->d_attrs; a.attr_stride = A; a.n_attrs = A; a.k_per_material = (int)Kper; a.n_materials = (int)M;
    a.metric = L->metric; a.component = L->component; a.want_grad = want_grad ? 1 : 0; a.partial = L->d_partial; a.P = P; a.sm_count = ctx->sm_count;
    bool done = false;
    if(n > 0)
    {
      const int m0 = shape.model[0];
      const char* nc = std::getenv("BBMCU_LOSS_NO_COMPACT");     // read per call: tests compare the two kernels in one process
      const bool no_compact = nc && *nc && *nc != '0';
      // the compact kernel (bbmcu_losscompact.cuh) where the lobe has one and both components are asked for
      if(shape.n_lobes == 1 && !shape.aggregate && !no_compact && a.component == FLAG_ALL) done = launch_loss_single_compact(m0, ctx->stream, a, (unsigned)Kper);
      if(done) {}
      else if(shape.n_lobes == 1 && !shape.aggregate)
        done = launch_loss_single_g0(m0, ctx->stream, a, bx, (unsigned)Kper) || launch_loss_single_g1(m0, ctx->stream, a, bx, (unsigned)Kper) ||
               launch_loss_single_g2(m0, ctx->stream, a, bx, (unsigned)Kper) || launch_loss_single_g3(m0, ctx->stream, a, bx, (unsigned)Kper);
      else if(shape.n_lobes == 2 && shape.aggregate && m0 == M_Lambertian)
      {
        const int m1 = shape.model[1];
        if(!no_compact && a.component == FLAG_ALL) done = launch_loss_pair_compact(m1, ctx->stream, a, (unsigned)Kper);
        if(!done) done = launch_loss_pair_g0(m1, ctx->stream, a, bx, (unsigned)Kper) || launch_loss_pair_g1(m1, ctx->stream, a, bx, (unsigned)Kper) ||
               launch_loss_pair_g2(m1, ctx->stream, a, bx, (unsigned)Kper) || launch_loss_pair_g3(m1, ctx->stream, a, bx, (unsigned)Kper);
      }
      if(!done) bind_device_tables();
      if(!done) k_loss_generic<<<dim3(bx, (unsigned)K), kLossThreads, (kLossThreads/32)*cols*sizeof(double), ctx->stream>>>(a, shape);
      BBMCU_CUDA(cudaGetLastError());
      ++ctx->launches;
    }
    else BBMCU_CUDA(cudaMemsetAsync(L->d_partial, 0, K*(size_t)bx*cols*sizeof(double), ctx->stream));
    double* result = device_out ? device_out : L->d_result;
    if(L->peer_connected && L->peer_world > 1)
    {
      if(K*(size_t)cols > L->peer_cap) throw std::invalid_argument("BBM: K*(1+P) = " + std::to_string(K*(size_t)cols) + " exceeds the peer window (" + std::to_string(L->peer_cap) + " values)");
      PeerArgs pa{};
      pa.rank = L->peer_rank; pa.world = L->peer_world; pa.cap = L->peer_cap;
      static const unsigned long long timeout_ns = [] { const char* e = std::getenv("BBMCU_PEER_TIMEOUT_MS"); long v = e ? std::atol(e) : 0; return (unsigned long long)(v > 0 ? v : 10000) * 1000000ull; }();
      pa.timeout_ns = timeout_ns;
      if(++L->peer_seq == 0u) L->peer_seq = 2u;                     // 0 marks a fresh window; 2 keeps the parities alternating across the wrap
      pa.seq = L->peer_seq;
      for(int r=0; r < L->peer_world; ++r) pa.win[r] = L->peer_win[r];
      BBMCU_CUDA(cudaHostGetDevicePointer((void**)&pa.status, L->h_peer_status, 0));
      k_loss_finish_send<<<dim3((unsigned)K, (unsigned)cols), kFinishThreads, 0, ctx->stream>>>(L->d_partial, (int)bx, cols, 1.0 / (double)L->N, pa);
      const unsigned nvals = (unsigned)(K*(size_t)cols);
      k_loss_gather<<<(nvals + kFinishThreads - 1) / kFinishThreads, kFinishThreads, 0, ctx->stream>>>(nvals, result, pa);
      ++ctx->launches;
    }
    else
      k_loss_finish<<<dim3((unsigned)K, (unsigned)cols), kFinishThreads, 0, ctx->stream>>>(L->d_partial, (int)bx, cols, 1.0 / (double)L->N, result);
    BBMCU_CUDA(cudaGetLastError());
    ++ctx->launches;
    if(device_out) return;                       // caller all-reduces / reads it on the stream
    if(K*(size_t)cols > L->h_result_cap)
    {
      if(L->h_result) { BBMCU_CUDA(cudaFreeHost(L->h_result)); L->h_result = nullptr; L->h_result_cap = 0; }
      BBMCU_CUDA(cudaMallocHost(&L->h_result, K*(size_t)cols*sizeof(double)));
      L->h_result_cap = K*(size_t)cols;
    }
    BBMCU_CUDA(cudaMemcpyAsync(L->h_result, L->d_result, K*(size_t)cols*sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
    if(L->h_peer_status && *(volatile unsigned int*)L->h_peer_status) throw std::runtime_error("BBM: a peer shard did not arrive at the loss exchange within 10 s");
    for(size_t k=0; k < K; ++k)
    {
      if(loss_out) loss_out[k] = L->h_result[k*cols];
      if(grad_out) for(int j=0; j < P; ++j) grad_out[k*P + j] = L->h_result[k*cols + 1 + j];
    }
  });
}

int bbmcu_loss_peer_init(bbmcu_loss* L, int rank, int world, size_t max_values, unsigned char handle_out[64], void** window_out)
{
  bbmcu_ctx* ctx = L ? L->ctx : nullptr;
  return guarded(ctx, [&] {
    if(!L) throw std::invalid_argument("BBM: null argument");
    if(world < 1 || world > kMaxPeers || rank < 0 || rank >= world) throw std::invalid_argument("BBM: peer rank/world out of range (at most " + std::to_string(kMaxPeers) + " shards)");
    if(max_values == 0) throw std::invalid_argument("BBM: max_values == 0");
    if(L->peer_win[L->peer_rank]) throw std::invalid_argument("BBM: the loss already has a peer window");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = 2*(size_t)world*max_values*sizeof(PeerWord);
    unsigned char* win = nullptr;
    BBMCU_CUDA(cudaMalloc(&win, bytes));
    BBMCU_CUDA(cudaMemset(win, 0, bytes));
    BBMCU_CUDA(cudaHostAlloc((void**)&L->h_peer_status, sizeof(unsigned int), cudaHostAllocMapped));
    *L->h_peer_status = 0u;
    BBMCU_CUDA(cudaDeviceSynchronize());
    L->peer_rank = rank; L->peer_world = world; L->peer_cap = max_values; L->peer_win[rank] = win; L->peer_seq = 0;
    if(handle_out)
    {
      cudaIpcMemHandle_t h;
      BBMCU_CUDA(cudaIpcGetMemHandle(&h, win));
      std::memcpy(handle_out, &h, 64);
    }
    if(window_out) *window_out = win;
  });
}

int bbmcu_loss_peer_connect(bbmcu_loss* L, const unsigned char* handles)
{
  bbmcu_ctx* ctx = L ? L->ctx : nullptr;
  return guarded(ctx, [&] {
    if(!L || !handles) throw std::invalid_argument("BBM: null argument");
    if(!L->peer_win[L->peer_rank]) throw std::invalid_argument("BBM: bbmcu_loss_peer_init first");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    for(int r=0; r < L->peer_world; ++r)
    {
      if(r == L->peer_rank) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, handles + 64*(size_t)r, 64);
      void* p = nullptr;
      BBMCU_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      L->peer_win[r] = static_cast<unsigned char*>(p); L->peer_ipc[r] = true;
    }
    L->peer_connected = true;
  });
}

int bbmcu_loss_peer_connect_ptrs(bbmcu_loss* L, void* const* windows)
{
  bbmcu_ctx* ctx = L ? L->ctx : nullptr;
  return guarded(ctx, [&] {
    if(!L || !windows) throw std::invalid_argument("BBM: null argument");
    if(!L->peer_win[L->peer_rank]) throw std::invalid_argument("BBM: bbmcu_loss_peer_init first");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    for(int r=0; r < L->peer_world; ++r)
    {
      if(r == L->peer_rank) continue;
      if(!windows[r]) throw std::invalid_argument("BBM: null peer window");
      cudaPointerAttributes at{};
      BBMCU_CUDA(cudaPointerGetAttributes(&at, windows[r]));
      if(at.type != cudaMemoryTypeDevice) throw std::invalid_argument("BBM: a peer window is not device memory");
      if(at.device != ctx->device)
      {
        cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
        if(e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else BBMCU_CUDA(e);
      }
      L->peer_win[r] = static_cast<unsigned char*>(windows[r]);
    }
    L->peer_connected = true;
  });
}

int bbmcu_loss_terms(bbmcu_loss* L, const bbmcu_bsdf* bsdf, float* terms) { return bbmcu_loss_terms_at(L, bsdf, 0, terms); }

int bbmcu_loss_terms_at(bbmcu_loss* L, const bbmcu_bsdf* bsdf, int material, float* terms)
{
  bbmcu_ctx* ctx = L ? L->ctx : nullptr;
  return guarded(ctx, [&] {
    if(!L || !bsdf || !terms) throw std::invalid_argument("BBM: null argument");
    if(material < 0 || material >= L->n_materials) throw std::out_of_range("BBM: material index out of range");
    if(L->count == 0) return;
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc shape = make_desc(bsdf->b, ctx->device);
    LossArgs a{};
    fill_sample_source(L, a);
    a.ref = L->d_ref + (size_t)material*a.ref_stride;
    a.metric = L->metric; a.component = L->component;
    const bool dev = is_device_pointer(terms);
    float* d_terms = terms;
    if(!dev) BBMCU_CUDA(cudaMalloc(&d_terms, L->count*sizeof(float)));
    bind_device_tables();
    k_loss_terms<<<grid_for(ctx, L->count), 256, 0, ctx->stream>>>(a, shape, d_terms);
    BBMCU_CUDA(cudaGetLastError());
    ++ctx->launches;
    if(!dev)
    {
      BBMCU_CUDA(cudaMemcpyAsync(terms, d_terms, L->count*sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
      BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));
      BBMCU_CUDA(cudaFree(d_terms));
    }
  });
}

} // extern "C"
