// libbbmcu.so C ABI (include/bbmcu.h): context, model registry, BSDF objects, batched
// eval / sample / pdf / reflectance, linearizers, MERL binaries and .fit files.
// Loss entry points live in bbmcu_loss.cu.
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <thread>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <emmintrin.h>

#include "bbmcu_launch.cuh"

using namespace bbmcu;

namespace bbmcu {

static thread_local std::string g_thread_error;
void set_thread_error(const std::string& msg) { g_thread_error = msg; }
const char* thread_error() { return g_thread_error.c_str(); }

bool is_device_pointer(const void* p)
{
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if(e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// the Holzschuch-Pacanowski G1 table: bbm_b200/data/epd_g1.f32 linked into the library (build.py: ld -r -b binary)
extern "C" { extern const unsigned char _binary_epd_g1_f32_start[]; extern const unsigned char _binary_epd_g1_f32_end[]; }
const float* epd_table_device(int device)
{
  static std::mutex mtx;
  static float* table[64] = {};
  std::lock_guard<std::mutex> lock(mtx);
  if(device < 0 || device >= 64) throw std::invalid_argument("BBM: device index out of range");
  if(!table[device])
  {
    const size_t bytes = (size_t)(_binary_epd_g1_f32_end - _binary_epd_g1_f32_start);
    if(bytes != sizeof(float)*kEpdRows*kEpdCols) throw std::runtime_error("BBM: embedded EPD G1 table has the wrong size");
    int cur = 0; BBMCU_CUDA(cudaGetDevice(&cur));
    BBMCU_CUDA(cudaSetDevice(device));
    float* p = nullptr;
    BBMCU_CUDA(cudaMalloc(&p, bytes));
    BBMCU_CUDA(cudaMemcpy(p, _binary_epd_g1_f32_start, bytes, cudaMemcpyHostToDevice));
    BBMCU_CUDA(cudaSetDevice(cur));
    table[device] = p;
  }
  return table[device];
}

// the separable merl_linearizer table (bbmcu_linearizer.cuh): 360 rows filled by the very device functions merl_dirs() uses
__global__ void k_merl_lin_tab(float* tab) { const int j = blockIdx.x * blockDim.x + threadIdx.x; if(j < 360) merl_lin_tab_fill(tab, j); }
const float* merl_lin_table_device(int device)
{
  static std::mutex mtx;
  static float* table[64] = {};
  std::lock_guard<std::mutex> lock(mtx);
  if(device < 0 || device >= 64) throw std::invalid_argument("BBM: device index out of range");
  if(!table[device])
  {
    int cur = 0; BBMCU_CUDA(cudaGetDevice(&cur));
    BBMCU_CUDA(cudaSetDevice(device));
    float* p = nullptr;
    BBMCU_CUDA(cudaMalloc(&p, kMerlLinTabFloats*sizeof(float)));
    k_merl_lin_tab<<<3, 128>>>(p);
    BBMCU_CUDA(cudaGetLastError());
    BBMCU_CUDA(cudaDeviceSynchronize());
    BBMCU_CUDA(cudaSetDevice(cur));
    table[device] = p;
  }
  return table[device];
}

// device copies of MERL tables: created on first use per device, released with the MerlData (its `release` hook)
static void merl_release(int device, void* ptr)
{
  int cur = 0;
  if(cudaGetDevice(&cur) != cudaSuccess) return;          // driver already shut down
  cudaSetDevice(device); cudaFree(ptr); cudaSetDevice(cur);
}
const float* merl_device_table(const bbmcu_host::MerlData& m, int device)
{
  static std::mutex mtx;
  std::lock_guard<std::mutex> lock(mtx);
  auto it = m.device.find(device);
  if(it != m.device.end()) return static_cast<const float*>(it->second);
  int cur = 0; BBMCU_CUDA(cudaGetDevice(&cur));
  BBMCU_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  BBMCU_CUDA(cudaMalloc(&p, m.rgb.size()*sizeof(float)));
  BBMCU_CUDA(cudaMemcpy(p, m.rgb.data(), m.rgb.size()*sizeof(float), cudaMemcpyHostToDevice));
  BBMCU_CUDA(cudaSetDevice(cur));
  m.device[device] = p;
  m.release = merl_release;
  return static_cast<const float*>(p);
}

namespace {

// 4-byte elements, SoA planes; an OUTPUT pointer may be null (not wanted).  How a host output travels back over PCIe:
//   XFER_PLAIN    as it is (4 bytes per element)
//   XFER_BYTES    one-plane int32 output with values in [0, 255] (bsdf_flag): narrowed to one byte per element on the device,
//                 copied into the pinned ring and widened into the caller's array by host threads (1 byte over the link, not 4)
//   XFER_COPY_OF / XFER_MASKED_BY  not transferred at all: equal to another output of the same call (`src`), or to it where a
//                 third output (`mask`: the flag) is non-zero and 0 elsewhere - the host threads write it from those
// (the fused pass returns sample.pdf next to pdf(sample.direction, out), which most models define as the same number:
// 36 -> 29 bytes per pair down the link, the side the pass is bound by)
//   XFER_GRAY_SCALED  a three-plane output whose planes are one device plane u times three launch constants (eval of a
//                 single scaled lobe whose unscaled value is gray): the kernel writes u, u crosses the link (4 bytes per
//                 element, not 12) into the pinned ring and host threads write u * scale[c] into the caller's planes - the
//                 IEEE single multiplications the device kernel does, so the same bits
enum : int { XFER_PLAIN = 0, XFER_BYTES, XFER_COPY_OF, XFER_MASKED_BY, XFER_GRAY_SCALED };
struct ArrayArg { const void* ptr; int planes; bool input; int xfer = XFER_PLAIN; int src = -1; int mask = -1; float scale[3] = {0.0f, 0.0f, 0.0f}; };

enum class Mem { Device, Pinned, Pageable };
Mem classify(const void* p)
{
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if(e != cudaSuccess) { cudaGetLastError(); return Mem::Pageable; }
  if(a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return Mem::Device;
  return a.type == cudaMemoryTypeHost ? Mem::Pinned : Mem::Pageable;       // Host = cudaMallocHost / cudaHostRegister'ed
}

// classify the buffers of one call: all device (true) or all host (false); mixing is an error
bool all_device(const std::vector<ArrayArg>& args, std::vector<Mem>& kind)
{
  int dev = 0, host = 0;
  kind.assign(args.size(), Mem::Device);
  for(size_t a=0; a < args.size(); ++a)
  {
    if(!args[a].ptr) { if(args[a].input) throw std::invalid_argument("BBM: null input pointer"); continue; }
    kind[a] = classify(args[a].ptr);
    (kind[a] == Mem::Device ? dev : host)++;
  }
  if(dev && host) throw std::invalid_argument("BBM: data pointers of one call must be all host or all device memory");
  if(!dev && !host) throw std::invalid_argument("BBM: every output pointer of the call is null");
  return dev > 0;
}

constexpr size_t kChunk = size_t(1) << 21;       // elements per host-pointer chunk

// rows x bytes strided copy on `threads` host threads (pageable memory <-> pinned staging)
void host_copy_rows(char* dst, size_t dst_pitch, const char* src, size_t src_pitch, size_t bytes, int rows, int threads)
{
  if(bytes * (size_t)rows < (size_t(1) << 20) || threads <= 1) { for(int r=0; r < rows; ++r) std::memcpy(dst + r*dst_pitch, src + r*src_pitch, bytes); return; }
  std::vector<std::thread> pool;
  const size_t per = ((bytes + threads - 1) / threads + 63) & ~size_t(63);
  for(int t=0; t < threads; ++t)
  {
    const size_t b0 = (size_t)t*per; if(b0 >= bytes) break;
    const size_t bn = std::min(per, bytes - b0);
    pool.emplace_back([=] { for(int r=0; r < rows; ++r) std::memcpy(dst + r*dst_pitch + b0, src + r*src_pitch + b0, bn); });
  }
  for(auto& th : pool) th.join();
}

__global__ void k_narrow_to_bytes(const int32_t* __restrict__ in, unsigned char* __restrict__ out, size_t n)
{
  for(size_t i = ((size_t)blockIdx.x*blockDim.x + threadIdx.x)*4; i < n; i += (size_t)gridDim.x*blockDim.x*4)
  {
    if(i + 4 <= n) { const int4 v = *reinterpret_cast<const int4*>(in + i); *reinterpret_cast<uchar4*>(out + i) = make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z, (unsigned char)v.w); }
    else for(size_t k = i; k < n; ++k) out[k] = (unsigned char)in[k];
  }
}

// A few persistent host threads for the per-chunk work of the host-pointer path (created on first use; creating threads
// per call cost more than the work of a chunk).  run(n, f) calls f(begin, end) on disjoint ranges covering [0, n).
class HostPool
{
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::function<void(size_t, size_t)> fn;
  size_t n = 0, per = 0;
  int pending = 0;
  uint64_t generation = 0;
  bool stop = false;
  void loop(int id)
  {
    uint64_t seen = 0;
    for(;;)
    {
      std::unique_lock<std::mutex> lk(m);
      cv_work.wait(lk, [&] { return stop || generation != seen; });
      if(stop) return;
      seen = generation;
      const size_t b = (size_t)id*per, e = std::min(n, b + per);
      lk.unlock();
      if(b < e) fn(b, e);
      lk.lock();
      if(--pending == 0) cv_done.notify_one();
    }
  }
public:
  explicit HostPool(int threads) { for(int i=0; i < threads; ++i) workers.emplace_back([this, i] { loop(i); }); }
  ~HostPool() { { std::lock_guard<std::mutex> lk(m); stop = true; } cv_work.notify_all(); for(auto& t : workers) t.join(); }
  int size() const { return (int)workers.size(); }
  void run(size_t count, std::function<void(size_t, size_t)> f)
  {
    std::unique_lock<std::mutex> lk(m);
    fn = std::move(f); n = count;
    per = ((count + workers.size() - 1) / workers.size() + 63) & ~size_t(63);
    pending = (int)workers.size(); ++generation;
    cv_work.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};
std::mutex g_pool_mutex;                                        // one call at a time uses the pool
HostPool& host_pool()
{
  // half the hardware threads, at most 16 (BBMCU_HOST_THREADS overrides: the loops are bound by the host's memory system)
  static HostPool pool([] { const char* e = std::getenv("BBMCU_HOST_THREADS"); const int v = e ? std::atoi(e) : 0;
                            return v > 0 ? std::min(v, 64) : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / 2)); }());
  return pool;
}
// f(begin, end) over [0, n) on the pool's threads (small n: on the caller)
template<class F> void host_parallel(size_t n, int threads, F&& f)
{
  if(n < (size_t(1) << 18) || threads <= 1) { f((size_t)0, n); return; }
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  host_pool().run(n, std::function<void(size_t, size_t)>(f));
}

// Streaming (non-temporal) forms of the per-chunk host loops: the destinations are whole output planes that are not
// read again here, so the stores bypass the cache instead of reading every line for ownership first.
inline bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
void widen_bytes(int32_t* dst, const unsigned char* src, size_t b, size_t e)
{
  size_t i = b;
  while(i < e && !aligned16p(dst + i)) { dst[i] = (int32_t)src[i]; ++i; }
  const __m128i z = _mm_setzero_si128();
  for(; i + 16 <= e; i += 16)
  {
    const __m128i x = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i lo = _mm_unpacklo_epi8(x, z), hi = _mm_unpackhi_epi8(x, z);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i),      _mm_unpacklo_epi16(lo, z));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4),  _mm_unpackhi_epi16(lo, z));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 8),  _mm_unpacklo_epi16(hi, z));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 12), _mm_unpackhi_epi16(hi, z));
  }
  for(; i < e; ++i) dst[i] = (int32_t)src[i];
  _mm_sfence();
}
void stream_copy(float* dst, const float* src, size_t b, size_t e)
{
  size_t i = b;
  while(i < e && !aligned16p(dst + i)) { dst[i] = src[i]; ++i; }
  for(; i + 4 <= e; i += 4) _mm_stream_ps(dst + i, _mm_loadu_ps(src + i));
  for(; i < e; ++i) dst[i] = src[i];
  _mm_sfence();
}
// d_c[i] = u[i] * s_c: IEEE single products, scalar or packed (the same operation, the same bits as the device's FMUL)
void stream_scaled3(float* d0, float* d1, float* d2, const float* u, float s0, float s1, float s2, size_t b, size_t e)
{
  size_t i = b;
  if(aligned16p(d0) && aligned16p(d1) && aligned16p(d2))
  {
    while(i < e && (i & 3)) { const float v = u[i]; d0[i] = v*s0; d1[i] = v*s1; d2[i] = v*s2; ++i; }
    const __m128 v0 = _mm_set1_ps(s0), v1 = _mm_set1_ps(s1), v2 = _mm_set1_ps(s2);
    for(; i + 4 <= e; i += 4)
    {
      const __m128 v = _mm_loadu_ps(u + i);
      _mm_stream_ps(d0 + i, _mm_mul_ps(v, v0)); _mm_stream_ps(d1 + i, _mm_mul_ps(v, v1)); _mm_stream_ps(d2 + i, _mm_mul_ps(v, v2));
    }
  }
  for(; i < e; ++i) { const float v = u[i]; d0[i] = v*s0; d1[i] = v*s1; d2[i] = v*s2; }
  _mm_sfence();
}

// Host-pointer path: chunks of the batch flow through kSlots device staging buffers, each chunk's H2D copies, kernel and
// D2H copies on its own stream so copies of one chunk overlap the kernel and the copies of its neighbours.  Pinned /
// registered caller memory is DMA'd in place; PAGEABLE caller memory goes through a pinned staging ring filled and
// drained by a few host threads (cudaMemcpyAsync from pageable memory would serialise the whole pipeline).  The staging
// planes are padded to a multiple of 4 floats, so the kernels keep their 16-byte accesses for any n.
// `launch(stream, device_pointers, chunk_n)` issues the kernel; device pointers of null outputs are null.
template<class Launch>
void run_hosted(bbmcu_ctx* ctx, size_t n, const std::vector<ArrayArg>& args, const std::vector<Mem>& kind, Launch&& launch)
{
  if(n == 0) return;
  const size_t cap = std::min(n, kChunk);
  const size_t ld = (cap + 3) & ~size_t(3);                    // staging plane stride
  const size_t user_ld = ctx->user_ld ? ctx->user_ld : n;      // the caller's plane stride
  if(user_ld < n) throw std::invalid_argument("BBM: plane stride smaller than the batch");
  size_t need = 0;
  bool any_pageable = false;                                    // per-chunk host work: pageable memory, narrowed or derived outputs
  std::vector<size_t> off(args.size());
  for(size_t a=0; a < args.size(); ++a)
  {
    off[a] = need;
    if(!args[a].ptr) continue;
    need += ((size_t)args[a].planes * ld * 4 + 255) & ~size_t(255);
    if(args[a].xfer == XFER_BYTES) need += (ld + 255) & ~size_t(255);             // the narrowed copy behind the int32 plane
    any_pageable |= (kind[a] == Mem::Pageable) || (args[a].xfer != XFER_PLAIN);
  }
  if(need > ctx->slot_bytes)
  {
    ctx->slot_bytes = 0;                                   // stays 0 if an allocation below throws: the next call re-allocates
    for(int s=0; s < bbmcu_ctx::kSlots; ++s) { if(ctx->slot_buf[s]) BBMCU_CUDA(cudaFree(ctx->slot_buf[s])); ctx->slot_buf[s] = nullptr; }
    for(int s=0; s < bbmcu_ctx::kSlots; ++s) BBMCU_CUDA(cudaMalloc(&ctx->slot_buf[s], need));
    ctx->slot_bytes = need;
  }
  if(any_pageable && need > ctx->pin_bytes)
  {
    ctx->pin_bytes = 0;
    for(int s=0; s < bbmcu_ctx::kSlots; ++s) { if(ctx->pin_buf[s]) BBMCU_CUDA(cudaFreeHost(ctx->pin_buf[s])); ctx->pin_buf[s] = nullptr; }
    for(int s=0; s < bbmcu_ctx::kSlots; ++s) BBMCU_CUDA(cudaMallocHost(&ctx->pin_buf[s], need));
    ctx->pin_bytes = need;
  }
  const int threads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
  BBMCU_CUDA(cudaStreamSynchronize(ctx->stream));          // order after earlier device-pointer work
  struct Pending { bool busy = false; size_t c0 = 0, cn = 0; } pending[bbmcu_ctx::kSlots];
  // wait for the chunk that last used slot s and hand its pageable outputs to the caller
  auto drain = [&](int s) {
    if(!pending[s].busy) return;
    BBMCU_CUDA(cudaStreamSynchronize(ctx->slot_stream[s]));
    pending[s].busy = false;
    const size_t c0 = pending[s].c0, cn = pending[s].cn;
    for(size_t a=0; a < args.size(); ++a)
      if(args[a].ptr && !args[a].input && args[a].xfer == XFER_PLAIN && kind[a] == Mem::Pageable)
        host_copy_rows((char*)args[a].ptr + c0*4, user_ld*4, (const char*)ctx->pin_buf[s] + off[a], ld*4, cn*4, args[a].planes, threads);
    // narrowed outputs: bytes in the ring -> the caller's int32 plane
    for(size_t a=0; a < args.size(); ++a)
      if(args[a].ptr && args[a].xfer == XFER_BYTES)
      {
        const unsigned char* src = (const unsigned char*)ctx->pin_buf[s] + off[a];
        int32_t* dst = (int32_t*)args[a].ptr + c0;
        host_parallel(cn, threads, [=](size_t b, size_t e) { widen_bytes(dst, src, b, e); });
      }
    // scaled outputs: u in the ring -> u * scale[c] in the caller's three planes
    for(size_t a=0; a < args.size(); ++a)
      if(args[a].ptr && args[a].xfer == XFER_GRAY_SCALED)
      {
        const float* src = (const float*)((const char*)ctx->pin_buf[s] + off[a]);
        float* d0 = (float*)args[a].ptr + c0; float* d1 = d0 + user_ld; float* d2 = d1 + user_ld;
        const float s0 = args[a].scale[0], s1 = args[a].scale[1], s2 = args[a].scale[2];
        host_parallel(cn, threads, [=](size_t b, size_t e) { stream_scaled3(d0, d1, d2, src, s0, s1, s2, b, e); });
      }
    // derived outputs, from what has just arrived of their sources (the caller's own plane, or the ring if that is pageable and not yet copied... it is: see above)
    for(size_t a=0; a < args.size(); ++a)
      if(args[a].ptr && (args[a].xfer == XFER_COPY_OF || args[a].xfer == XFER_MASKED_BY))
      {
        const float* src = (const float*)args[args[a].src].ptr + c0;
        float* dst = (float*)args[a].ptr + c0;
        if(args[a].xfer == XFER_COPY_OF) host_parallel(cn, threads, [=](size_t b, size_t e) { stream_copy(dst, src, b, e); });
        else
        {
          const unsigned char* m = (const unsigned char*)ctx->pin_buf[s] + off[args[a].mask];     // the flag bytes of this chunk
          host_parallel(cn, threads, [=](size_t b, size_t e) { for(size_t i = b; i < e; ++i) dst[i] = m[i] ? src[i] : 0.0f; });
        }
      }
  };
  const size_t saved_ld = ctx->ld;
  ctx->ld = ld;
  try
  {
    size_t chunk_id = 0;
    for(size_t c0 = 0; c0 < n; c0 += cap, ++chunk_id)
    {
      const size_t cn = std::min(cap, n - c0);
      const int s = (int)(chunk_id % bbmcu_ctx::kSlots);
      cudaStream_t st = ctx->slot_stream[s];                  // in-order reuse of the slot buffer
      if(any_pageable) drain(s);
      std::vector<void*> dptr(args.size(), nullptr);
      for(size_t a=0; a < args.size(); ++a)
      {
        if(!args[a].ptr) continue;
        dptr[a] = (char*)ctx->slot_buf[s] + off[a];
        if(!args[a].input) continue;
        const char* src = (const char*)args[a].ptr + c0*4;
        if(kind[a] == Mem::Pageable)
        {
          host_copy_rows((char*)ctx->pin_buf[s] + off[a], ld*4, src, user_ld*4, cn*4, args[a].planes, threads);
          BBMCU_CUDA(cudaMemcpy2DAsync(dptr[a], ld*4, (const char*)ctx->pin_buf[s] + off[a], ld*4, cn*4, args[a].planes, cudaMemcpyHostToDevice, st));
        }
        else BBMCU_CUDA(cudaMemcpy2DAsync(dptr[a], ld*4, src, user_ld*4, cn*4, args[a].planes, cudaMemcpyHostToDevice, st));
      }
      // a derived output is still computed by the kernel when its staging pointer is given; it is not: null = not stored
      for(size_t a=0; a < args.size(); ++a) if(args[a].ptr && (args[a].xfer == XFER_COPY_OF || args[a].xfer == XFER_MASKED_BY)) dptr[a] = nullptr;
      launch(st, dptr, cn);
      for(size_t a=0; a < args.size(); ++a)
      {
        if(!args[a].ptr || args[a].input) continue;
        if(args[a].xfer == XFER_COPY_OF || args[a].xfer == XFER_MASKED_BY) continue;
        if(args[a].xfer == XFER_BYTES)
        {
          unsigned char* bytes = (unsigned char*)dptr[a] + (((size_t)args[a].planes * ld * 4 + 255) & ~size_t(255));      // behind the int32 plane
          k_narrow_to_bytes<<<(unsigned)std::min<size_t>(1184, (cn/4 + 255)/256 + 1), 256, 0, st>>>((const int32_t*)dptr[a], bytes, cn);
          BBMCU_CUDA(cudaGetLastError());
          BBMCU_CUDA(cudaMemcpyAsync((char*)ctx->pin_buf[s] + off[a], bytes, cn, cudaMemcpyDeviceToHost, st));
          continue;
        }
        if(args[a].xfer == XFER_GRAY_SCALED) { BBMCU_CUDA(cudaMemcpyAsync((char*)ctx->pin_buf[s] + off[a], dptr[a], cn*4, cudaMemcpyDeviceToHost, st)); continue; }
        if(kind[a] == Mem::Pageable) BBMCU_CUDA(cudaMemcpy2DAsync((char*)ctx->pin_buf[s] + off[a], ld*4, dptr[a], ld*4, cn*4, args[a].planes, cudaMemcpyDeviceToHost, st));
        else BBMCU_CUDA(cudaMemcpy2DAsync((char*)args[a].ptr + c0*4, user_ld*4, dptr[a], ld*4, cn*4, args[a].planes, cudaMemcpyDeviceToHost, st));
      }
      pending[s].busy = true; pending[s].c0 = c0; pending[s].cn = cn;
    }
    for(int s=0; s < bbmcu_ctx::kSlots; ++s) { if(any_pageable) drain(s); else BBMCU_CUDA(cudaStreamSynchronize(ctx->slot_stream[s])); }
  }
  catch(...)
  {
    // copies into the caller's memory may still be in flight: do not return before they have landed
    for(int s=0; s < bbmcu_ctx::kSlots; ++s) cudaStreamSynchronize(ctx->slot_stream[s]);
    ctx->ld = saved_ld;
    throw;
  }
  ctx->ld = saved_ld;
}

template<class Launch>
void run_any(bbmcu_ctx* ctx, size_t n, const std::vector<ArrayArg>& args, Launch&& launch)
{
  if(!ctx) throw std::invalid_argument("BBM: null context");
  BBMCU_CUDA(cudaSetDevice(ctx->device));
  std::vector<Mem> kind;
  if(all_device(args, kind))
  {
    std::vector<void*> p(args.size());
    for(size_t a=0; a < args.size(); ++a) p[a] = const_cast<void*>(args[a].ptr);
    if(ctx->user_ld && ctx->user_ld < n) throw std::invalid_argument("BBM: plane stride smaller than the batch");
    const size_t saved_ld = ctx->ld;
    ctx->ld = ctx->user_ld;
    try { launch(ctx->stream, p, n); } catch(...) { ctx->ld = saved_ld; throw; }
    ctx->ld = saved_ld;
  }
  else run_hosted(ctx, n, args, kind, launch);
}

// What the fused pass returns twice: sample.pdf against pdf(sample.direction, out).
//   an aggregate:                         sample.pdf IS that pdf (aggregatebsdf.h:119-126 and :183; SampleEvalPdfOpT copies it)
//   the hand-merged GGX kernel:           the same number (0 when the sample is invalid)
//   models with kSamplePdfIsPdf:          pdf where the sample's flag is not None, else 0
// For those the host-pointer path does not send sample.pdf over the link: host threads write it from pdf (and the flag).
// The flag itself travels as one byte per element.  BBMCU_HOST_TRANSFER_PLAIN=1 (tests) sends every plane as it is.
enum : int { SPDF_INDEPENDENT = 0, SPDF_IS_PDF, SPDF_IS_PDF_WHERE_FLAGGED };
int sample_pdf_rule(const BsdfDesc& d)
{
  if(d.aggregate) return SPDF_IS_PDF;
  if(d.n_lobes != 1) return SPDF_INDEPENDENT;
  int rule = SPDF_INDEPENDENT;
  dispatch_model_host(d.model[0], [&](auto* m) {
    using M = typename std::remove_pointer<decltype(m)>::type;
    if(HandFused<M>::value) rule = SPDF_IS_PDF;
    else if(SamplePdfIsPdf<M>::value) rule = SPDF_IS_PDF_WHERE_FLAGGED;
  });
  return rule;
}
void compress_fused_outputs(const BsdfDesc& d, std::vector<ArrayArg>& args, int i_spdf, int i_flag, int i_pdf, int i_rgb = -1)
{
  static const bool plain = [] { const char* e = std::getenv("BBMCU_HOST_TRANSFER_PLAIN"); return e && e[0] == '1'; }();
  if(plain) return;
  // eval of a single hand-merged lobe = u * its leading RGB scale (attributes 0..2): send u.  Host pointers only - the
  // caller of a device-pointer call reads the planes the kernel wrote.
  if(i_rgb >= 0 && args[i_rgb].ptr && classify(args[i_rgb].ptr) != Mem::Device && launch_sample_eval_pdf_gray_capable(d))
  { args[i_rgb].xfer = XFER_GRAY_SCALED; args[i_rgb].scale[0] = d.attrs[0]; args[i_rgb].scale[1] = d.attrs[1]; args[i_rgb].scale[2] = d.attrs[2]; }
  if(args[i_flag].ptr) args[i_flag].xfer = XFER_BYTES;
  if(!args[i_spdf].ptr || !args[i_pdf].ptr) return;
  const int rule = sample_pdf_rule(d);
  if(rule == SPDF_IS_PDF) { args[i_spdf].xfer = XFER_COPY_OF; args[i_spdf].src = i_pdf; }
  else if(rule == SPDF_IS_PDF_WHERE_FLAGGED && args[i_flag].ptr) { args[i_spdf].xfer = XFER_MASKED_BY; args[i_spdf].src = i_pdf; args[i_spdf].mask = i_flag; }
}

void check_flags(int component, int unit)
{
  if(component < 0 || component > 3) throw std::invalid_argument("BBM: invalid bsdf_flag " + std::to_string(component));
  if(unit != BBMCU_RADIANCE && unit != BBMCU_IMPORTANCE) throw std::invalid_argument("BBM: invalid unit " + std::to_string(unit));
}

} // anonymous namespace
} // namespace bbmcu

extern "C" {

// ---- context ------------------------------------------------------------------------------------------
int bbmcu_init(int device, bbmcu_ctx** out)
{
  return guarded(nullptr, [&] {
    if(!out) throw std::invalid_argument("BBM: null output pointer");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if(e != cudaSuccess || count == 0) throw CudaError(std::string("BBM: no usable CUDA device (") + cudaGetErrorString(e) + "); the CUDA backbone has no CPU fallback");
    if(device < 0 || device >= count) throw std::invalid_argument("BBM: device " + std::to_string(device) + " out of range (" + std::to_string(count) + " devices)");
    BBMCU_CUDA(cudaSetDevice(device));
    std::unique_ptr<bbmcu_ctx> ctx(new bbmcu_ctx);
    ctx->device = device;
    cudaDeviceProp prop;
    BBMCU_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    BBMCU_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for(int s=0; s < bbmcu_ctx::kSlots; ++s) BBMCU_CUDA(cudaStreamCreateWithFlags(&ctx->slot_stream[s], cudaStreamNonBlocking));
    *out = ctx.release();
  });
}

void bbmcu_destroy(bbmcu_ctx* ctx)
{
  if(!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for(int s=0; s < bbmcu_ctx::kSlots; ++s) { if(ctx->slot_buf[s]) cudaFree(ctx->slot_buf[s]); if(ctx->pin_buf[s]) cudaFreeHost(ctx->pin_buf[s]); if(ctx->slot_stream[s]) cudaStreamDestroy(ctx->slot_stream[s]); }
  if(ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* bbmcu_last_error(bbmcu_ctx* ctx) { return ctx ? ctx->error.c_str() : thread_error(); }
int bbmcu_synchronize(bbmcu_ctx* ctx) { return guarded(ctx, [&] { if(!ctx) throw std::invalid_argument("BBM: null context"); BBMCU_CUDA(cudaStreamSynchronize(ctx->stream)); }); }
void* bbmcu_stream(bbmcu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t bbmcu_launch_count(bbmcu_ctx* ctx) { return ctx ? ctx->launches : 0; }
int bbmcu_set_plane_stride(bbmcu_ctx* ctx, size_t ld) { return guarded(ctx, [&] { if(!ctx) throw std::invalid_argument("BBM: null context"); ctx->user_ld = ld; }); }
int bbmcu_host_register(bbmcu_ctx* ctx, void* ptr, size_t bytes)
{
  return guarded(ctx, [&] {
    if(!ctx || !ptr || !bytes) throw std::invalid_argument("BBM: null argument");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BBMCU_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  });
}
int bbmcu_host_unregister(bbmcu_ctx* ctx, void* ptr)
{
  return guarded(ctx, [&] { if(!ctx || !ptr) throw std::invalid_argument("BBM: null argument"); BBMCU_CUDA(cudaSetDevice(ctx->device)); BBMCU_CUDA(cudaHostUnregister(ptr)); });
}

// ---- registry -----------------------------------------------------------------------------------------
int bbmcu_model_count(void) { return (int)bbmcu_host::model_table().size(); }
const char* bbmcu_model_name(int id)
{
  auto& t = bbmcu_host::model_table();
  return (id >= 0 && id < (int)t.size()) ? t[id].name.c_str() : nullptr;
}
int bbmcu_model_lookup(const char* name, int* id)
{
  return guarded(nullptr, [&] {
    auto* m = name ? bbmcu_host::find_model(name) : nullptr;
    if(!m) throw std::invalid_argument(std::string("BBM: unrecognized BSDF model: ") + (name ? name : "(null)"));
    if(id) *id = m->id;
  });
}
int bbmcu_model_layout(int id, bbmcu_attr* attrs, int* n_attrs)
{
  return guarded(nullptr, [&] {
    auto& t = bbmcu_host::model_table();
    if(id < 0 || id >= (int)t.size()) throw std::invalid_argument("BBM: invalid model id");
    int off = 0, k = 0;
    for(auto& a : t[id].attrs)
    {
      if(attrs) { attrs[k].name = a.name.c_str(); attrs[k].width = a.width; attrs[k].rows = a.rows; attrs[k].flag = a.flag; attrs[k].offset = off; }
      off += a.width; ++k;
    }
    if(n_attrs) *n_attrs = k;
  });
}

// ---- BSDF objects -------------------------------------------------------------------------------------
int bbmcu_bsdf_from_string(bbmcu_ctx* ctx, const char* str, bbmcu_bsdf** out) { return bbmcu_bsdf_from_string_ex(ctx, str, BBMCU_FLOAT_RGB, out); }
int bbmcu_bsdf_from_string_ex(bbmcu_ctx* ctx, const char* str, int config, bbmcu_bsdf** out)
{
  return guarded(ctx, [&] {
    if(!str || !out) throw std::invalid_argument("BBM: null argument");
    if(config != BBMCU_FLOAT_RGB && config != BBMCU_DOUBLE_RGB) throw std::invalid_argument("BBM: unknown configuration");
    std::unique_ptr<bbmcu_bsdf> b(new bbmcu_bsdf);
    b->b = bbmcu_host::parse_bsdf(str, config == BBMCU_DOUBLE_RGB);
    make_desc(b->b, -1);                                // validates lobe/attribute limits early (no device touched)
    *out = b.release();
  });
}
void bbmcu_bsdf_free(bbmcu_bsdf* b) { delete b; }
int bbmcu_bsdf_to_string(const bbmcu_bsdf* b, char* buf, size_t cap)
{
  return guarded(nullptr, [&] {
    if(!b || !buf || !cap) throw std::invalid_argument("BBM: null argument");
    std::string s = b->b.to_string();
    if(s.size() + 1 > cap) throw std::out_of_range("BBM: buffer too small for BSDF string (" + std::to_string(s.size() + 1) + " bytes needed)");
    std::memcpy(buf, s.c_str(), s.size() + 1);
  });
}
int bbmcu_bsdf_param_count(const bbmcu_bsdf* b, int flags) { return b ? b->b.param_count(flags) : -1; }
int bbmcu_bsdf_get_params(const bbmcu_bsdf* b, int which, int flags, double* values, int* count)
{
  return guarded(nullptr, [&] {
    if(!b) throw std::invalid_argument("BBM: null bsdf");
    if(which < 0 || which > 3) throw std::invalid_argument("BBM: invalid parameter vector selector");
    auto v = b->b.params(which, flags);
    if(values) std::copy(v.begin(), v.end(), values);
    if(count) *count = (int)v.size();
  });
}
int bbmcu_bsdf_set_params(bbmcu_bsdf* b, int flags, const double* values, int count)
{
  return guarded(nullptr, [&] {
    if(!b || !values) throw std::invalid_argument("BBM: null argument");
    b->b.set_params(flags, values, count);
  });
}

// ---- batched BSDF concept -----------------------------------------------------------------------------
int bbmcu_eval(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit, const float* in, const float* out, size_t n, float* rgb)
{
  return guarded(ctx, [&] {
    if(!bsdf) throw std::invalid_argument("BBM: null bsdf");
    check_flags(component, unit);
    if(n == 0) return;
    if(!ctx) throw std::invalid_argument("BBM: null context");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc d = make_desc(bsdf->b, ctx->device);
    run_any(ctx, n, {{in, 3, true}, {out, 3, true}, {rgb, 3, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      launch_eval(ctx, s, d, component, (const float*)p[0], (const float*)p[1], (float*)p[2], cn); });
  });
}

int bbmcu_pdf(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit, const float* in, const float* out, size_t n, float* pdf)
{
  return guarded(ctx, [&] {
    if(!bsdf) throw std::invalid_argument("BBM: null bsdf");
    check_flags(component, unit);
    if(n == 0) return;
    if(!ctx) throw std::invalid_argument("BBM: null context");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc d = make_desc(bsdf->b, ctx->device);
    run_any(ctx, n, {{in, 3, true}, {out, 3, true}, {pdf, 1, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      launch_pdf(ctx, s, d, component, (const float*)p[0], (const float*)p[1], (float*)p[2], cn); });
  });
}

int bbmcu_reflectance(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit, const float* out, size_t n, float* rgb)
{
  return guarded(ctx, [&] {
    if(!bsdf) throw std::invalid_argument("BBM: null bsdf");
    check_flags(component, unit);
    if(n == 0) return;
    if(!ctx) throw std::invalid_argument("BBM: null context");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc d = make_desc(bsdf->b, ctx->device);
    run_any(ctx, n, {{out, 3, true}, {rgb, 3, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      launch_reflectance(ctx, s, d, component, (const float*)p[0], (float*)p[1], cn); });
  });
}

int bbmcu_sample(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit, const float* out, const float* xi, size_t n, float* dir, float* pdf, int32_t* flag)
{
  return guarded(ctx, [&] {
    if(!bsdf) throw std::invalid_argument("BBM: null bsdf");
    check_flags(component, unit);
    if(n == 0) return;
    if(!ctx) throw std::invalid_argument("BBM: null context");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc d = make_desc(bsdf->b, ctx->device);
    run_any(ctx, n, {{out, 3, true}, {xi, 2, true}, {dir, 3, false}, {pdf, 1, false}, {flag, 1, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      launch_sample(ctx, s, d, component, (const float*)p[0], (const float*)p[1], (float*)p[2], (float*)p[3], (int32_t*)p[4], cn); });
  });
}

int bbmcu_sample_eval_pdf(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit, const float* out, const float* xi, size_t n,
                          float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf)
{
  return guarded(ctx, [&] {
    if(!bsdf) throw std::invalid_argument("BBM: null bsdf");
    check_flags(component, unit);
    if(n == 0) return;
    if(!ctx) throw std::invalid_argument("BBM: null context");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc d = make_desc(bsdf->b, ctx->device);
    std::vector<ArrayArg> args = {{out, 3, true}, {xi, 2, true}, {dir, 3, false}, {spdf, 1, false}, {flag, 1, false}, {rgb, 3, false}, {pdf, 1, false}};
    compress_fused_outputs(d, args, 3, 4, 6, 5);
    const bool gray = args[5].xfer == XFER_GRAY_SCALED;
    run_any(ctx, n, args, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      if(gray && launch_sample_eval_pdf_gray(ctx, s, d, component, (const float*)p[0], (const float*)p[1], (float*)p[2], (float*)p[3], (int32_t*)p[4], (float*)p[5], (float*)p[6], cn)) return;
      launch_sample_eval_pdf(ctx, s, d, component, (const float*)p[0], (const float*)p[1], (float*)p[2], (float*)p[3], (int32_t*)p[4], (float*)p[5], (float*)p[6], cn); });
  });
}

int bbmcu_sample_eval_pdf_generated(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit, uint64_t seed, uint64_t first, size_t n,
                                    float* gen_out, float* gen_xi, float* dir, float* spdf, int32_t* flag, float* rgb, float* pdf)
{
  return guarded(ctx, [&] {
    if(!bsdf) throw std::invalid_argument("BBM: null bsdf");
    check_flags(component, unit);
    if(n == 0) return;
    if(!ctx) throw std::invalid_argument("BBM: null context");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc d = make_desc(bsdf->b, ctx->device);
    size_t done = 0;                                     // chunks arrive in order: the running element offset
    std::vector<ArrayArg> args = {{gen_out, 3, false}, {gen_xi, 2, false}, {dir, 3, false}, {spdf, 1, false}, {flag, 1, false}, {rgb, 3, false}, {pdf, 1, false}};
    compress_fused_outputs(d, args, 3, 4, 6);
    run_any(ctx, n, args, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      GenArgs g; g.gen = 1; g.seed = seed; g.first = first + done; g.out = (float*)p[0]; g.xi = (float*)p[1];
      launch_sample_eval_pdf(ctx, s, d, component, nullptr, nullptr, (float*)p[2], (float*)p[3], (int32_t*)p[4], (float*)p[5], (float*)p[6], cn, g);
      done += cn; });
  });
}

int bbmcu_eval_merl_grid(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit, uint32_t first, size_t n, float* rgb, float* in, float* out)
{
  return guarded(ctx, [&] {
    if(!bsdf) throw std::invalid_argument("BBM: null bsdf");
    check_flags(component, unit);
    if(n == 0) return;
    if(!ctx) throw std::invalid_argument("BBM: null context");
    if((uint64_t)first + n > kMerlBins) throw std::out_of_range("BBM: bins beyond the MERL grid");
    if(!rgb) throw std::invalid_argument("BBM: null output pointer");
    BBMCU_CUDA(cudaSetDevice(ctx->device));
    BsdfDesc d = make_desc(bsdf->b, ctx->device);
    size_t done = 0;
    run_any(ctx, n, {{rgb, 3, false}, {in, 3, false}, {out, 3, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      launch_eval_grid(ctx, s, d, component, first + (uint32_t)done, (float*)p[0], (float*)p[1], (float*)p[2], cn); done += cn; });
  });
}

// ---- linearizers ----------------------------------------------------------------------------------------
int bbmcu_merl_index(bbmcu_ctx* ctx, const float* in, const float* out, size_t n, uint32_t* index)
{
  return guarded(ctx, [&] {
    if(n == 0) return;
    run_any(ctx, n, {{in, 3, true}, {out, 3, true}, {index, 1, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      MerlIndexOp op; op.in = (const float*)p[0]; op.out = (const float*)p[1]; op.index = (uint32_t*)p[2]; op.n = cn;
      op.aligned = aligned16(p[0]) && aligned16(p[1]) && aligned16(p[2]);
      launch_foreach4(ctx, s, op, cn); });
  });
}

int bbmcu_merl_dirs(bbmcu_ctx* ctx, uint32_t first, size_t n, float* in, float* out)
{
  return guarded(ctx, [&] {
    if(n == 0) return;
    size_t done = 0;     // run_any hands out consecutive chunks; track the running first index
    run_any(ctx, n, {{in, 3, false}, {out, 3, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      MerlDirsOp op; op.first = first + (uint32_t)done; op.in = (float*)p[0]; op.out = (float*)p[1]; op.n = cn;
      op.aligned = aligned16(p[0]) && aligned16(p[1]);
      launch_foreach4(ctx, s, op, cn); done += cn; });
  });
}

void bbmcu_spherical_grid_default(bbmcu_spherical_grid* g, uint32_t in_phi, uint32_t in_theta, uint32_t out_phi, uint32_t out_theta)
{
  if(!g) return;
  g->samples_in[0] = in_phi; g->samples_in[1] = in_theta; g->samples_out[0] = out_phi; g->samples_out[1] = out_theta;
  g->start_in[0] = g->start_in[1] = g->start_out[0] = g->start_out[1] = 0.0f;
  g->end_in[0] = g->end_out[0] = kTwoPi; g->end_in[1] = g->end_out[1] = kHalfPi;    // Constants::Hemisphere()
}

} // extern "C"

namespace bbmcu {
SphericalGrid to_device_grid(const bbmcu_spherical_grid& g)
{
  if(!g.samples_in[0] || !g.samples_in[1] || !g.samples_out[0] || !g.samples_out[1]) throw std::invalid_argument("BBM: spherical grid with zero samples");
  SphericalGrid d;
  d.n_in_phi = g.samples_in[0]; d.n_in_theta = g.samples_in[1]; d.n_out_phi = g.samples_out[0]; d.n_out_theta = g.samples_out[1];
  d.start_in_phi = g.start_in[0]; d.start_in_theta = g.start_in[1]; d.start_out_phi = g.start_out[0]; d.start_out_theta = g.start_out[1];
  d.size_in_phi = g.end_in[0] - g.start_in[0]; d.size_in_theta = g.end_in[1] - g.start_in[1];
  d.size_out_phi = g.end_out[0] - g.start_out[0]; d.size_out_theta = g.end_out[1] - g.start_out[1];
  return d;
}
}

extern "C" {

int bbmcu_spherical_dirs(bbmcu_ctx* ctx, const bbmcu_spherical_grid* grid, uint64_t first, size_t n, float* in, float* out)
{
  return guarded(ctx, [&] {
    if(!grid) throw std::invalid_argument("BBM: null grid");
    if(n == 0) return;
    SphericalGrid g = to_device_grid(*grid);
    size_t done = 0;
    run_any(ctx, n, {{in, 3, false}, {out, 3, false}}, [&](cudaStream_t s, const std::vector<void*>& p, size_t cn) {
      SphericalDirsOp op; op.grid = g; op.first = first + done; op.in = (float*)p[0]; op.out = (float*)p[1]; op.n = cn;
      op.aligned = aligned16(p[0]) && aligned16(p[1]);
      launch_foreach4(ctx, s, op, cn); done += cn; });
  });
}

// ---- MERL binaries (staticmodel/merl.h:173-206) ------------------------------------------------------------
int bbmcu_merl_read(bbmcu_ctx* ctx, const char* filename, float* rgb)
{
  return guarded(ctx, [&] {
    if(!filename || !rgb) throw std::invalid_argument("BBM: null argument");
    std::vector<float> t = bbmcu_host::read_merl(filename);
    std::memcpy(rgb, t.data(), t.size()*sizeof(float));
  });
}

int bbmcu_merl_write(bbmcu_ctx* ctx, const char* filename, const float* rgb)
{
  return guarded(ctx, [&] {
    if(!filename || !rgb) throw std::invalid_argument("BBM: null argument");
    bbmcu_host::write_merl(filename, rgb);
  });
}

// ---- .fit files ------------------------------------------------------------------------------------------------
int bbmcu_fit_import(bbmcu_ctx* ctx, const char* filename, bbmcu_fit** out) { return bbmcu_fit_import_ex(ctx, filename, BBMCU_FLOAT_RGB, out); }
int bbmcu_fit_import_ex(bbmcu_ctx* ctx, const char* filename, int config, bbmcu_fit** out)
{
  return guarded(ctx, [&] {
    if(!filename || !out) throw std::invalid_argument("BBM: null argument");
    if(config != BBMCU_FLOAT_RGB && config != BBMCU_DOUBLE_RGB) throw std::invalid_argument("BBM: unknown configuration");
    std::unique_ptr<bbmcu_fit> f(new bbmcu_fit);
    f->entries = bbmcu_host::import_fit(filename, config == BBMCU_DOUBLE_RGB);
    *out = f.release();
  });
}
int bbmcu_fit_count(const bbmcu_fit* f) { return f ? (int)f->entries.size() : -1; }
const char* bbmcu_fit_key(const bbmcu_fit* f, int i) { return (f && i >= 0 && i < (int)f->entries.size()) ? f->entries[i].first.c_str() : nullptr; }
int bbmcu_fit_bsdf(const bbmcu_fit* f, int i, bbmcu_bsdf** out)
{
  return guarded(nullptr, [&] {
    if(!f || !out || i < 0 || i >= (int)f->entries.size()) throw std::invalid_argument("BBM: invalid fit entry");
    *out = new bbmcu_bsdf{f->entries[i].second};
  });
}
int bbmcu_fit_create(bbmcu_fit** out) { return guarded(nullptr, [&] { if(!out) throw std::invalid_argument("BBM: null argument"); *out = new bbmcu_fit; }); }
int bbmcu_fit_add(bbmcu_fit* f, const char* key, const bbmcu_bsdf* b)
{
  return guarded(nullptr, [&] { if(!f || !key || !b) throw std::invalid_argument("BBM: null argument"); f->entries.emplace_back(key, b->b); });
}
int bbmcu_fit_export(bbmcu_ctx* ctx, const bbmcu_fit* f, const char* filename, const char* comment)
{
  return guarded(ctx, [&] { if(!f || !filename) throw std::invalid_argument("BBM: null argument"); bbmcu_host::export_fit(filename, f->entries, comment ? comment : ""); });
}
void bbmcu_fit_free(bbmcu_fit* f) { delete f; }

} // extern "C"
