// Internal: context, object structs and small helpers shared by the .cu translation units of libbbmcu.so
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/bbmcu.h"
#include "bbmcu_host.hpp"
#include "bbmcu_bsdf.cuh"
#include "bbmcu_desc.hpp"

struct bbmcu_ctx
{
  int device = 0;
  int sm_count = 148;
  static constexpr int kSlots = 3;                 // host-pointer pipeline depth
  cudaStream_t stream = nullptr;                   // all device-pointer work
  cudaStream_t slot_stream[kSlots] = {};           // host-pointer chunks round-robin here
  void* slot_buf[kSlots] = {};                     // device staging, one per slot
  size_t slot_bytes = 0;
  // plane stride (floats) of SoA arguments: `ld` is what the kernels launched right now use (0 = n), `user_ld` what
  // bbmcu_set_plane_stride asked for (device and host pointers of the caller)
  size_t ld = 0, user_ld = 0;
  void* pin_buf[kSlots] = {};                      // pinned staging for PAGEABLE host memory, one per slot
  size_t pin_bytes = 0;
  cudaEvent_t slot_done[kSlots] = {};              // recorded after the last D2H of the chunk in the slot
  std::string error;
  uint64_t launches = 0;
};

struct bbmcu_bsdf { bbmcu_host::Bsdf b; };

struct bbmcu_fit { std::vector<std::pair<std::string, bbmcu_host::Bsdf>> entries; };

namespace bbmcu {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

#define BBMCU_CUDA(call) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) throw ::bbmcu::CudaError(std::string(#call) + ": " + cudaGetErrorString(e__)); } while(0)

void set_thread_error(const std::string& msg);
const char* thread_error();

// run f(); translate exceptions into status codes + messages (ctx may be null)
template<class F> int guarded(bbmcu_ctx* ctx, F&& f)
{
  int rc = BBMCU_OK; std::string msg;
  try { f(); return BBMCU_OK; }
  catch(const CudaError& e) { rc = BBMCU_CUDA_ERROR; msg = e.what(); }
  catch(const std::invalid_argument& e) { rc = BBMCU_INVALID_ARGUMENT; msg = e.what(); }
  catch(const std::out_of_range& e) { rc = BBMCU_OUT_OF_RANGE; msg = e.what(); }
  catch(const std::exception& e) { rc = BBMCU_RUNTIME_ERROR; msg = e.what(); }
  catch(...) { rc = BBMCU_RUNTIME_ERROR; msg = "unknown error"; }
  if(ctx) ctx->error = msg;
  set_thread_error(msg);
  return rc;
}

bool is_device_pointer(const void* p);

// grid for an element-wise launch over `groups` thread-items: whole waves of the SMs, capped
inline unsigned grid_for(const bbmcu_ctx* ctx, size_t groups, int threads = 256, int blocks_per_sm = 8)
{
  size_t need = (groups + threads - 1) / threads;
  size_t cap = (size_t)ctx->sm_count * blocks_per_sm * 4;
  if(need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

} // namespace bbmcu
