// Shared helpers for the recguru_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <utility>

#include "../../include/recguru_b200.h"

namespace rg {

// ---- error plumbing (thread-local message, global launch counter) -------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define RG_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      rg::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

// Checks the launch that was just enqueued (no synchronisation).
#define RG_CHECK_LAUNCH(name)                                                     \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      rg::set_error("%s: CUDA launch error: %s", name, cudaGetErrorString(e__));  \
      return RECGURU_E_CUDA;                                                      \
    }                                                                             \
    rg::count_launch();                                                           \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t min64(int64_t a, int64_t b) { return a < b ? a : b; }

// ---- per-device state ------------------------------------------------------------------
// The library may be driven from one process on several devices and from several host threads
// (the reference's only parallel mode is nn.DataParallel: one Python thread per GPU,
// train_auto.py:350-351).  Everything cached on the host side is therefore keyed by the CURRENT
// device; nothing is process-wide.
constexpr int kMaxDevices = 64;
int current_device();          // cudaGetDevice, -1 on error
int num_sms();                 // multiprocessor count of the current device (148 on B200), cached
// "done once per device" flag for per-device function attributes (cudaFuncSetAttribute is per
// device, a process-wide static would leave every device but the first without it).
struct PerDeviceOnce {
  unsigned long long mask = 0;   // bit d set <=> done on device d (accessed with atomics)
  bool need(int dev) const {
    if (dev < 0 || dev >= kMaxDevices) return true;
    return ((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> dev) & 1ull) == 0;
  }
  void done(int dev) {
    if (dev >= 0 && dev < kMaxDevices) __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELEASE);
  }
};
// RECGURU_CHECK_IDS=1 (read once): every entry point that takes item ids verifies
// 0 <= id < V1 on the device before launching and SYNCHRONISES the stream to report the first
// offender (RECGURU_E_INDEX).  Debug aid: off by default, skipped while the stream is capturing.
bool check_ids_enabled();
int check_ids(const char* who, const int64_t* ids, int64_t n, int64_t lo, int64_t hi, cudaStream_t st);
#define RG_CHECK_IDS(who, ids, n, lo, hi, st)                                  \
  do {                                                                         \
    if (rg::check_ids_enabled()) {                                             \
      const int rc__ = rg::check_ids(who, ids, n, lo, hi, st);                 \
      if (rc__) return rc__;                                                   \
    }                                                                          \
  } while (0)

// ---- programmatic dependent launch ---------------------------------------------------
// The training-step path is a chain of short kernels on one stream.  Every kernel of the chain
// starts with griddepcontrol.wait (nothing is read or written before the previous grid has
// completed and flushed) and ends with griddepcontrol.launch_dependents, and is launched with
// the programmatic-stream-serialization attribute: the next kernel's launch and CTA scheduling
// overlap this kernel's tail instead of waiting for the full drain.  RECGURU_PDL=0 turns the
// attribute off (the device instructions are then no-ops).
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RECGURU_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);   // error picked up by RG_CHECK_LAUNCH
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
#endif

// ---- device helpers --------------------------------------------------------------
#ifdef __CUDACC__
// streaming 128-bit load that does not pollute L1 (read-once data)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// streaming 128-bit store (written once, read by a later kernel)
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace rg
