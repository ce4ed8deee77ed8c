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
// Philox4x32-10 (Salmon et al.), the generator of cuRAND's curandStatePhilox4_32_10 and of torch's
// CUDA generator: counter (c0..c3), key (k0, k1).
struct Philox4 { uint32_t x, y, z, w; };
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return {c0, c1, c2, c3};
}

// torch's fused CUDA dropout (ATen/native/cuda/Dropout.cu, fused_dropout_kernel_vec<float, ..., 4>)
// restated per float4: torch launches T = grid * 256 threads, thread t takes the float4 groups
// q = t, t + T, t + 2T, ...; for its i-th group it draws curand_uniform4 number i of the Philox
// stream (seed, subsequence t, offset), keeps element c iff uniform_c < keep_p, and writes
// (x * keep) * (1 / keep_p).  cuRAND: subsequence -> counter words 2-3, offset / 4 + i -> words 0-1,
// uniform = u32 * 2^-32 + 2^-33.  So group q needs one Philox block, whoever computes it.
struct DropoutParams {
  float keep_p;          // 1 - p; 0 disables
  float scale;           // 1 / keep_p (fp32, as torch computes it)
  uint32_t seed_lo, seed_hi;
  uint64_t offset4;      // philox offset / 4
  uint32_t T;            // threads of torch's launch
};
__device__ __forceinline__ float4 torch_dropout4(float4 v, int64_t q, const DropoutParams& dp) {
  const uint32_t t = static_cast<uint32_t>(q % dp.T);
  const uint64_t c = dp.offset4 + static_cast<uint64_t>(q / dp.T);
  const Philox4 r = philox4x32_10(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), t, 0u, dp.seed_lo, dp.seed_hi);
  constexpr float kInv = 2.3283064365386963e-10f;             // 2^-32
  const float u0 = __fmaf_rn(static_cast<float>(r.x), kInv, kInv * 0.5f);
  const float u1 = __fmaf_rn(static_cast<float>(r.y), kInv, kInv * 0.5f);
  const float u2 = __fmaf_rn(static_cast<float>(r.z), kInv, kInv * 0.5f);
  const float u3 = __fmaf_rn(static_cast<float>(r.w), kInv, kInv * 0.5f);
  v.x = __fmul_rn(__fmul_rn(v.x, u0 < dp.keep_p ? 1.f : 0.f), dp.scale);
  v.y = __fmul_rn(__fmul_rn(v.y, u1 < dp.keep_p ? 1.f : 0.f), dp.scale);
  v.z = __fmul_rn(__fmul_rn(v.z, u2 < dp.keep_p ? 1.f : 0.f), dp.scale);
  v.w = __fmul_rn(__fmul_rn(v.w, u3 < dp.keep_p ? 1.f : 0.f), dp.scale);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace rg
