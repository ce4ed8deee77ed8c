// Error plumbing, bf16 cast, per-shard top-K merge (K5) and the rank histogram (H1).
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace rg {

static thread_local char t_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

int num_sms() {
  static std::atomic<int> cache[kMaxDevices];   // 0 = not queried yet
  const int dev = current_device();
  if (dev >= 0 && dev < kMaxDevices) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n = 0;
  if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = 148;   // B200
  if (dev >= 0 && dev < kMaxDevices) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

// ---- RECGURU_CHECK_IDS=1: debug range check of item ids ---------------------------------
// g_bad_id: {position of the first offender, its value}; one instance per device (module scope).
__device__ unsigned long long g_bad_id[2];

__global__ void check_ids_reset_kernel() { g_bad_id[0] = ~0ull; g_bad_id[1] = 0ull; }

__global__ void __launch_bounds__(256)
check_ids_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t lo, int64_t hi) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t v = ids[i];
    if (v < lo || v >= hi) atomicMin(&g_bad_id[0], static_cast<unsigned long long>(i));
  }
}

__global__ void check_ids_value_kernel(const int64_t* __restrict__ ids) {
  if (g_bad_id[0] != ~0ull) g_bad_id[1] = static_cast<unsigned long long>(ids[g_bad_id[0]]);
}

bool check_ids_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RECGURU_CHECK_IDS");
    v = (e && e[0] != '\0' && e[0] != '0') ? 1 : 0;
  }
  return v != 0;
}

int check_ids(const char* who, const int64_t* ids, int64_t n, int64_t lo, int64_t hi, cudaStream_t st) {
  if (n <= 0 || ids == nullptr) return RECGURU_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)
    return RECGURU_OK;     // a capturing stream cannot be synchronised: the check is skipped
  check_ids_reset_kernel<<<1, 1, 0, st>>>();
  const int64_t blocks = rg::min64(ceil_div(n, 256 * 4), static_cast<int64_t>(num_sms()) * 8);
  check_ids_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(ids, n, lo, hi);
  check_ids_value_kernel<<<1, 1, 0, st>>>(ids);
  unsigned long long bad[2] = {~0ull, 0ull};
  cudaError_t e = cudaMemcpyFromSymbolAsync(bad, g_bad_id, sizeof(bad), 0, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  RG_REQUIRE(e == cudaSuccess, RECGURU_E_CUDA, "%s: id check failed to run: %s", who, cudaGetErrorString(e));
  count_launch(3);
  RG_REQUIRE(bad[0] == ~0ull, RECGURU_E_INDEX, "%s: id %lld at position %llu is outside [%lld, %lld)", who,
             static_cast<long long>(bad[1]), bad[0], static_cast<long long>(lo), static_cast<long long>(hi));
  return RECGURU_OK;
}

// fp32 -> bf16, round-to-nearest-even (same as torch .to(torch.bfloat16))
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float4* __restrict__ in, int64_t n_vec, const float* __restrict__ in_tail,
                 int64_t n, uint2* __restrict__ out, __nv_bfloat16* __restrict__ out_tail) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n_vec; i += stride) {
    const float4 v = ld_stream_f4(in + i);
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    const __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b);
    out[i] = o;
  }
  for (int64_t i = n_vec * 4 + t0; i < n; i += stride) out_tail[i] = __float2bfloat16_rn(in_tail[i]);
}

// Merge G sorted (desc score, asc id) lists of length k per user; one thread per user.
// G*k <= 8*32; a G-way head-pointer merge keeps it O(k*G) compares with no scratch.
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ vals, const int64_t* __restrict__ ids,
                  const int64_t* __restrict__ counts, int G, int64_t U, int k,
                  float* __restrict__ out_val, int64_t* __restrict__ out_id,
                  int64_t* __restrict__ out_rank) {
  const int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (u >= U) return;
  int head[64];
  for (int g = 0; g < G; ++g) head[g] = 0;
  for (int j = 0; j < k; ++j) {
    int best = -1;
    float bv = -CUDART_INF_F;
    int64_t bi = -1;
    for (int g = 0; g < G; ++g) {
      if (head[g] >= k) continue;
      const int64_t off = (static_cast<int64_t>(g) * U + u) * k + head[g];
      const int64_t id = ids[off];
      if (id < 0) { head[g] = k; continue; }  // padding: list g is exhausted
      const float v = vals[off];
      if (best < 0 || v > bv || (v == bv && id < bi)) { best = g; bv = v; bi = id; }
    }
    if (best >= 0) ++head[best];
    out_val[u * k + j] = best >= 0 ? bv : -CUDART_INF_F;
    out_id[u * k + j] = best >= 0 ? bi : -1;
  }
  if (out_rank) {
    int64_t r = 0;
    for (int g = 0; g < G; ++g) r += counts[static_cast<int64_t>(g) * U + u];
    out_rank[u] = r;
  }
}

// hist[min(rank, kmax)] += 1  -- integer atomics only (exact, order independent)
__global__ void __launch_bounds__(256)
rank_hist_kernel(const int64_t* __restrict__ rank, int64_t U, int kmax,
                 unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int s_hist[];
  for (int b = threadIdx.x; b <= kmax; b += blockDim.x) s_hist[b] = 0;
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < U; i += stride) {
    const int64_t r = rank[i];
    atomicAdd(&s_hist[r < kmax ? (r < 0 ? 0 : static_cast<int>(r)) : kmax], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b <= kmax; b += blockDim.x)
    if (s_hist[b]) atomicAdd(&hist[b], static_cast<unsigned long long>(s_hist[b]));
}

// int32 -> int64 ids: 4 per thread, 128-bit load, two 128-bit stores
__global__ void __launch_bounds__(256)
widen_ids_kernel(const int4* __restrict__ in4, const int32_t* __restrict__ in, int64_t n_vec, int64_t n,
                 longlong2* __restrict__ out2, int64_t* __restrict__ out) {
  rg::pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n_vec; i += stride) {
    const int4 v = __ldg(in4 + i);
    out2[2 * i] = make_longlong2(v.x, v.y);
    out2[2 * i + 1] = make_longlong2(v.z, v.w);
  }
  for (int64_t i = n_vec * 4 + t0; i < n; i += stride) out[i] = in[i];
  rg::pdl_trigger();
}

__global__ void timestamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *slot = t;
}

__global__ void zero_u64_kernel(unsigned long long* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0ull;
}

}  // namespace rg

extern "C" int recguru_abi_version(void) { return 6; }
extern "C" const char* recguru_last_error(void) { return rg::t_error; }
extern "C" int64_t recguru_launch_count(void) { return rg::g_launches.load(); }

extern "C" int recguru_cast_bf16(const float* in, int64_t n, void* out_bf16,
                                 recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n >= 0, RECGURU_E_SHAPE, "cast_bf16: n < 0");
  if (n == 0) return RECGURU_OK;
  RG_REQUIRE(in && out_bf16 && aligned16(in) && (reinterpret_cast<uintptr_t>(out_bf16) & 7u) == 0,
             RECGURU_E_SHAPE, "cast_bf16: null or unaligned pointer");
  const int64_t n_vec = n / 4;
  const int64_t blocks = rg::min64(ceil_div(n_vec > 0 ? n_vec : 1, 256 * 4), static_cast<int64_t>(num_sms()) * 16);
  cast_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(in), n_vec, in, n, static_cast<uint2*>(out_bf16),
      static_cast<__nv_bfloat16*>(out_bf16));
  RG_CHECK_LAUNCH("cast_bf16");
  return RECGURU_OK;
}

extern "C" int recguru_topk_merge(const float* vals, const int64_t* ids, const int64_t* counts,
                                  int64_t G, int64_t U, int64_t k, float* out_val, int64_t* out_id,
                                  int64_t* out_rank, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(k == 0 || (vals && ids && out_val && out_id), RECGURU_E_SHAPE, "topk_merge: null pointer");
  RG_REQUIRE(out_rank == nullptr || counts != nullptr, RECGURU_E_SHAPE,
             "topk_merge: out_rank requires counts");
  RG_REQUIRE(G >= 1 && G <= 64 && U >= 0 && k >= 0 && k <= 32, RECGURU_E_UNSUPPORTED,
             "topk_merge: G=%lld (1..64) k=%lld (0..32)", (long long)G, (long long)k);
  if (U == 0) return RECGURU_OK;
  topk_merge_kernel<<<static_cast<unsigned>(ceil_div(U, 128)), 128, 0,
                      static_cast<cudaStream_t>(stream)>>>(vals, ids, counts, static_cast<int>(G),
                                                           U, static_cast<int>(k), out_val, out_id,
                                                           out_rank);
  RG_CHECK_LAUNCH("topk_merge");
  return RECGURU_OK;
}

extern "C" int recguru_rank_histogram(const int64_t* rank, int64_t U, int64_t kmax, int64_t* hist,
                                      int accumulate, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(hist && (rank || U == 0), RECGURU_E_SHAPE, "rank_histogram: null pointer");
  RG_REQUIRE(kmax >= 1 && kmax <= 4096 && U >= 0, RECGURU_E_UNSUPPORTED,
             "rank_histogram: kmax=%lld not in [1,4096]", (long long)kmax);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
  if (!accumulate) {
    zero_u64_kernel<<<static_cast<unsigned>(ceil_div(kmax + 1, 256)), 256, 0, st>>>(
        h, static_cast<int>(kmax + 1));
    RG_CHECK_LAUNCH("zero_u64");
  }
  if (U == 0) return RECGURU_OK;
  const int64_t blocks = rg::min64(ceil_div(U, 256 * 8), static_cast<int64_t>(num_sms()) * 4);
  rank_hist_kernel<<<static_cast<unsigned>(blocks), 256, (kmax + 1) * sizeof(unsigned int), st>>>(
      rank, U, static_cast<int>(kmax), h);
  RG_CHECK_LAUNCH("rank_hist");
  return RECGURU_OK;
}

extern "C" int recguru_widen_ids(const int32_t* ids32_in, int64_t n, int64_t* ids_out, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n >= 0, RECGURU_E_SHAPE, "widen_ids: n < 0");
  if (n == 0) return RECGURU_OK;
  RG_REQUIRE(ids32_in && ids_out && aligned16(ids32_in) && aligned16(ids_out), RECGURU_E_SHAPE,
             "widen_ids: null or unaligned pointer");
  const int64_t n_vec = n / 4;
  const int64_t blocks = rg::min64(ceil_div(n_vec > 0 ? n_vec : 1, 256 * 2), static_cast<int64_t>(num_sms()) * 8);
  rg::launch_pdl((widen_ids_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const int4*>(ids32_in), ids32_in, n_vec, n,
                 reinterpret_cast<longlong2*>(ids_out), ids_out);
  RG_CHECK_LAUNCH("widen_ids");
  return RECGURU_OK;
}

extern "C" int recguru_debug_timestamp(uint64_t* slot, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(slot != nullptr, RECGURU_E_SHAPE, "debug_timestamp: null pointer");
  timestamp_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(slot));
  RG_CHECK_LAUNCH("debug_timestamp");
  return RECGURU_OK;
}
