// Piece 1 (K1): fused item-embedding gather + positional add + pad mask.
//
// Replaces  src_emb(ids) -> (x + pe[:, :L]) * mask[..., None]   (before dropout)
//   /root/reference/GURU/AutoEnc4Rec.py:71-75,:89-93,:176-181,:189-193
//   /root/reference/GURU/AutoEnc4Rec_cross.py:98-102,:124-128,:156-160
//   /root/reference/GURU/Transformer/transformer.py:104-106
//
// HBM-bound (no reuse except hot rows): per token 8 B id + 4d B row read + 4d B out write
// + 4 B mask write.  Layout: a CTA owns a contiguous tile of tokens; the tile's ids are
// loaded once, coalesced, into shared memory (narrowed to int32) together with the 0/1
// mask; then every thread issues kUnroll independent 128-bit row loads before the first
// use (memory-level parallelism), adds the PE vector (L2-resident, [L,d]) and streams the
// result out with 128-bit stores.  Arithmetic is an fp32 add followed by an fp32 multiply
// by 0/1 (no FMA contraction), i.e. bit-identical to the reference's two ATen kernels.
#include <cuda_bf16.h>
#include <limits.h>

#include "common.cuh"

namespace rg {

constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll = 8;

// DROP: the fused-dropout instantiation.  Compiled without the Philox path the plain kernel is 10 %
// faster (30.1 -> 26.8 us at cfg2) at the same four CTAs per SM; forcing five or six CTAs per SM
// (48 / 40 registers) made K1 no faster and the step slower (the plan branch beside it starves).
template <int VEC, bool DROP>  // VEC = float4 per row (d/4); 0 = runtime
__global__ void __launch_bounds__(kGatherThreads, 4)
gather_pe_mask_kernel(const float4* __restrict__ weight, const int64_t* __restrict__ ids,
                      const float4* __restrict__ pe, const float* __restrict__ mask_in,
                      int64_t pad_index, int64_t n_tokens, int L, int vec_rt,
                      float4* __restrict__ out, float* __restrict__ mask_out, DropoutParams drop) {
  rg::pdl_wait();
  extern __shared__ int smem_raw[];
  const int vec = VEC > 0 ? VEC : vec_rt;
  const int tok_per_block = (kGatherThreads * kGatherUnroll) / vec;
  int* s_id = smem_raw;
  float* s_mask = reinterpret_cast<float*>(smem_raw + tok_per_block);

  const int64_t tok0 = static_cast<int64_t>(blockIdx.x) * tok_per_block;
  // position of the tile's first token: ONE 64-bit modulo per thread; the per-element `tok % L` it
  // replaces (eight emulated 64-bit divisions per thread) was a third of the kernel's issue slots
  const uint32_t l0 = static_cast<uint32_t>(tok0 % L);
  for (int t = threadIdx.x; t < tok_per_block; t += kGatherThreads) {
    const int64_t tok = tok0 + t;
    int id = 0;
    float m = 0.f;
    if (tok < n_tokens) {
      const int64_t id64 = ids[tok];
      id = static_cast<int>(id64);
      m = mask_in ? mask_in[tok] : (id64 != pad_index ? 1.f : 0.f);
      if (mask_out) mask_out[tok] = m;
    }
    s_id[t] = id;
    s_mask[t] = m;
  }
  __syncthreads();

  float4 row[kGatherUnroll];
  int tl[kGatherUnroll], vv[kGatherUnroll];
#pragma unroll
  for (int u = 0; u < kGatherUnroll; ++u) {
    const int flat = u * kGatherThreads + threadIdx.x;
    tl[u] = flat / vec;
    vv[u] = flat - tl[u] * vec;
    if (tl[u] >= tok_per_block) tl[u] = INT_MAX / 2;  // d/4 does not divide the tile: idle slot
    if (tok0 + tl[u] < n_tokens) {
      row[u] = __ldg(weight + static_cast<int64_t>(s_id[tl[u]]) * vec + vv[u]);
    }
  }
  if (pe != nullptr) {
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const int64_t tok = tok0 + tl[u];
      if (tok < n_tokens) {
        const int l = static_cast<int>((l0 + static_cast<uint32_t>(tl[u])) % static_cast<uint32_t>(L));
        const float4 p = __ldg(pe + static_cast<int64_t>(l) * vec + vv[u]);
        const float m = s_mask[tl[u]];
        float4 r = row[u];
        r.x = __fmul_rn(__fadd_rn(r.x, p.x), m);
        r.y = __fmul_rn(__fadd_rn(r.y, p.y), m);
        r.z = __fmul_rn(__fadd_rn(r.z, p.z), m);
        r.w = __fmul_rn(__fadd_rn(r.w, p.w), m);
        // fused nn.Dropout (Transformer/transformer.py:104-106), bit-compatible with torch's kernel and
        // generator stream; rows under the mask are (signed) zeros whatever is drawn: no Philox there
        if (DROP && drop.keep_p > 0.f && m != 0.f) r = torch_dropout4(r, tok * vec + vv[u], drop);
        st_stream_f4(out + tok * vec + vv[u], r);
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
      const int64_t tok = tok0 + tl[u];
      if (tok < n_tokens) st_stream_f4(out + tok * vec + vv[u], row[u]);
    }
  }
  rg::pdl_trigger();
}

// PositionalEncoding.forward(x, mask) on an already materialised x.
__global__ void __launch_bounds__(256)
pe_mask_kernel(const float4* __restrict__ x, const float4* __restrict__ pe,
               const float* __restrict__ mask, int64_t n_vec, int vec, int L,
               float4* __restrict__ out) {
  rg::pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
       i += stride) {
    const int64_t tok = i / vec;
    const int v = static_cast<int>(i - tok * vec);
    const int l = static_cast<int>(tok % L);
    float4 r = ld_stream_f4(x + i);
    const float4 p = __ldg(pe + static_cast<int64_t>(l) * vec + v);
    const float m = __ldg(mask + tok);
    r.x = __fmul_rn(__fadd_rn(r.x, p.x), m);
    r.y = __fmul_rn(__fadd_rn(r.y, p.y), m);
    r.z = __fmul_rn(__fadd_rn(r.z, p.z), m);
    r.w = __fmul_rn(__fadd_rn(r.w, p.w), m);
    st_stream_f4(out + i, r);
  }
  rg::pdl_trigger();
}

}  // namespace rg

namespace rg {
// mask[i] = ids[i] != pad_index (M1 on its own: lets the table-backward plan start before K1)
__global__ void __launch_bounds__(256)
pad_mask_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t pad_index, float* __restrict__ mask) {
  rg::pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    mask[i] = ids[i] != pad_index ? 1.f : 0.f;
  rg::pdl_trigger();
}
// The same mask plus its population count: count3 = {accumulator, tickets, result}, zeroed once by
// the caller.  Integer atomics (exact, order independent); the last CTA to finish publishes the
// count in count3[2] and clears the other two words for the next call.
__global__ void __launch_bounds__(256)
pad_mask_count_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t pad_index, float* __restrict__ mask,
                      uint32_t* __restrict__ count3) {
  rg::pdl_wait();
  __shared__ uint32_t s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  uint32_t mine = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool live = ids[i] != pad_index;
    mask[i] = live ? 1.f : 0.f;
    mine += live ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&count3[0], s_cnt);
    __threadfence();
    const uint32_t ticket = atomicAdd(&count3[1], 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      count3[2] = atomicExch(&count3[0], 0u);
      count3[1] = 0u;
    }
  }
  rg::pdl_trigger();
}
}  // namespace rg

extern "C" int recguru_pad_mask_count(const int64_t* ids, int64_t n, int64_t pad_index, float* mask_out,
                                      uint32_t* count3, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n > 0 && ids && mask_out && count3, RECGURU_E_SHAPE, "pad_mask_count: null pointer or n <= 0");
  const int64_t blocks = rg::min64(ceil_div(n, 256 * 4), static_cast<int64_t>(num_sms()) * 8);
  rg::launch_pdl((pad_mask_count_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream),
                 ids, n, pad_index, mask_out, count3);
  RG_CHECK_LAUNCH("pad_mask_count");
  return RECGURU_OK;
}

extern "C" int recguru_pad_mask(const int64_t* ids, int64_t n, int64_t pad_index, float* mask_out,
                                recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n >= 0 && (n == 0 || (ids && mask_out)), RECGURU_E_SHAPE, "pad_mask: null pointer");
  if (n == 0) return RECGURU_OK;
  const int64_t blocks = rg::min64(ceil_div(n, 256 * 4), static_cast<int64_t>(num_sms()) * 8);
  rg::launch_pdl((pad_mask_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), ids, n,
                 pad_index, mask_out);
  RG_CHECK_LAUNCH("pad_mask");
  return RECGURU_OK;
}

static int gather_pe_mask_impl(const float* weight, int64_t V1, int64_t d,
                               const int64_t* ids, int64_t n_tokens, int64_t L,
                               const float* pe, const float* mask_in, int64_t pad_index,
                               float* out, float* mask_out, rg::DropoutParams drop, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(weight && (n_tokens == 0 || (ids && out)), RECGURU_E_SHAPE,
             "gather_pe_mask: null pointer");
  RG_REQUIRE(V1 > 0 && V1 < (1ll << 31) && n_tokens >= 0 && L > 0, RECGURU_E_SHAPE,
             "gather_pe_mask: bad sizes V1=%lld n_tokens=%lld L=%lld", (long long)V1,
             (long long)n_tokens, (long long)L);
  RG_REQUIRE(d > 0 && d % 4 == 0 && d <= 4096, RECGURU_E_UNSUPPORTED,
             "gather_pe_mask: d=%lld must be a multiple of 4 and <= 4096", (long long)d);
  RG_REQUIRE(aligned16(weight) && aligned16(out) && aligned16(pe), RECGURU_E_SHAPE,
             "gather_pe_mask: weight/pe/out must be 16-byte aligned");
  if (n_tokens == 0) return RECGURU_OK;
  RG_CHECK_IDS("gather_pe_mask", ids, n_tokens, 0, V1, static_cast<cudaStream_t>(stream));
  const int vec = static_cast<int>(d / 4);
  const int tok_per_block = (kGatherThreads * kGatherUnroll) / vec;
  const int64_t blocks = ceil_div(n_tokens, tok_per_block);
  const size_t smem = static_cast<size_t>(tok_per_block) * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float4* w4 = reinterpret_cast<const float4*>(weight);
  const float4* pe4 = reinterpret_cast<const float4*>(pe);
  float4* o4 = reinterpret_cast<float4*>(out);
#define RG_LAUNCH_GATHER(V)                                                              \
  do {                                                                                   \
    if (drop.keep_p > 0.f)                                                               \
      rg::launch_pdl((gather_pe_mask_kernel<V, true>), static_cast<unsigned>(blocks), kGatherThreads, smem, st,  \
          w4, ids, pe4, mask_in, pad_index, n_tokens, static_cast<int>(L), vec, o4, mask_out, drop);  \
    else                                                                                 \
      rg::launch_pdl((gather_pe_mask_kernel<V, false>), static_cast<unsigned>(blocks), kGatherThreads, smem, st, \
          w4, ids, pe4, mask_in, pad_index, n_tokens, static_cast<int>(L), vec, o4, mask_out, drop);  \
  } while (0)
  switch (vec) {
    case 8: RG_LAUNCH_GATHER(8); break;
    case 16: RG_LAUNCH_GATHER(16); break;
    case 32: RG_LAUNCH_GATHER(32); break;
    case 64: RG_LAUNCH_GATHER(64); break;
    default: RG_LAUNCH_GATHER(0); break;
  }
#undef RG_LAUNCH_GATHER
  RG_CHECK_LAUNCH("gather_pe_mask");
  return RECGURU_OK;
}

extern "C" int recguru_gather_pe_mask(const float* weight, int64_t V1, int64_t d,
                                      const int64_t* ids, int64_t n_tokens, int64_t L,
                                      const float* pe, const float* mask_in, int64_t pad_index,
                                      float* out, float* mask_out, recguru_stream_t stream) {
  rg::DropoutParams none = {};
  return gather_pe_mask_impl(weight, V1, d, ids, n_tokens, L, pe, mask_in, pad_index, out, mask_out, none, stream);
}

namespace rg {
static int make_dropout(const char* who, double p_drop, uint64_t seed, uint64_t offset, int64_t torch_threads,
                        DropoutParams* dp) {
  RG_REQUIRE(p_drop >= 0.0 && p_drop < 1.0, RECGURU_E_SHAPE, "%s: dropout p=%g not in [0, 1)", who, p_drop);
  RG_REQUIRE(offset % 4 == 0 && torch_threads > 0 && torch_threads < (1ll << 32), RECGURU_E_SHAPE,
             "%s: philox offset must be a multiple of 4 and 0 < torch_threads < 2^32", who);
  const float keep = static_cast<float>(1.0 - p_drop);       // torch: pa = (float)(1. - p)
  dp->keep_p = p_drop == 0.0 ? 0.f : keep;
  dp->scale = 1.0f / keep;
  dp->seed_lo = static_cast<uint32_t>(seed);
  dp->seed_hi = static_cast<uint32_t>(seed >> 32);
  dp->offset4 = offset / 4;
  dp->T = static_cast<uint32_t>(torch_threads);
  return RECGURU_OK;
}

__global__ void __launch_bounds__(256)
dropout_kernel(const float4* __restrict__ x, const float* __restrict__ row_mask, int vec, int64_t n_vec,
               float4* __restrict__ out, DropoutParams drop) {
  rg::pdl_wait();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    float4 v = ld_stream_f4(x + i);
    // rows whose mask is 0 hold zeros on the forward side and receive no gradient use on the backward
    // side: they pass through unchanged (sign of zero included)
    if (row_mask == nullptr || __ldg(row_mask + i / vec) != 0.f) v = torch_dropout4(v, i, drop);
    st_stream_f4(out + i, v);
  }
  rg::pdl_trigger();
}
}  // namespace rg

extern "C" int recguru_gather_pe_mask_dropout(const float* weight, int64_t V1, int64_t d,
                                              const int64_t* ids, int64_t n_tokens, int64_t L,
                                              const float* pe, const float* mask_in, int64_t pad_index,
                                              double p_drop, uint64_t seed, uint64_t philox_offset,
                                              int64_t torch_threads, float* out, float* mask_out,
                                              recguru_stream_t stream) {
  rg::DropoutParams dp = {};
  int rc = rg::make_dropout("gather_pe_mask_dropout", p_drop, seed, philox_offset, torch_threads, &dp);
  if (rc) return rc;
  RG_REQUIRE(pe != nullptr, RECGURU_E_SHAPE, "gather_pe_mask_dropout: pe is null (dropout follows the positional add)");
  return gather_pe_mask_impl(weight, V1, d, ids, n_tokens, L, pe, mask_in, pad_index, out, mask_out, dp, stream);
}

extern "C" int recguru_dropout(const float* x, const float* row_mask, int64_t n_rows, int64_t d, double p_drop,
                               uint64_t seed, uint64_t philox_offset, int64_t torch_threads, float* out,
                               recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n_rows >= 0 && d > 0 && d % 4 == 0, RECGURU_E_SHAPE, "dropout: bad sizes (d must be a multiple of 4)");
  if (n_rows == 0) return RECGURU_OK;
  RG_REQUIRE(x && out && aligned16(x) && aligned16(out), RECGURU_E_SHAPE, "dropout: null or unaligned pointer");
  DropoutParams dp = {};
  int rc = make_dropout("dropout", p_drop, seed, philox_offset, torch_threads, &dp);
  if (rc) return rc;
  RG_REQUIRE(dp.keep_p > 0.f, RECGURU_E_SHAPE, "dropout: p must be > 0");
  const int64_t n_vec = n_rows * (d / 4);
  const int64_t blocks = rg::min64(ceil_div(n_vec, 256 * 4), static_cast<int64_t>(num_sms()) * 16);
  rg::launch_pdl((dropout_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const float4*>(x), row_mask, static_cast<int>(d / 4), n_vec,
                 reinterpret_cast<float4*>(out), dp);
  RG_CHECK_LAUNCH("dropout");
  return RECGURU_OK;
}

extern "C" int recguru_pe_mask(const float* x, const float* pe, const float* mask,
                               int64_t n_tokens, int64_t L, int64_t d, float* out,
                               recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(x && pe && mask && out, RECGURU_E_SHAPE, "pe_mask: null pointer");
  RG_REQUIRE(d > 0 && d % 4 == 0, RECGURU_E_UNSUPPORTED, "pe_mask: d=%lld must be a multiple of 4",
             (long long)d);
  RG_REQUIRE(aligned16(x) && aligned16(pe) && aligned16(out), RECGURU_E_SHAPE,
             "pe_mask: x/pe/out must be 16-byte aligned");
  RG_REQUIRE(n_tokens >= 0 && L > 0, RECGURU_E_SHAPE, "pe_mask: bad sizes");
  if (n_tokens == 0) return RECGURU_OK;
  const int vec = static_cast<int>(d / 4);
  const int64_t n_vec = n_tokens * vec;
  int64_t blocks = ceil_div(n_vec, 256 * 4);
  if (blocks > static_cast<int64_t>(num_sms()) * 16) blocks = static_cast<int64_t>(num_sms()) * 16;
  rg::launch_pdl((pe_mask_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(pe), mask, n_vec, vec,
      static_cast<int>(L), reinterpret_cast<float4*>(out));
  RG_CHECK_LAUNCH("pe_mask");
  return RECGURU_OK;
}

// ---- P1' (learned positions, Transformer/transformer.py:327-342) and user-rep extraction --------
namespace rg {
// dP[l, :] = sum_b mask[b, l] * g[b, l, :]  -- the gradient of a LEARNED position table behind
// x + pos_emb(arange(L)).  One CTA per position l (x one 128-float column block): warp w sums the
// batch rows b = w, w + 8, ... in ascending order, the eight partial rows are combined in warp order:
// a fixed summation tree, bit-reproducible.
__global__ void __launch_bounds__(256)
pe_grad_kernel(const float4* __restrict__ g, const float* __restrict__ mask, int64_t B, int L, int vec,
               int accumulate, float4* __restrict__ dP) {
  rg::pdl_wait();
  __shared__ float4 s_part[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = blockIdx.x;
  const int col = blockIdx.y * 32 + lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < vec) {
    for (int64_t b = warp; b < B; b += 8) {
      const float m = __ldg(mask + b * L + l);
      if (m != 0.f) {
        const float4 v = ld_stream_f4(g + (b * L + l) * vec + col);
        acc.x = fmaf(m, v.x, acc.x); acc.y = fmaf(m, v.y, acc.y);
        acc.z = fmaf(m, v.z, acc.z); acc.w = fmaf(m, v.w, acc.w);
      }
    }
  }
  s_part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && col < vec) {
    float4 t = s_part[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      t.x += s_part[w][lane].x; t.y += s_part[w][lane].y; t.z += s_part[w][lane].z; t.w += s_part[w][lane].w;
    }
    float4* dst = dP + static_cast<int64_t>(l) * vec + col;
    if (accumulate) { const float4 o = *dst; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
    *dst = t;
  }
  rg::pdl_trigger();
}

// u[b, :] = h[b, L-1, :]  (dec_embedding[:, -1, :] / get_seq_embed(...)[:, -1, :]), fp32 and/or bf16
__global__ void __launch_bounds__(256)
extract_last_kernel(const float4* __restrict__ h, int64_t B, int L, int vec, float4* __restrict__ out_f32,
                    uint2* __restrict__ out_bf16) {
  rg::pdl_wait();
  const int64_t n = B * vec;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t b = i / vec;
    const int c = static_cast<int>(i - b * vec);
    const float4 v = __ldg(h + (b * L + (L - 1)) * vec + c);
    if (out_f32) out_f32[i] = v;
    if (out_bf16) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
      const __nv_bfloat162 bb = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&a);
      o.y = *reinterpret_cast<const uint32_t*>(&bb);
      out_bf16[i] = o;
    }
  }
  rg::pdl_trigger();
}
}  // namespace rg

extern "C" int recguru_pe_grad(const float* g, const float* mask, int64_t B, int64_t L, int64_t d, int accumulate,
                               float* dP, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(B >= 0 && L > 0 && L < (1 << 30) && d > 0 && d % 4 == 0, RECGURU_E_SHAPE, "pe_grad: bad sizes");
  RG_REQUIRE(dP && aligned16(dP) && (B == 0 || (g && mask && aligned16(g))), RECGURU_E_SHAPE,
             "pe_grad: null or unaligned pointer");
  const int vec = static_cast<int>(d / 4);
  const dim3 grid(static_cast<unsigned>(L), static_cast<unsigned>(ceil_div(vec, 32)));
  rg::launch_pdl((pe_grad_kernel), grid, dim3(256), 0, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const float4*>(g), mask, B, static_cast<int>(L), vec, accumulate,
                 reinterpret_cast<float4*>(dP));
  RG_CHECK_LAUNCH("pe_grad");
  return RECGURU_OK;
}

extern "C" int recguru_extract_last(const float* h, int64_t B, int64_t L, int64_t d, float* out_f32,
                                    void* out_bf16, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(B >= 0 && L > 0 && d > 0 && d % 4 == 0, RECGURU_E_SHAPE, "extract_last: bad sizes");
  if (B == 0) return RECGURU_OK;
  RG_REQUIRE(h && aligned16(h) && (out_f32 || out_bf16), RECGURU_E_SHAPE, "extract_last: null pointer");
  RG_REQUIRE((!out_f32 || aligned16(out_f32)) && (reinterpret_cast<uintptr_t>(out_bf16) & 7u) == 0, RECGURU_E_SHAPE,
             "extract_last: unaligned output");
  const int vec = static_cast<int>(d / 4);
  const int64_t blocks = rg::min64(ceil_div(B * vec, 256), static_cast<int64_t>(num_sms()) * 8);
  rg::launch_pdl((extract_last_kernel), static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream),
                 reinterpret_cast<const float4*>(h), B, static_cast<int>(L), vec, reinterpret_cast<float4*>(out_f32),
                 static_cast<uint2*>(out_bf16));
  RG_CHECK_LAUNCH("extract_last");
  return RECGURU_OK;
}
