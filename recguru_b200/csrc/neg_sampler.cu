// the negative sampler of the training loader, on the GPU.
//
// Reference (CPU, per sample, inside pickle_loader.__getitem__,
// /root/reference/GURU/data/data_loader.py:289-314): a weight vector over the vocabulary --
// np.ones(v_size), or item_freq^0.75 (:251-254) -- with entry 0, the user's whole sequence, the
// validation and the test item zeroed, then torch.multinomial(weights, enc_maxlen*num_n,
// replacement=True).  That is O(v_size) work and memory per sequence on the host and
// 8 B * L * K of H2D traffic per sequence.
//
// Here: one CTA per sequence draws its L*K negatives by rejection from the same distribution --
// uniform over [1, v_size) or Walker's alias table of the base weights -- rejecting id 0 and
// anything in the user's exclusion list (history in shared memory + val + test).  Rejection from
// the base distribution conditioned on "not excluded" IS the reference's distribution (zeroed
// weights renormalised).  Randomness: counter-based Philox4x32-10 keyed by (seed), counter
// (slot, attempt, sequence, step): the same (seed, step) always yields the same negatives,
// independent of launch geometry.
#include "common.cuh"

namespace rg {

constexpr int kSamplerThreads = 256;
constexpr int kMaxAttempts = 4096;   // P(all rejected) is astronomically small unless nothing is allowed

__global__ void __launch_bounds__(kSamplerThreads)
sample_negatives_kernel(const int64_t* __restrict__ history, int H, int64_t pad_index,
                        const int64_t* __restrict__ val, const int64_t* __restrict__ test,
                        uint32_t v_size, const float* __restrict__ alias_prob,
                        const int32_t* __restrict__ alias_idx, int64_t num_n, uint32_t seed_lo,
                        uint32_t seed_hi, uint32_t step, const uint64_t* __restrict__ step_dev,
                        const int64_t* __restrict__ offsets, const int64_t* __restrict__ user_idx,
                        int64_t* __restrict__ out) {
  rg::pdl_wait();
  if (step_dev != nullptr) step = static_cast<uint32_t>(*step_dev);   // device-side step counter (CUDA graphs)
  extern __shared__ int s_hist[];
  __shared__ int s_n;
  const int64_t b = blockIdx.x;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  // Where this sequence's history lives: row b of a padded [B, H] matrix, or (offsets != NULL) the
  // ragged store -- history = all sequences back to back, user u = user_idx[b] owns
  // [offsets[u], offsets[u+1]), and val / test are indexed by user.
  const int64_t u = offsets ? (user_idx ? user_idx[b] : b) : b;
  const int64_t* hist = offsets ? history + offsets[u] : history + b * H;
  const int n_h = offsets ? static_cast<int>(rg::min64(offsets[u + 1] - offsets[u], H)) : H;
  // exclusion list: the non-pad entries of this sequence's history (order irrelevant)
  for (int i = threadIdx.x; i < n_h; i += kSamplerThreads) {
    const int64_t id = hist[i];
    if (id != pad_index && id > 0 && id < v_size) s_hist[atomicAdd(&s_n, 1)] = static_cast<int>(id);
  }
  __syncthreads();
  const int n_hist = s_n;
  const int v0 = val ? static_cast<int>(val[u]) : 0;
  const int t0 = test ? static_cast<int>(test[u]) : 0;
  const uint32_t n_choices = v_size - 1;                        // ids 1 .. v_size-1
  for (int64_t slot = threadIdx.x; slot < num_n; slot += kSamplerThreads) {
    int id = 1;
    for (int attempt = 0; attempt < kMaxAttempts; ++attempt) {
      const Philox4 r = philox4x32_10(static_cast<uint32_t>(slot), static_cast<uint32_t>(attempt),
                                      static_cast<uint32_t>(b), step, seed_lo, seed_hi);
      // unbiased integer in [0, n_choices): Lemire's multiply-shift with rejection of the short range
      uint64_t m = static_cast<uint64_t>(r.x) * n_choices;
      if (static_cast<uint32_t>(m) < n_choices) {
        const uint32_t thresh = (0u - n_choices) % n_choices;
        if (static_cast<uint32_t>(m) < thresh) continue;        // redraw (next attempt = new counter)
      }
      id = 1 + static_cast<int>(m >> 32);
      if (alias_prob != nullptr) {
        // Walker alias over ALL v_size entries (entry 0 carries weight 0 in the table)
        uint64_t m2 = static_cast<uint64_t>(r.x) * v_size;
        int col = static_cast<int>(m2 >> 32);
        const float u = (r.y >> 8) * (1.0f / 16777216.0f);
        id = u < alias_prob[col] ? col : alias_idx[col];
        if (id <= 0) continue;
      }
      bool excluded = id == v0 || id == t0;
      for (int h = 0; h < n_hist && !excluded; ++h) excluded = s_hist[h] == id;
      if (!excluded) break;
    }
    out[b * num_n + slot] = id;
  }
  rg::pdl_trigger();
}

}  // namespace rg

namespace rg {
__global__ void counter_add_kernel(uint64_t* ctr, uint64_t inc) {
  rg::pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) *ctr += inc;
  rg::pdl_trigger();
}
}  // namespace rg

extern "C" int recguru_counter_add(uint64_t* counter, uint64_t inc, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(counter != nullptr, RECGURU_E_SHAPE, "counter_add: null counter");
  rg::launch_pdl((counter_add_kernel), 1, 32, 0, static_cast<cudaStream_t>(stream), counter, inc);
  RG_CHECK_LAUNCH("counter_add");
  return RECGURU_OK;
}

static int sample_negatives_impl(const int64_t* history, int64_t B, int64_t H, int64_t pad_index,
                                 const int64_t* val, const int64_t* test, int64_t v_size,
                                 const float* alias_prob, const int32_t* alias_idx, int64_t num_n,
                                 uint64_t seed, uint64_t step, const uint64_t* step_dev,
                                 const int64_t* offsets, const int64_t* user_idx,
                                 int64_t* n_items_out, recguru_stream_t stream);

extern "C" int recguru_sample_negatives(const int64_t* history, int64_t B, int64_t H, int64_t pad_index,
                                        const int64_t* val, const int64_t* test, int64_t v_size,
                                        const float* alias_prob, const int32_t* alias_idx,
                                        int64_t num_n, uint64_t seed, uint64_t step,
                                        int64_t* n_items_out, recguru_stream_t stream) {
  return sample_negatives_impl(history, B, H, pad_index, val, test, v_size, alias_prob, alias_idx, num_n, seed,
                               step, nullptr, nullptr, nullptr, n_items_out, stream);
}

extern "C" int recguru_sample_negatives_devstep(const int64_t* history, int64_t B, int64_t H, int64_t pad_index,
                                                const int64_t* val, const int64_t* test, int64_t v_size,
                                                const float* alias_prob, const int32_t* alias_idx,
                                                int64_t num_n, uint64_t seed, const uint64_t* step_dev,
                                                int64_t* n_items_out, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(step_dev != nullptr, RECGURU_E_SHAPE, "sample_negatives_devstep: null step counter");
  return sample_negatives_impl(history, B, H, pad_index, val, test, v_size, alias_prob, alias_idx, num_n, seed,
                               0, step_dev, nullptr, nullptr, n_items_out, stream);
}

extern "C" int recguru_sample_negatives_csr(const int64_t* items, const int64_t* offsets, const int64_t* user_idx,
                                            int64_t B, int64_t max_len, const int64_t* val, const int64_t* test,
                                            int64_t v_size, const float* alias_prob, const int32_t* alias_idx,
                                            int64_t num_n, uint64_t seed, uint64_t step,
                                            const uint64_t* step_dev, int64_t* n_items_out,
                                            recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(offsets != nullptr, RECGURU_E_SHAPE, "sample_negatives_csr: null offsets");
  return sample_negatives_impl(items, B, max_len, /*pad_index=*/0, val, test, v_size, alias_prob, alias_idx, num_n,
                               seed, step, step_dev, offsets, user_idx, n_items_out, stream);
}

static int sample_negatives_impl(const int64_t* history, int64_t B, int64_t H, int64_t pad_index,
                                 const int64_t* val, const int64_t* test, int64_t v_size,
                                 const float* alias_prob, const int32_t* alias_idx, int64_t num_n,
                                 uint64_t seed, uint64_t step, const uint64_t* step_dev,
                                 const int64_t* offsets, const int64_t* user_idx,
                                 int64_t* n_items_out, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(B >= 0 && H >= 0 && num_n >= 0 && v_size >= 2 && v_size < (1ll << 31), RECGURU_E_SHAPE,
             "sample_negatives: bad sizes B=%lld H=%lld num_n=%lld v_size=%lld", (long long)B, (long long)H,
             (long long)num_n, (long long)v_size);
  RG_REQUIRE(H <= 12000, RECGURU_E_UNSUPPORTED, "sample_negatives: history length %lld > 12000", (long long)H);
  RG_REQUIRE((alias_prob == nullptr) == (alias_idx == nullptr), RECGURU_E_SHAPE,
             "sample_negatives: alias_prob and alias_idx come together");
  if (B == 0 || num_n == 0) return RECGURU_OK;
  RG_REQUIRE((H == 0 || history) && n_items_out, RECGURU_E_SHAPE, "sample_negatives: null pointer");
  rg::launch_pdl((sample_negatives_kernel), static_cast<unsigned>(B), kSamplerThreads, static_cast<size_t>(H) * 4 + 16,
                 static_cast<cudaStream_t>(stream), history, static_cast<int>(H), pad_index, val, test,
                 static_cast<uint32_t>(v_size), alias_prob, alias_idx, num_n, static_cast<uint32_t>(seed),
                 static_cast<uint32_t>(seed >> 32), static_cast<uint32_t>(step), step_dev, offsets, user_idx,
                 n_items_out);
  RG_CHECK_LAUNCH("sample_negatives");
  return RECGURU_OK;
}
