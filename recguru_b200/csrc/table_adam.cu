// the optimizer step of the item table.
//
// The reference trains with torch.optim.Adam(betas=(0.9, 0.99), eps=1e-9) over dense gradients
// (/root/reference/GURU/train_auto.py:354-364, Transformer/transformer.py:15-51 for the
// scheduled variant): every row of the [V1, d] table is updated every step, touched or not
// (moments of untouched rows decay with g = 0).  torch's default CUDA path is the foreach
// implementation: ~10 elementwise kernels, each streaming the table once.  This is the same
// arithmetic in ONE pass: read W, m, v, dW; write W, m, v  (28 B per element, HBM bound).
//
//   m  = m + (1 - beta1) * (g - m)                      (lerp_)
//   v  = v * beta2 + (1 - beta2) * g * g                (mul_ + addcmul_)
//   W  = W - step_size * m / (sqrt(v) / bc2_sqrt + eps) (sqrt, div, add_, addcdiv_)
// with step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t) computed by the host in fp64
// exactly as torch does.  maximize / amsgrad / weight_decay are not used by the reference.
#include "common.cuh"

namespace rg {

__device__ __forceinline__ void adam_elem(float& w, float& m, float& v, float g, float one_minus_b1,
                                          float b2, float one_minus_b2, float neg_step, float bc2_sqrt,
                                          float eps) {
  m = fmaf(one_minus_b1, g - m, m);
  v = fmaf(one_minus_b2 * g, g, v * b2);
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  w = fmaf(neg_step, m / denom, w);
}

__global__ void __launch_bounds__(256)
table_adam_kernel(float4* __restrict__ W, float4* __restrict__ M, float4* __restrict__ V,
                  const float4* __restrict__ G, int64_t n_vec, float one_minus_b1, float b2,
                  float one_minus_b2, float neg_step, float bc2_sqrt, float eps) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    float4 w = W[i], m = M[i], v = V[i];
    const float4 g = ld_stream_f4(G + i);
    adam_elem(w.x, m.x, v.x, g.x, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    adam_elem(w.y, m.y, v.y, g.y, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    adam_elem(w.z, m.z, v.z, g.z, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    adam_elem(w.w, m.w, v.w, g.w, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    W[i] = w; M[i] = m; V[i] = v;
  }
}

// Sparse-gradient variant: dW is a ROW STORE, valid only where row_epoch[row] == epoch (the rows
// the table backward wrote this step, recguru_embedding_bwd_mark_rows); every other row takes
// g = 0 through the same arithmetic -- what dense Adam does with a zero-filled gradient, bit for
// bit -- without the 4*V1*d zero fill of dW and without reading the untouched gradient rows.
__global__ void __launch_bounds__(256)
table_adam_rows_kernel(float4* __restrict__ W, float4* __restrict__ M, float4* __restrict__ V,
                       const float4* __restrict__ G, const uint32_t* __restrict__ row_epoch, uint32_t epoch,
                       int64_t n_vec, int vec, float one_minus_b1, float b2, float one_minus_b2,
                       float neg_step, float bc2_sqrt, float eps) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const int64_t row = i / vec;
    const bool live = row_epoch[row] == epoch;
    float4 w = W[i], m = M[i], v = V[i];
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) g = ld_stream_f4(G + i);
    adam_elem(w.x, m.x, v.x, g.x, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    adam_elem(w.y, m.y, v.y, g.y, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    adam_elem(w.z, m.z, v.z, g.z, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    adam_elem(w.w, m.w, v.w, g.w, one_minus_b1, b2, one_minus_b2, neg_step, bc2_sqrt, eps);
    W[i] = w; M[i] = m; V[i] = v;
  }
}

}  // namespace rg

extern "C" int recguru_table_adam_rows(float* weight, float* exp_avg, float* exp_avg_sq, const float* grad_rows,
                                       const uint32_t* row_epoch, uint32_t epoch, int64_t V1, int64_t d,
                                       double lr, double beta1, double beta2, double eps, int64_t step,
                                       recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(V1 >= 0 && d > 0 && d % 4 == 0, RECGURU_E_SHAPE, "table_adam_rows: d must be a positive multiple of 4");
  RG_REQUIRE(step >= 1, RECGURU_E_SHAPE, "table_adam_rows: step counts from 1");
  if (V1 == 0) return RECGURU_OK;
  RG_REQUIRE(weight && exp_avg && exp_avg_sq && grad_rows && row_epoch, RECGURU_E_SHAPE,
             "table_adam_rows: null pointer");
  RG_REQUIRE(aligned16(weight) && aligned16(exp_avg) && aligned16(exp_avg_sq) && aligned16(grad_rows),
             RECGURU_E_SHAPE, "table_adam_rows: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  const double step_size = lr / bc1;
  const int64_t n_vec = V1 * d / 4;
  const int64_t blocks = rg::min64(ceil_div(n_vec, 256 * 2), static_cast<int64_t>(num_sms()) * 16);
  table_adam_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(weight), reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
      reinterpret_cast<const float4*>(grad_rows), row_epoch, epoch, n_vec, static_cast<int>(d / 4),
      static_cast<float>(1.0 - beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta2),
      static_cast<float>(-step_size), static_cast<float>(sqrt(bc2)), static_cast<float>(eps));
  RG_CHECK_LAUNCH("table_adam_rows");
  return RECGURU_OK;
}

extern "C" int recguru_table_adam(float* weight, float* exp_avg, float* exp_avg_sq, const float* grad,
                                  int64_t n_floats, double lr, double beta1, double beta2, double eps,
                                  int64_t step, recguru_stream_t stream) {
  using namespace rg;
  RG_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, RECGURU_E_SHAPE, "table_adam: element count must be a multiple of 4");
  RG_REQUIRE(step >= 1, RECGURU_E_SHAPE, "table_adam: step counts from 1");
  if (n_floats == 0) return RECGURU_OK;
  RG_REQUIRE(weight && exp_avg && exp_avg_sq && grad, RECGURU_E_SHAPE, "table_adam: null pointer");
  RG_REQUIRE(aligned16(weight) && aligned16(exp_avg) && aligned16(exp_avg_sq) && aligned16(grad), RECGURU_E_SHAPE,
             "table_adam: buffers must be 16-byte aligned");
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  const double step_size = lr / bc1;
  const int64_t n_vec = n_floats / 4;
  const int64_t blocks = rg::min64(ceil_div(n_vec, 256 * 2), static_cast<int64_t>(num_sms()) * 16);
  table_adam_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(weight), reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
      reinterpret_cast<const float4*>(grad), n_vec, static_cast<float>(1.0 - beta1), static_cast<float>(beta2),
      static_cast<float>(1.0 - beta2), static_cast<float>(-step_size), static_cast<float>(sqrt(bc2)),
      static_cast<float>(eps));
  RG_CHECK_LAUNCH("table_adam");
  return RECGURU_OK;
}
