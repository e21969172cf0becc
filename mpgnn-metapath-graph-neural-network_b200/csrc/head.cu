// MPNetm head pieces that are not GEMMs: log_softmax + nll on an index set and its
// gradient (model.py:226, main.py:1065), device-side macro-F1 (K6; main.py:1090-1099) and
// the fused Adam step (main.py:1119).  All reductions are fixed-order (deterministic).
#include "common.cuh"

namespace mpgnn {

constexpr int MAX_CLASSES = 64;

__global__ void logsoftmax_kernel(const float* __restrict__ logits, int64_t n, int c, float* __restrict__ logp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* row = logits + i * c;
  float mx = row[0];
  for (int j = 1; j < c; ++j) mx = fmaxf(mx, row[j]);
  float s = 0.f;
  for (int j = 0; j < c; ++j) s += expf(row[j] - mx);
  const float lse = logf(s);
  for (int j = 0; j < c; ++j) logp[i * c + j] = (row[j] - mx) - lse;
}

// stage 1: per-block partial sums of -logp[idx[i], y[i]] in a fixed tree; stage 2: one block.
constexpr int NLL_THREADS = 256;

__device__ __forceinline__ float block_sum_fixed(float v, float* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = NLL_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  return sh[0];
}

__global__ void __launch_bounds__(NLL_THREADS) nll_partial_kernel(const float* __restrict__ logp, int c,
                                                                   const int64_t* __restrict__ idx,
                                                                   const int64_t* __restrict__ y, int64_t n_idx,
                                                                   float* __restrict__ partial) {
  __shared__ float sh[NLL_THREADS];
  float v = 0.f;
  // contiguous chunk per block, strided inside the block: a pure function of (n_idx, grid)
  const int64_t per_block = (n_idx + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = (int64_t)blockIdx.x * per_block;
  const int64_t b1 = min(n_idx, b0 + per_block);
  for (int64_t i = b0 + threadIdx.x; i < b1; i += NLL_THREADS) v -= logp[idx[i] * c + y[i]];
  const float s = block_sum_fixed(v, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(NLL_THREADS) nll_final_kernel(const float* __restrict__ partial, int n_partial,
                                                                 int64_t n_idx, float* __restrict__ loss) {
  __shared__ float sh[NLL_THREADS];
  float v = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += NLL_THREADS) v += partial[i];
  const float s = block_sum_fixed(v, sh);
  if (threadIdx.x == 0) *loss = s / (float)n_idx;
}

__global__ void nll_grad_kernel(const float* __restrict__ logp, int c, const int64_t* __restrict__ idx,
                                const int64_t* __restrict__ y, int64_t n_idx, float* __restrict__ glogits) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_idx) return;
  const int64_t row = idx[i];
  const int64_t yi = y[i];
  const float inv_n = 1.0f / (float)n_idx;
  for (int j = 0; j < c; ++j) {
    const float g = (expf(logp[row * c + j]) - (j == yi ? 1.f : 0.f)) * inv_n;
    atomicAdd(glogits + row * c + j, g);  // one add per element when idx has no duplicates
  }
}

int launch_logsoftmax_nll(const float* logits, int64_t n, int64_t c, const int64_t* idx, const int64_t* y,
                          int64_t n_idx, float* logp, float* loss, float* glogits, void* ws, int64_t ws_bytes,
                          cudaStream_t s) {
  MPGNN_REQUIRE(c >= 1 && c <= MAX_CLASSES, MPGNN_ENOTSUP, "logsoftmax_nll: %lld classes unsupported (max %d)",
                (long long)c, MAX_CLASSES);
  MPGNN_REQUIRE(logits && logp, MPGNN_EINVAL, "logsoftmax_nll: NULL logits/logp");
  if (n > 0) {
    logsoftmax_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(logits, n, (int)c, logp);
    MPGNN_LAUNCH_CHECK();
  }
  if (loss != nullptr) {
    MPGNN_REQUIRE(n_idx > 0 && idx && y, MPGNN_EINVAL, "logsoftmax_nll: empty index set");
    int64_t want = ceil_div(n_idx, NLL_THREADS * 4);
    int blocks = want > 1024 ? 1024 : (want < 1 ? 1 : (int)want);
    Workspace w(ws, ws_bytes);
    float* partial = w.take<float>(blocks);
    MPGNN_REQUIRE(partial != nullptr, MPGNN_EINVAL, "logsoftmax_nll: workspace too small (need %d floats)", blocks);
    nll_partial_kernel<<<blocks, NLL_THREADS, 0, s>>>(logp, (int)c, idx, y, n_idx, partial);
    MPGNN_LAUNCH_CHECK();
    nll_final_kernel<<<1, NLL_THREADS, 0, s>>>(partial, blocks, n_idx, loss);
    MPGNN_LAUNCH_CHECK();
  }
  if (glogits != nullptr) {
    MPGNN_REQUIRE(n_idx > 0 && idx && y, MPGNN_EINVAL, "logsoftmax_nll: empty index set");
    MPGNN_CUDA_CHECK(cudaMemsetAsync(glogits, 0, (size_t)(n * c) * sizeof(float), s));
    nll_grad_kernel<<<(unsigned)ceil_div(n_idx, 256), 256, 0, s>>>(logp, (int)c, idx, y, n_idx, glogits);
    MPGNN_LAUNCH_CHECK();
  }
  return MPGNN_OK;
}

// ---- macro-F1 -----------------------------------------------------------------------------
__global__ void confusion_kernel(const float* __restrict__ logp, int c, const int64_t* __restrict__ idx,
                                 const int64_t* __restrict__ y, int64_t n_idx, int32_t* __restrict__ cm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_idx) return;
  const float* row = logp + idx[i] * c;
  int best = 0;
  float bv = row[0];
  for (int j = 1; j < c; ++j)
    if (row[j] > bv) {  // first maximum wins, like torch.argmax
      bv = row[j];
      best = j;
    }
  const int64_t t = y[i];
  if (t >= 0 && t < c) atomicAdd(cm + t * c + best, 1);  // integer counts: order independent
}

__global__ void f1_from_confusion_kernel(const int32_t* __restrict__ cm, int c, double* __restrict__ f1) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sum = 0.0;
  int present = 0;
  for (int k = 0; k < c; ++k) {
    long long tp = cm[k * c + k], row = 0, col = 0;
    for (int j = 0; j < c; ++j) {
      row += cm[k * c + j];  // true == k
      col += cm[j * c + k];  // pred == k
    }
    if (row + col == 0) continue;  // label in neither array: not part of sklearn's label union
    ++present;
    const long long den = row + col;  // 2tp + fp + fn
    sum += den > 0 ? (2.0 * (double)tp) / (double)den : 0.0;
  }
  *f1 = present > 0 ? sum / (double)present : 0.0;
}

int launch_macro_f1(const float* logp, int64_t c, const int64_t* idx, const int64_t* y, int64_t n_idx, int32_t* cm,
                    double* f1, cudaStream_t s) {
  MPGNN_REQUIRE(c >= 1 && c <= MAX_CLASSES, MPGNN_ENOTSUP, "macro_f1: %lld classes unsupported", (long long)c);
  MPGNN_REQUIRE(logp && idx && y && cm && f1 && n_idx > 0, MPGNN_EINVAL, "macro_f1: bad arguments");
  MPGNN_CUDA_CHECK(cudaMemsetAsync(cm, 0, (size_t)(c * c) * sizeof(int32_t), s));
  confusion_kernel<<<(unsigned)ceil_div(n_idx, 256), 256, 0, s>>>(logp, (int)c, idx, y, n_idx, cm);
  MPGNN_LAUNCH_CHECK();
  f1_from_confusion_kernel<<<1, 32, 0, s>>>(cm, (int)c, f1);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

// ---- Adam ---------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float step_size, float inv_sqrt_bc2,
                            float one_minus_beta1, float beta2, float one_minus_beta2, float eps, float wd) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);                 // grad.add(param, alpha=weight_decay)
    const float mi = m[i] + (gi - m[i]) * one_minus_beta1;  // exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * beta2 + one_minus_beta2 * gi * gi;  // mul_(beta2).addcmul_(g, g, 1-beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, int64_t step, double lr, double beta1,
                double beta2, double eps, double wd, cudaStream_t s) {
  if (n <= 0) return MPGNN_OK;
  MPGNN_REQUIRE(step >= 1 && p && g && m && v, MPGNN_EINVAL, "adam: bad arguments");
  // scalars are formed in double exactly as torch.optim.Adam forms them in Python, then narrowed
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  int64_t blocks = ceil_div(n, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(p, g, m, v, n, step_size, inv_sqrt_bc2, (float)(1.0 - beta1),
                                               (float)beta2, (float)(1.0 - beta2), (float)eps, (float)wd);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

}  // namespace mpgnn
