// K5 -- search-stage relation scorer, non-bag mode.
//
// Reference: Score/OutputLayer.forward (model.py:75-88) + train (main.py:641-673) called 100 times
// by score_relation_parallel (main.py:727-760):
//   pred[src] = max_{dst in N_r(src)} w[dst]   (first maximum in edge order)
//   loss      = mean_src (pred[src] - label[src])^2
//   w        <- clamp(Adam_lr(w, dloss/dw), 0, 1)
// The reference walks Python dicts per source every epoch; here a source is a CSR row of the
// relation and a destination a CSC bucket, both already built by K1.
//   score_fwd:    thread per row   -> argmax edge id + (pred - label), block-tree partial loss
//   score_loss:   one block        -> loss[epoch] (fixed-order sum)
//   score_update: thread per node  -> gradient gathered over the node's CSC bucket IN ORDER (an
//                 edge contributes iff it is its source's argmax edge), fused Adam + clamp
// No floating-point atomics: the result is identical run to run and on any GPU count.
#include "common.cuh"

namespace mpgnn {

constexpr int SC_THREADS = 256;

__global__ void __launch_bounds__(SC_THREADS) score_fwd_kernel(const int32_t* __restrict__ ptr,
                                                                const int32_t* __restrict__ idx,
                                                                const int32_t* __restrict__ eid, int64_t n,
                                                                const float* __restrict__ w,
                                                                const float* __restrict__ labels,
                                                                const uint8_t* __restrict__ src_mask,
                                                                int32_t* __restrict__ best_eid,
                                                                int32_t* __restrict__ best_dst,
                                                                float* __restrict__ diff, float* __restrict__ partial,
                                                                int32_t* __restrict__ partial_cnt) {
  __shared__ float sh[SC_THREADS];
  __shared__ int shc[SC_THREADS];
  const int64_t i = (int64_t)blockIdx.x * SC_THREADS + threadIdx.x;
  float sq = 0.f;
  int is_src = 0;
  if (i < n) {
    const int32_t b = ptr[i], e = ptr[i + 1];
    int32_t be = -1, bd = -1;
    float d = 0.f;
    // sources: every node with an edge of the relation (first iteration), or the given mask --
    // masked sources without such an edge predict 0 and still count in the mean (main.py:653-656)
    const bool masked_in = src_mask != nullptr && src_mask[i] != 0;
    if (e > b && (src_mask == nullptr || masked_in)) {
      float best = w[idx[b]];
      int32_t bp = b;
      for (int32_t p = b + 1; p < e; ++p) {
        const float v = w[idx[p]];
        if (v > best) {  // strict: the first maximum wins, like torch.argmax
          best = v;
          bp = p;
        }
      }
      be = eid[bp];
      bd = idx[bp];
      d = best - labels[i];
      sq = d * d;
      is_src = 1;
    } else if (masked_in) {
      d = 0.f - labels[i];
      sq = d * d;
      is_src = 1;
    }
    best_eid[i] = be;
    best_dst[i] = bd;
    diff[i] = d;
  }
  sh[threadIdx.x] = sq;
  shc[threadIdx.x] = is_src;
  __syncthreads();
  for (int o = SC_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[threadIdx.x] += sh[threadIdx.x + o];
      shc[threadIdx.x] += shc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = sh[0];
    partial_cnt[blockIdx.x] = shc[0];
  }
}

__global__ void __launch_bounds__(SC_THREADS) score_loss_kernel(const float* __restrict__ partial,
                                                                 const int32_t* __restrict__ partial_cnt, int n_partial,
                                                                 float* __restrict__ loss_out,
                                                                 float* __restrict__ inv_sources) {
  __shared__ float sh[SC_THREADS];
  __shared__ int shc[SC_THREADS];
  float v = 0.f;
  int c = 0;
  for (int i = threadIdx.x; i < n_partial; i += SC_THREADS) {
    v += partial[i];
    c += partial_cnt[i];
  }
  sh[threadIdx.x] = v;
  shc[threadIdx.x] = c;
  __syncthreads();
  for (int o = SC_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[threadIdx.x] += sh[threadIdx.x + o];
      shc[threadIdx.x] += shc[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float s = shc[0] > 0 ? (float)shc[0] : 1.f;
    *loss_out = sh[0] / s;
    *inv_sources = 1.f / s;
  }
}

__global__ void __launch_bounds__(SC_THREADS) score_update_kernel(
    const int32_t* __restrict__ cptr, const int32_t* __restrict__ cidx, const int32_t* __restrict__ ceid, int64_t n,
    const int32_t* __restrict__ best_eid, const float* __restrict__ diff, const float* __restrict__ inv_sources,
    float* __restrict__ w, float* __restrict__ m, float* __restrict__ v, float step_size, float inv_sqrt_bc2,
    float one_minus_b1, float b2, float one_minus_b2, float eps) {
  const int64_t j = (int64_t)blockIdx.x * SC_THREADS + threadIdx.x;
  if (j >= n) return;
  const float scale = 2.f * (*inv_sources);
  float g = 0.f;
  for (int32_t p = cptr[j]; p < cptr[j + 1]; ++p) {
    const int32_t src = cidx[p];
    if (best_eid[src] == ceid[p]) g += diff[src] * scale;  // this edge is its source's argmax edge
  }
  const float mj = m[j] + (g - m[j]) * one_minus_b1;
  const float vj = v[j] * b2 + one_minus_b2 * g * g;
  m[j] = mj;
  v[j] = vj;
  float wj = w[j] - step_size * (mj / (sqrtf(vj) * inv_sqrt_bc2 + eps));
  w[j] = fminf(fmaxf(wj, 0.f), 1.f);  // torch.clamp(weights, 0, 1) (main.py:667)
}

int64_t score_workspace_bytes(int64_t n) {
  const int64_t blocks = ceil_div(n, SC_THREADS);
  return align_up(n * 4, 256) * 3 + align_up(blocks * 4, 256) * 2 + 256;
}

int score_relation(const mpgnn_graph_impl* g, int64_t rel, float* w, const float* labels, const uint8_t* src_mask,
                   int64_t epochs, double lr,
                   float* m, float* v, float* loss_traj, int32_t* argmax_dst, void* ws_ptr, int64_t ws_bytes,
                   cudaStream_t s) {
  MPGNN_REQUIRE(g && w && labels && m && v && loss_traj && argmax_dst, MPGNN_EINVAL, "score_relation: NULL argument");
  MPGNN_REQUIRE(rel >= 0 && rel < g->r, MPGNN_ERANGE, "score_relation: relation %lld outside [0,%lld)",
                (long long)rel, (long long)g->r);
  MPGNN_REQUIRE(epochs >= 1, MPGNN_EINVAL, "score_relation: epochs must be >= 1");
  const int64_t n = g->n;
  const int blocks = (int)ceil_div(n, SC_THREADS);
  Workspace ws(ws_ptr, ws_bytes);
  int32_t* best_eid = ws.take<int32_t>(n);
  float* diff = ws.take<float>(n);
  float* inv_sources = ws.take<float>(1);
  float* partial = ws.take<float>(blocks);
  int32_t* partial_cnt = ws.take<int32_t>(blocks);
  MPGNN_REQUIRE(best_eid && diff && inv_sources && partial && partial_cnt, MPGNN_EINVAL,
                "score_relation: workspace too small");
  MPGNN_CUDA_CHECK(cudaMemsetAsync(m, 0, (size_t)n * 4, s));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(v, 0, (size_t)n * 4, s));
  const int32_t* ptr = g->csr_ptr + rel * n;
  const int32_t* cptr = g->csc_ptr + rel * n;
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  for (int64_t ep = 1; ep <= epochs; ++ep) {
    score_fwd_kernel<<<blocks, SC_THREADS, 0, s>>>(ptr, g->csr_idx, g->csr_eid, n, w, labels, src_mask, best_eid, argmax_dst, diff,
                                                  partial, partial_cnt);
    MPGNN_LAUNCH_CHECK();
    score_loss_kernel<<<1, SC_THREADS, 0, s>>>(partial, partial_cnt, blocks, loss_traj + (ep - 1), inv_sources);
    MPGNN_LAUNCH_CHECK();
    const double bc1 = 1.0 - pow(b1, (double)ep), bc2 = 1.0 - pow(b2, (double)ep);
    score_update_kernel<<<blocks, SC_THREADS, 0, s>>>(cptr, g->csc_idx, g->csc_eid, n, best_eid, diff, inv_sources, w, m,
                                                     v, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), (float)(1.0 - b1),
                                                     (float)b2, (float)(1.0 - b2), (float)eps);
    MPGNN_LAUNCH_CHECK();
  }
  return MPGNN_OK;
}

}  // namespace mpgnn

// ============================================================================================
// K5, bag mode (score_relation_bags_parallel / retrain_bags, main.py:814-917; OutputLayer.forward
// BAGS branch, model.py:45-72).  A bag is a list of source nodes; its prediction is
//   max over its sources s of  w[argmax_d (w[d] * a_s)] * a_s ,   a_s = <x[s], lin>
// with first-maximum ties, `lin` the 1 x F LinearLayerAttri weight.  MSE(mean) against the bag
// labels; Adam(lr) trains w (masked by grad_mask) and lin; both are clamped to [0,1].
//   bag_fwd:    thread per bag   -> prediction, (destination, source) that produced it, diff;
//               gradient contributions added as 2^48-scaled int64 (integer atomics are order
//               independent, so the sums are deterministic without a per-epoch sort)
//   bag_loss:   one block        -> loss[epoch]
//   bag_update: thread per node / feature -> Adam + clamp, accumulators cleared
// ============================================================================================
namespace mpgnn {

constexpr double kFix = 281474976710656.0;  // 2^48

__global__ void __launch_bounds__(SC_THREADS) bag_fwd_kernel(
    const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int32_t* __restrict__ bag_ptr,
    const int32_t* __restrict__ bag_src, int n_bags, const float* __restrict__ bag_labels,
    const float* __restrict__ x, int feat, const float* __restrict__ lin, const float* __restrict__ w,
    int32_t* __restrict__ best_dst, int32_t* __restrict__ best_src, float* __restrict__ diff,
    float* __restrict__ src_val, long long* __restrict__ acc_w, long long* __restrict__ acc_lin,
    float* __restrict__ partial) {
  __shared__ float sh[SC_THREADS];
  const int b = blockIdx.x * SC_THREADS + threadIdx.x;
  float sq = 0.f;
  if (b < n_bags) {
    float cur = -10.f, cur_a = 0.f;
    int32_t bd = -1, bs = -1;
    float pred = 0.f;
    for (int32_t q = bag_ptr[b]; q < bag_ptr[b + 1]; ++q) {
      const int32_t s = bag_src[q];
      float a = 0.f;
      for (int f = 0; f < feat; ++f) a = fmaf(x[(int64_t)s * feat + f], lin[f], a);
      const int32_t p0 = ptr[s], p1 = ptr[s + 1];
      if (p1 <= p0) continue;                      // cleaned bags only hold sources with an edge
      float best = w[idx[p0]] * a;
      int32_t bp = p0;
      for (int32_t p = p0 + 1; p < p1; ++p) {
        const float v = w[idx[p]] * a;
        if (v > best) {
          best = v;
          bp = p;
        }
      }
      const float val = w[idx[bp]] * a;
      src_val[s] = val;                            // same value from every bag that holds s
      if (val > cur) {
        cur = val;
        cur_a = a;
        pred = val;
        bd = idx[bp];
        bs = s;
      }
    }
    const float d = pred - bag_labels[b];
    best_dst[b] = bd;
    best_src[b] = bs;
    diff[b] = d;
    sq = d * d;
    if (bd >= 0) {
      const double gi = 2.0 * (double)d / (double)n_bags;
      atomicAdd(reinterpret_cast<unsigned long long*>(acc_w + bd),
                (unsigned long long)(long long)llrint(gi * (double)cur_a * kFix));
      const double wb = (double)w[bd];
      for (int f = 0; f < feat; ++f)
        atomicAdd(reinterpret_cast<unsigned long long*>(acc_lin + f),
                  (unsigned long long)(long long)llrint(gi * wb * (double)x[(int64_t)bs * feat + f] * kFix));
    }
  }
  sh[threadIdx.x] = sq;
  __syncthreads();
  for (int o = SC_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(SC_THREADS) bag_loss_kernel(const float* __restrict__ partial, int n_partial,
                                                               int n_bags, float* __restrict__ loss_out) {
  __shared__ float sh[SC_THREADS];
  float v = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += SC_THREADS) v += partial[i];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = SC_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = sh[0] / (float)n_bags;
}

__global__ void __launch_bounds__(SC_THREADS) bag_update_kernel(
    int64_t n, int feat, long long* __restrict__ acc_w, long long* __restrict__ acc_lin,
    const uint8_t* __restrict__ grad_mask, int use_mask, float* __restrict__ w, float* __restrict__ lin,
    float* __restrict__ m, float* __restrict__ v, float* __restrict__ ml, float* __restrict__ vl, float step_size,
    float inv_sqrt_bc2, float one_minus_b1, float b2, float one_minus_b2, float eps) {
  const int64_t j = (int64_t)blockIdx.x * SC_THREADS + threadIdx.x;
  if (j < n) {
    float g = (float)((double)acc_w[j] / kFix);
    acc_w[j] = 0;
    if (use_mask && grad_mask[j] == 0) g = 0.f;          // frozen destinations (main.py:663-664)
    const float mj = m[j] + (g - m[j]) * one_minus_b1;
    const float vj = v[j] * b2 + one_minus_b2 * g * g;
    m[j] = mj;
    v[j] = vj;
    const float wj = w[j] - step_size * (mj / (sqrtf(vj) * inv_sqrt_bc2 + eps));
    w[j] = fminf(fmaxf(wj, 0.f), 1.f);
  }
  if (j < feat) {
    const float g = (float)((double)acc_lin[j] / kFix);
    acc_lin[j] = 0;
    const float mj = ml[j] + (g - ml[j]) * one_minus_b1;
    const float vj = vl[j] * b2 + one_minus_b2 * g * g;
    ml[j] = mj;
    vl[j] = vj;
    const float lj = lin[j] - step_size * (mj / (sqrtf(vj) * inv_sqrt_bc2 + eps));
    lin[j] = fminf(fmaxf(lj, 0.f), 1.f);                  // main.py:668
  }
}

int64_t score_bags_workspace_bytes(int64_t n, int64_t n_bags, int64_t feat) {
  const int64_t blocks = ceil_div(n_bags > 0 ? n_bags : 1, SC_THREADS);
  return align_up(n * 8, 256) + align_up(feat * 8, 256) + align_up(n * 4, 256) * 2 + align_up(feat * 4, 256) * 2 +
         align_up(blocks * 4, 256) + 256;
}

int score_bags(const mpgnn_graph_impl* g, int64_t rel, const int32_t* bag_ptr, const int32_t* bag_src, int64_t n_bags,
               const float* bag_labels, const float* x, int64_t feat, float* w, float* lin, const uint8_t* grad_mask,
               int use_mask, int64_t epochs, double lr, float* loss_traj, int32_t* best_dst, int32_t* best_src,
               float* diff, float* src_val, void* ws_ptr, int64_t ws_bytes, cudaStream_t s) {
  MPGNN_REQUIRE(g && bag_ptr && bag_src && bag_labels && x && w && lin && loss_traj && best_dst && best_src && diff &&
                    src_val, MPGNN_EINVAL, "score_bags: NULL argument");
  MPGNN_REQUIRE(rel >= 0 && rel < g->r, MPGNN_ERANGE, "score_bags: relation %lld outside [0,%lld)", (long long)rel,
                (long long)g->r);
  MPGNN_REQUIRE(n_bags >= 1 && feat >= 1 && feat <= SC_THREADS && epochs >= 1 && (!use_mask || grad_mask), MPGNN_EINVAL,
                "score_bags: bad sizes");
  const int64_t n = g->n;
  const int blocks_b = (int)ceil_div(n_bags, SC_THREADS), blocks_n = (int)ceil_div(n, SC_THREADS);
  Workspace ws(ws_ptr, ws_bytes);
  long long* acc_w = ws.take<long long>(n);
  long long* acc_lin = ws.take<long long>(feat);
  float* m = ws.take<float>(n);
  float* v = ws.take<float>(n);
  float* ml = ws.take<float>(feat);
  float* vl = ws.take<float>(feat);
  float* partial = ws.take<float>(blocks_b);
  MPGNN_REQUIRE(acc_w && acc_lin && m && v && ml && vl && partial, MPGNN_EINVAL, "score_bags: workspace too small");
  MPGNN_CUDA_CHECK(cudaMemsetAsync(acc_w, 0, (size_t)n * 8, s));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(acc_lin, 0, (size_t)feat * 8, s));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(m, 0, (size_t)n * 4, s));      // a fresh optimiser per restart (main.py:887)
  MPGNN_CUDA_CHECK(cudaMemsetAsync(v, 0, (size_t)n * 4, s));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(ml, 0, (size_t)feat * 4, s));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(vl, 0, (size_t)feat * 4, s));
  const int32_t* ptr = g->csr_ptr + rel * n;
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  for (int64_t ep = 1; ep <= epochs; ++ep) {
    bag_fwd_kernel<<<blocks_b, SC_THREADS, 0, s>>>(ptr, g->csr_idx, bag_ptr, bag_src, (int)n_bags, bag_labels, x,
                                                  (int)feat, lin, w, best_dst, best_src, diff, src_val, acc_w, acc_lin,
                                                  partial);
    MPGNN_LAUNCH_CHECK();
    bag_loss_kernel<<<1, SC_THREADS, 0, s>>>(partial, blocks_b, (int)n_bags, loss_traj + (ep - 1));
    MPGNN_LAUNCH_CHECK();
    if (ep == epochs) {
      // the reference reads predictions / argmax maps of the LAST forward, i.e. before the last update:
      // best_dst/best_src/diff/src_val are already final; the update below only moves w and lin
    }
    const double bc1 = 1.0 - pow(b1, (double)ep), bc2 = 1.0 - pow(b2, (double)ep);
    bag_update_kernel<<<blocks_n, SC_THREADS, 0, s>>>(n, (int)feat, acc_w, acc_lin, grad_mask, use_mask, w, lin, m, v,
                                                     ml, vl, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)),
                                                     (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)eps);
    MPGNN_LAUNCH_CHECK();
  }
  return MPGNN_OK;
}

}  // namespace mpgnn
