// Device-resident candidate trainer: one call scores one candidate metapath the way
// mpgnn_parallel_multiple does (main.py:1117-1134) -- `epochs` x (mpgnn_train, mpgnn_validation) of
// an MPNetm with ONE metapath (model.py:179-228) -- without returning to the host between epochs.
// The whole epoch (train forward, loss, backward, Adam, eval forward, validation loss, macro-F1)
// is captured once in a CUDA graph and replayed: configs C1-C3 are launch-bound, so
// the per-epoch cost becomes one graph launch instead of ~60 kernel launches + a host sync.
// Everything that changes from epoch to epoch (Adam step, dropout offsets, trace slot) lives in
// device words advanced by the first kernel of the graph.
#include <mutex>
#include <vector>

#include "common.cuh"

namespace mpgnn {

int64_t hop_workspace_bytes(int64_t n, int64_t f_in, int64_t f_out);
int hop_fwd(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t f_in, const float* w, const float* root,
            const float* bias, int64_t f_out, uint32_t flags, double p, uint64_t seed, uint64_t offset,
            const uint8_t* mask_bits, float* h, float* y, uint32_t* actmask, void* ws_ptr, int64_t ws_bytes,
            cudaStream_t s, const uint64_t* offset_ptr, bool h_precomputed);
int hop_bwd(const mpgnn_graph_impl* g, int64_t rel, const float* x, const float* h, const float* y,
            const uint32_t* actmask, const float* gy, int64_t f_in, const float* w, const float* root, int64_t f_out,
            uint32_t flags, double p, float* gx, float* gw, float* groot, float* gbias, void* ws_ptr, int64_t ws_bytes,
            cudaStream_t s);
int proj_tcgen05_supported(int64_t m, int64_t k1, int64_t k2, int64_t n, uint32_t flags);
int64_t proj_tcgen05_workspace_floats(int64_t k, int64_t n);
int launch_proj_tcgen05_ws(const GemmRowsArgs& a, uint32_t flags, float* b_img, cudaStream_t s);
int wgrad_tcgen05_supported(int64_t m, int64_t k1, int64_t k2, int64_t n, uint32_t flags);
int64_t wgrad_tcgen05_workspace_floats(int64_t m, int64_t n);
int launch_wgrad_tcgen05(const GemmTnArgs& a, float* ws, cudaStream_t s);
int launch_logsoftmax_nll(const float* logits, int64_t n, int64_t c, const int64_t* idx, const int64_t* y,
                          int64_t n_idx, float* logp, float* loss, float* glogits, void* ws, int64_t ws_bytes,
                          cudaStream_t s);
int launch_macro_f1(const float* logp, int64_t c, const int64_t* idx, const int64_t* y, int64_t n_idx, int32_t* cm,
                    double* f1, cudaStream_t s);

constexpr int kMaxPaths = 8;     // metapaths of one model (final_selection trains unions of up to 3, main.py:1465)
constexpr int kMaxLayers = 32;   // conv layers over all metapaths

struct TrainerState {      // device words advanced once per epoch
  int64_t epoch;           // 1-based after the bump
  uint64_t drop_off[kMaxLayers];   // per-layer dropout offset of this epoch: epoch * n_layers + layer
};

__global__ void trainer_bump_kernel(TrainerState* st, int n_layers) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->epoch += 1;
    for (int k = 0; k < n_layers; ++k) st->drop_off[k] = (uint64_t)st->epoch * (uint64_t)n_layers + (uint64_t)k;
  }
}

// torch.optim.Adam step with the step count read from the device (graph replay safe)
__global__ void trainer_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, int64_t n, const TrainerState* __restrict__ st, double lr,
                                    double b1, double b2, double eps, double wd) {
  const double t = (double)st->epoch;
  const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
  const float step_size = (float)(lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float omb1 = (float)(1.0 - b1), fb2 = (float)b2, omb2 = (float)(1.0 - b2), feps = (float)eps, fwd = (float)wd;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gi = g[i];
    const float pi = p[i];
    if (fwd != 0.f) gi = fmaf(fwd, pi, gi);
    const float mi = m[i] + (gi - m[i]) * omb1;
    const float vi = v[i] * fb2 + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + feps));
  }
}

// validated == 0: an epoch without the validation pass (MPGNN_TRAINER_VALIDATE_LAST): NaN in the three columns it fills
__global__ void trainer_trace_kernel(const TrainerState* st, const float* loss_train, const float* loss_val,
                                     const double* f1_train, const double* f1_val, double* trace, int64_t capacity,
                                     int validated) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int64_t e = st->epoch - 1;
    if (e >= 0 && e < capacity) {
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      trace[4 * e + 0] = (double)*loss_train;
      trace[4 * e + 1] = validated ? (double)*loss_val : nan;
      trace[4 * e + 2] = validated ? *f1_train : nan;
      trace[4 * e + 3] = validated ? *f1_val : nan;
    }
  }
}

struct Trainer {
  const mpgnn_graph_impl* g;
  const float* x;
  int64_t n, f_in, hidden, classes;
  int n_layers;                        // conv layers over all metapaths (flat index l = path_off[i] + k)
  int n_paths;
  int path_off[kMaxPaths + 1];
  int64_t rel[kMaxLayers];
  int64_t cat;                         // width of the concatenated embedding = hidden * n_paths (model.py:220)
  const int64_t *train_idx, *train_y, *val_idx, *val_y;
  int64_t n_train, n_val;
  double dropout_p;
  uint64_t seed;
  uint32_t flags;
  // parameters: per layer [W (fin x H), root (fin x H), bias (H)], then W1t (H x H), b1 (H), W2t (H x C), b2 (C)
  int64_t n_params;
  int64_t off_w[kMaxLayers], off_root[kMaxLayers], off_bias[kMaxLayers], off_w1, off_b1, off_w2, off_b2;
  float *params, *grads, *adam_m, *adam_v;
  float *h[kMaxLayers], *y[kMaxLayers];   // aggregated inputs and activations per layer
  float* am[kMaxLayers];        // activation bitmasks [y > 0] (uint32 words) when hidden % 32 == 0, else null
  float *gxa, *gxb;             // ping-pong activation gradients
  float *emb, *gemb;            // [N, cat] concatenated embeddings and their gradient (n_paths > 1 only)
  bool h0_ready;                // the first layer's aggregation mean_r(x) is a constant of (graph, relation, x): done once
  float *a1, *lg, *logp, *glg, *gz1, *packed;
  float *loss_train, *loss_val;
  double *f1_train, *f1_val, *trace;
  int64_t trace_capacity;
  int32_t* cm;
  TrainerState* st;
  void* ws;
  int64_t ws_bytes;
  void* slab;                // the one device allocation every buffer above is carved from
  size_t slab_bytes;
  int slab_device;
  cudaGraphExec_t exec;       // train step + validation
  cudaGraphExec_t exec_train; // train step only
  cudaStream_t own_stream;   // capture is not allowed on the legacy default stream
  double cap_lr, cap_b1, cap_b2, cap_eps, cap_wd;
  bool params_set;           // run()/evaluate() refuse to work on the recycled slab's stale bytes
};

// split-K partials of the head's two weight gradients: [gW2t; gb2] = a1^T glg (H+1 x C) and [gW1t; gb1] = E^T g_z1
// (cat+1 x H).  The split is a function of the shape and NOT monotone in it (thin outputs get more, shorter row
// ranges), so the buffer is sized from the shapes actually launched.
static int64_t trainer_tn_floats(const Trainer& t) {
  const int64_t a = gemm_tn_partial_floats(t.n, t.hidden + 1, t.classes);
  const int64_t b = gemm_tn_partial_floats(t.n, t.cat + 1, t.hidden);
  return a > b ? a : b;
}

static int64_t trainer_ws_bytes(const Trainer& t) {
  int64_t a = hop_workspace_bytes(t.n, t.f_in, t.hidden), b = hop_workspace_bytes(t.n, t.hidden, t.hidden);
  int64_t hop = a > b ? a : b;
  int64_t tn = trainer_tn_floats(t) * 4 + 4096;
  return (hop > tn ? hop : tn) + 1024 * 8;
}

// ---- trainer memory -------------------------------------------------------------------------------------------------
// A candidate trainer needs ~(4 L + 8) N H floats in about forty buffers and lives for one candidate (0.7 s at the
// configs[1] size).  Forty cudaMalloc + forty cudaFree per candidate cost up to a second of driver time (unmapping
// half a gigabyte synchronises the device), more than the training itself, so every trainer takes ONE slab and
// slabs are recycled through a small process-wide cache instead of going back to the driver.
struct SlabCache {
  struct Entry { void* p; size_t bytes; int device; };
  std::mutex mu;
  std::vector<Entry> free_list;
  static constexpr size_t kMaxEntries = 16;
  static constexpr size_t kMaxBytes = (size_t)48 << 30;      // of a 180 GB part
  size_t cached_bytes = 0;

  cudaError_t get(size_t bytes, int device, void** p, size_t* got) {
    {
      std::lock_guard<std::mutex> lk(mu);
      int best = -1;
      for (int i = 0; i < (int)free_list.size(); ++i) {
        const Entry& e = free_list[i];
        if (e.device == device && e.bytes >= bytes && e.bytes <= bytes + bytes / 2 &&
            (best < 0 || e.bytes < free_list[best].bytes))
          best = i;
      }
      if (best >= 0) {
        *p = free_list[best].p; *got = free_list[best].bytes;
        cached_bytes -= free_list[best].bytes;
        free_list.erase(free_list.begin() + best);
        return cudaSuccess;
      }
    }
    cudaError_t ce = cudaMalloc(p, bytes);
    if (ce == cudaErrorMemoryAllocation) {       // give the cached slabs back and retry once
      (void)cudaGetLastError();
      release(device);
      ce = cudaMalloc(p, bytes);
    }
    *got = bytes;
    return ce;
  }
  void put(void* p, size_t bytes, int device) {
    if (p == nullptr) return;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (free_list.size() < kMaxEntries && cached_bytes + bytes <= kMaxBytes) {
        free_list.push_back({p, bytes, device});
        cached_bytes += bytes;
        return;
      }
    }
    cudaFree(p);
  }
  void release(int device) {
    std::vector<Entry> drop;
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t i = 0; i < free_list.size();) {
        if (free_list[i].device == device) {
          drop.push_back(free_list[i]);
          cached_bytes -= free_list[i].bytes;
          free_list.erase(free_list.begin() + i);
        } else {
          ++i;
        }
      }
    }
    for (const Entry& e : drop) cudaFree(e.p);
  }
};
static SlabCache g_slabs;

// Hands out 256-byte aligned pieces of a slab; with base == nullptr it only measures.
struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  void take(T** p, int64_t count) {
    const size_t bytes = ((size_t)(count > 0 ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    if (base != nullptr) *p = reinterpret_cast<T*>(base + off);
    off += bytes;
  }
};

static void trainer_layout(Trainer* t, Carver& c) {
  const int64_t n = t->n, hidden = t->hidden, classes = t->classes, np = t->n_params;
  c.take(&t->params, np); c.take(&t->grads, np); c.take(&t->adam_m, np); c.take(&t->adam_v, np);
  for (int i = 0; i < t->n_paths; ++i)
    for (int l = t->path_off[i]; l < t->path_off[i + 1]; ++l) {
      c.take(&t->h[l], n * (l == t->path_off[i] ? t->f_in : hidden));
      c.take(&t->y[l], n * hidden);
      if (hidden % 32 == 0) c.take(&t->am[l], n * (hidden / 32));
    }
  c.take(&t->gxa, n * hidden); c.take(&t->gxb, n * hidden);
  if (t->n_paths > 1) { c.take(&t->emb, n * t->cat); c.take(&t->gemb, n * t->cat); }
  c.take(&t->a1, n * hidden); c.take(&t->lg, n * classes); c.take(&t->logp, n * classes);
  c.take(&t->glg, n * classes); c.take(&t->gz1, n * hidden);
  c.take(&t->packed, hidden * (t->cat > classes ? t->cat : classes));
  c.take(&t->loss_train, 1); c.take(&t->loss_val, 1);
  c.take(&t->f1_train, 1); c.take(&t->f1_val, 1); c.take(&t->trace, 4 * t->trace_capacity);
  c.take(&t->cm, classes * classes); c.take(&t->st, 1);
  char* ws = nullptr;
  c.take(&ws, t->ws_bytes);
  t->ws = ws;
}

void trainer_free(Trainer* t) {
  if (!t) return;
  if (t->exec) cudaGraphExecDestroy(t->exec);
  if (t->exec_train) cudaGraphExecDestroy(t->exec_train);
  if (t->own_stream) cudaStreamDestroy(t->own_stream);
  if (t->slab != nullptr) {
    // the slab may be handed to the next trainer at once: everything queued on it must have finished (what the
    // implicit synchronisation of cudaFree used to guarantee)
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != t->slab_device) cudaSetDevice(t->slab_device);
    cudaDeviceSynchronize();
    g_slabs.put(t->slab, t->slab_bytes, t->slab_device);
    if (cur != t->slab_device) cudaSetDevice(cur);
  }
  delete t;
}

int trainer_create(const mpgnn_graph_impl* g, const float* x, int64_t f_in, int64_t hidden, int64_t classes,
                   const int64_t* h_rel, const int64_t* h_path_ptr, int64_t n_paths, const int64_t* train_idx,
                   const int64_t* train_y, int64_t n_train, const int64_t* val_idx, const int64_t* val_y, int64_t n_val,
                   double dropout_p, uint64_t seed, uint32_t flags, int64_t max_epochs, Trainer** out) {
  MPGNN_REQUIRE(g && x && h_rel && h_path_ptr && train_idx && train_y && val_idx && val_y && out, MPGNN_EINVAL, "trainer: NULL argument");
  MPGNN_REQUIRE(n_paths >= 1 && n_paths <= kMaxPaths, MPGNN_ENOTSUP, "trainer: %lld metapaths outside [1,%d]", (long long)n_paths, kMaxPaths);
  MPGNN_REQUIRE(h_path_ptr[0] == 0, MPGNN_EINVAL, "trainer: path_ptr[0] must be 0");
  const int64_t n_layers = h_path_ptr[n_paths];
  MPGNN_REQUIRE(n_layers >= 1 && n_layers <= kMaxLayers, MPGNN_ENOTSUP, "trainer: %lld conv layers outside [1,%d]", (long long)n_layers, kMaxLayers);
  for (int64_t i = 0; i < n_paths; ++i)
    MPGNN_REQUIRE(h_path_ptr[i + 1] - h_path_ptr[i] >= 1 && h_path_ptr[i + 1] - h_path_ptr[i] <= 8, MPGNN_ENOTSUP,
                  "trainer: metapath length %lld outside [1,8]", (long long)(h_path_ptr[i + 1] - h_path_ptr[i]));
  MPGNN_REQUIRE(f_in >= 1 && hidden >= 1 && classes >= 1 && classes <= 64 && n_train > 0 && n_val > 0, MPGNN_EINVAL, "trainer: bad sizes");
  MPGNN_REQUIRE(dropout_p >= 0.0 && dropout_p < 1.0, MPGNN_EINVAL, "trainer: dropout p=%g", dropout_p);
  MPGNN_REQUIRE(!(flags & MPGNN_F_BF16), MPGNN_ENOTSUP, "trainer: MPGNN_F_BF16 is not built; use MPGNN_F_TF32X3");
  for (int64_t k = 0; k < n_layers; ++k)
    MPGNN_REQUIRE(h_rel[k] >= 0 && h_rel[k] < g->r, MPGNN_ERANGE, "trainer: relation %lld outside [0,%lld)", (long long)h_rel[k], (long long)g->r);
  Trainer* t = new Trainer();
  memset(t, 0, sizeof(*t));
  t->g = g; t->x = x; t->n = g->n; t->f_in = f_in; t->hidden = hidden; t->classes = classes;
  t->n_layers = (int)n_layers; t->n_paths = (int)n_paths; t->cat = hidden * n_paths;
  for (int i = 0; i <= n_paths; ++i) t->path_off[i] = (int)h_path_ptr[i];
  for (int k = 0; k < n_layers; ++k) t->rel[k] = h_rel[k];
  t->train_idx = train_idx; t->train_y = train_y; t->n_train = n_train;
  t->val_idx = val_idx; t->val_y = val_y; t->n_val = n_val;
  t->dropout_p = dropout_p; t->seed = seed; t->flags = flags;
  int64_t off = 0;
  for (int i = 0; i < n_paths; ++i)
    for (int l = t->path_off[i]; l < t->path_off[i + 1]; ++l) {      // state_dict order: layers_list.{i}.{k}.{weight,root,bias}
      const int64_t fi = l == t->path_off[i] ? f_in : hidden;
      t->off_w[l] = off; off += fi * hidden;
      t->off_root[l] = off; off += fi * hidden;
      t->off_bias[l] = off; off += hidden;
    }
  t->off_w1 = off; off += hidden * t->cat;
  t->off_b1 = off; off += hidden;
  t->off_w2 = off; off += hidden * classes;
  t->off_b2 = off; off += classes;
  t->n_params = off;
  t->trace_capacity = max_epochs > 0 ? max_epochs : 1;
  t->ws_bytes = trainer_ws_bytes(*t);
  Carver measure{nullptr};
  trainer_layout(t, measure);
  cudaError_t ce = cudaGetDevice(&t->slab_device);
  if (ce == cudaSuccess) ce = g_slabs.get(measure.off, t->slab_device, &t->slab, &t->slab_bytes);
  if (ce != cudaSuccess) {
    set_error("trainer: allocation of %zu bytes failed: %s", measure.off, cudaGetErrorString(ce));
    t->slab = nullptr;
    trainer_free(t);
    return MPGNN_ECUDA;
  }
  Carver carve{static_cast<char*>(t->slab)};
  trainer_layout(t, carve);
  // a recycled slab holds another trainer's bytes: parameters, optimiser state and the epoch words start from zero
  ce = cudaMemset(t->params, 0, (size_t)(reinterpret_cast<char*>(t->adam_v + t->n_params) - reinterpret_cast<char*>(t->params)));
  if (ce == cudaSuccess) ce = cudaMemset(t->st, 0, sizeof(TrainerState));
  if (ce != cudaSuccess) {
    set_error("trainer: clearing the parameter block failed: %s", cudaGetErrorString(ce));
    trainer_free(t);
    return MPGNN_ECUDA;
  }
  *out = t;
  return MPGNN_OK;
}

// mean_r(x) of every metapath's first layer: x never changes, so neither does this -- computed once per
// trainer, outside the captured epoch, instead of twice per epoch
static int trainer_prepare(Trainer* t, cudaStream_t s) {
  if (t->h0_ready) return MPGNN_OK;
  for (int i = 0; i < t->n_paths; ++i) {
    const int l = t->path_off[i];
    MPGNN_PROPAGATE(launch_spmm_graph(t->g, t->rel[l], /*transpose=*/0, /*mean=*/1, t->x, t->f_in, t->f_in, nullptr, 0, t->h[l],
                                      t->f_in, s));
  }
  t->h0_ready = true;
  return MPGNN_OK;
}

// one forward pass; train=true applies the seeded dropout and produces glg for the backward
static int trainer_forward(Trainer* t, bool train, cudaStream_t s) {
  const int64_t n = t->n, H = t->hidden, C = t->classes, K1 = t->cat;
  const float* emb = nullptr;
  for (int i = 0; i < t->n_paths; ++i) {
    const float* in = t->x;
    for (int l = t->path_off[i]; l < t->path_off[i + 1]; ++l) {
      const bool first = l == t->path_off[i];
      const int64_t fi = first ? t->f_in : H;
      uint32_t fl = MPGNN_F_RELU | (t->flags & (MPGNN_F_TF32X3 | MPGNN_F_COMPACT_H | MPGNN_F_DENSE_H));
      if (train && t->dropout_p > 0.0) fl |= MPGNN_F_DROPOUT_SEED;
      MPGNN_PROPAGATE(hop_fwd(t->g, t->rel[l], in, fi, t->params + t->off_w[l], t->params + t->off_root[l],
                              t->params + t->off_bias[l], H, fl, t->dropout_p, t->seed, 0, nullptr, t->h[l], t->y[l],
                              train ? reinterpret_cast<uint32_t*>(t->am[l]) : nullptr, t->ws, t->ws_bytes, s,
                              &t->st->drop_off[l], first && t->h0_ready));
      in = t->y[l];
    }
    if (t->n_paths > 1)      // torch.cat(embeddings, 1) (model.py:220): this metapath's columns of the [N, H * n_paths] matrix
      MPGNN_CUDA_CHECK(cudaMemcpy2DAsync(t->emb + (int64_t)i * H, (size_t)K1 * 4, in, (size_t)H * 4, (size_t)H * 4, (size_t)n,
                                         cudaMemcpyDeviceToDevice, s));
    else
      emb = in;
  }
  if (t->n_paths > 1) emb = t->emb;
  GemmRowsArgs a{};
  a.a1 = emb; a.lda1 = K1; a.k1 = K1; a.b = t->params + t->off_w1; a.m = n; a.n = H;
  a.bias = t->params + t->off_b1; a.relu = 1; a.out = t->a1; a.ldo = H;
  // a1 = relu(E W1t + b1); the scratch at the head of t->ws is free between hops (everything is stream ordered)
  const bool head_tc = (t->flags & MPGNN_F_TF32X3) && proj_tcgen05_supported(n, K1, 0, H, t->flags) &&
                       proj_tcgen05_workspace_floats(K1, H) * 4 <= t->ws_bytes;
  if (head_tc) MPGNN_PROPAGATE(launch_proj_tcgen05_ws(a, t->flags, static_cast<float*>(t->ws), s));
  else MPGNN_PROPAGATE(launch_gemm_rows(a, s));
  GemmRowsArgs b{};
  b.a1 = t->a1; b.lda1 = H; b.k1 = H; b.b = t->params + t->off_w2; b.m = n; b.n = C;
  b.bias = t->params + t->off_b2; b.out = t->lg; b.ldo = C;
  MPGNN_PROPAGATE(launch_gemm_rows(b, s));                                   // logits
  if (train)
    return launch_logsoftmax_nll(t->lg, n, C, t->train_idx, t->train_y, t->n_train, t->logp, t->loss_train, t->glg,
                                 t->ws, t->ws_bytes, s);
  return launch_logsoftmax_nll(t->lg, n, C, t->val_idx, t->val_y, t->n_val, t->logp, t->loss_val, nullptr, t->ws,
                               t->ws_bytes, s);
}

static int trainer_backward(Trainer* t, cudaStream_t s) {
  const int64_t n = t->n, H = t->hidden, C = t->classes, K1 = t->cat;
  const float* emb = t->n_paths > 1 ? t->emb : t->y[t->n_layers - 1];
  Workspace ws(t->ws, t->ws_bytes);
  const int64_t pf = trainer_tn_floats(*t);
  float* partials = ws.take<float>(pf);
  MPGNN_REQUIRE(partials != nullptr, MPGNN_EINVAL, "trainer: workspace too small");
  // fc2: [gW2t; gb2] = a1^T glg
  GemmTnArgs t2{};
  t2.a1 = t->a1; t2.lda1 = H; t2.k1 = H; t2.ones_row = 1; t2.b = t->glg; t2.ldb = C; t2.n = C; t2.m = n;
  t2.out1 = t->grads + t->off_w2; t2.ldo1 = C; t2.out_ones = t->grads + t->off_b2;
  t2.partials = partials; t2.partial_capacity_floats = pf;
  MPGNN_PROPAGATE(launch_gemm_tn(t2, s));
  // g_z1 = (glg W2t^T) * [a1 > 0]
  MPGNN_PROPAGATE(launch_pack_b(t->packed, H, t->params + t->off_w2, 1, C, C, H, s));   // B(k=c, n=h) = W2t[h*C + c]
  GemmRowsArgs a{};
  a.a1 = t->glg; a.lda1 = C; a.k1 = C; a.b = t->packed; a.m = n; a.n = H;
  a.gate = t->a1; a.ldgate = H; a.out = t->gz1; a.ldo = H;
  MPGNN_PROPAGATE(launch_gemm_rows(a, s));
  // fc1: [gW1t; gb1] = E^T g_z1   (E = the concatenated embeddings, K1 = H * n_paths wide)
  GemmTnArgs t1{};
  t1.a1 = emb; t1.lda1 = K1; t1.k1 = K1; t1.ones_row = 1; t1.b = t->gz1; t1.ldb = H; t1.n = H; t1.m = n;
  t1.out1 = t->grads + t->off_w1; t1.ldo1 = H; t1.out_ones = t->grads + t->off_b1;
  t1.partials = partials; t1.partial_capacity_floats = pf;
  if ((t->flags & MPGNN_F_TF32X3) && wgrad_tcgen05_supported(n, K1, 0, H, t->flags) &&
      wgrad_tcgen05_workspace_floats(n, H) * 4 <= t->ws_bytes) {
    t1.ones_row = 0;
    MPGNN_PROPAGATE(launch_wgrad_tcgen05(t1, static_cast<float*>(t->ws), s));      // also fills out_ones = colsum(g_z1)
  } else {
    MPGNN_PROPAGATE(launch_gemm_tn(t1, s));
  }
  // g_E = g_z1 W1t^T
  MPGNN_PROPAGATE(launch_pack_b(t->packed, K1, t->params + t->off_w1, 1, H, H, K1, s));   // B(k=o, n=i) = W1t[i*H + o]
  float* ge = t->n_paths > 1 ? t->gemb : t->gxa;
  GemmRowsArgs b{};
  b.a1 = t->gz1; b.lda1 = H; b.k1 = H; b.b = t->packed; b.m = n; b.n = K1; b.out = ge; b.ldo = K1;
  if ((t->flags & MPGNN_F_TF32X3) && proj_tcgen05_supported(n, H, 0, K1, t->flags) &&
      proj_tcgen05_workspace_floats(H, K1) * 4 <= t->ws_bytes)
    MPGNN_PROPAGATE(launch_proj_tcgen05_ws(b, t->flags, static_cast<float*>(t->ws), s));
  else
    MPGNN_PROPAGATE(launch_gemm_rows(b, s));
  for (int i = 0; i < t->n_paths; ++i) {
    float* gy = t->gxa;
    float* gx = t->gxb;
    if (t->n_paths > 1)      // this metapath's columns of g_E, made contiguous
      MPGNN_CUDA_CHECK(cudaMemcpy2DAsync(gy, (size_t)H * 4, t->gemb + (int64_t)i * H, (size_t)K1 * 4, (size_t)H * 4, (size_t)n,
                                         cudaMemcpyDeviceToDevice, s));
    for (int l = t->path_off[i + 1] - 1; l >= t->path_off[i]; --l) {
      const bool first = l == t->path_off[i];
      const int64_t fi = first ? t->f_in : H;
      const float* in = first ? t->x : t->y[l - 1];
      uint32_t fl = MPGNN_F_RELU | (t->flags & (MPGNN_F_TF32X3 | MPGNN_F_COMPACT_H | MPGNN_F_DENSE_H));
      if (t->dropout_p > 0.0) fl |= MPGNN_F_DROPOUT_SEED;
      if (!first) fl |= MPGNN_F_NEED_GX;
      MPGNN_PROPAGATE(hop_bwd(t->g, t->rel[l], in, t->h[l], t->y[l], reinterpret_cast<const uint32_t*>(t->am[l]), gy, fi,
                              t->params + t->off_w[l], t->params + t->off_root[l], H, fl, t->dropout_p, first ? nullptr : gx,
                              t->grads + t->off_w[l], t->grads + t->off_root[l], t->grads + t->off_bias[l], t->ws,
                              t->ws_bytes, s));
      float* tmp = gy; gy = gx; gx = tmp;
    }
  }
  return MPGNN_OK;
}

static int trainer_epoch(Trainer* t, double lr, double b1, double b2, double eps, double wd, bool validate, cudaStream_t s) {
  trainer_bump_kernel<<<1, 32, 0, s>>>(t->st, t->n_layers);
  MPGNN_LAUNCH_CHECK();
  MPGNN_PROPAGATE(trainer_forward(t, true, s));                 // mpgnn_train: forward, nll on train idx
  MPGNN_PROPAGATE(trainer_backward(t, s));                      // backward
  int64_t blocks = ceil_div(t->n_params, 256);
  trainer_adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(t->params, t->grads, t->adam_m, t->adam_v, t->n_params, t->st,
                                                       lr, b1, b2, eps, wd);
  MPGNN_LAUNCH_CHECK();
  if (validate) {
    MPGNN_PROPAGATE(trainer_forward(t, false, s));              // mpgnn_validation: eval forward, nll on val idx
    MPGNN_PROPAGATE(launch_macro_f1(t->logp, t->classes, t->train_idx, t->train_y, t->n_train, t->cm, t->f1_train, s));
    MPGNN_PROPAGATE(launch_macro_f1(t->logp, t->classes, t->val_idx, t->val_y, t->n_val, t->cm, t->f1_val, s));
  }
  trainer_trace_kernel<<<1, 32, 0, s>>>(t->st, t->loss_train, t->loss_val, t->f1_train, t->f1_val, t->trace,
                                        t->trace_capacity, validate ? 1 : 0);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

// mode bit 0: replay the epoch as a CUDA graph; bit 1 (MPGNN_TRAINER_VALIDATE_LAST): run the validation pass only in the
// last epoch of this call.  The validation of the reference (main.py:1084-1100) runs under no_grad in eval mode and
// consumes no random numbers, so skipping it changes nothing but the trace; the returned number is the last epoch's.
int trainer_run(Trainer* t, int64_t epochs, double lr, double b1, double b2, double eps, double wd, int mode,
                cudaStream_t s, double* h_trace, double* h_last_val_f1) {
  MPGNN_REQUIRE(t != nullptr && epochs >= 1, MPGNN_EINVAL, "trainer_run: bad arguments");
  MPGNN_REQUIRE(t->params_set, MPGNN_EINVAL, "trainer_run: call mpgnn_trainer_set_params first");
  const bool use_graph = mode & 1, validate_last = mode & 2;
  cudaStream_t caller = s;
  if (use_graph && (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread)) {
    // stream capture is not permitted on the default streams: run on a private stream, ordered after
    // everything the caller queued so far, and drain it before returning
    if (t->own_stream == nullptr) MPGNN_CUDA_CHECK(cudaStreamCreateWithFlags(&t->own_stream, cudaStreamNonBlocking));
    MPGNN_CUDA_CHECK(cudaStreamSynchronize(caller));
    s = t->own_stream;
  }
  MPGNN_PROPAGATE(trainer_prepare(t, s));
  if (use_graph) {
    const bool stale = lr != t->cap_lr || b1 != t->cap_b1 || b2 != t->cap_b2 || eps != t->cap_eps || wd != t->cap_wd;
    if (stale) {
      if (t->exec) { cudaGraphExecDestroy(t->exec); t->exec = nullptr; }
      if (t->exec_train) { cudaGraphExecDestroy(t->exec_train); t->exec_train = nullptr; }
    }
    auto capture = [&](bool validate, cudaGraphExec_t* exec) -> int {
      if (*exec != nullptr) return MPGNN_OK;
      cudaGraph_t graph = nullptr;
      MPGNN_CUDA_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      int rc = trainer_epoch(t, lr, b1, b2, eps, wd, validate, s);
      cudaError_t ce = cudaStreamEndCapture(s, &graph);
      if (rc != MPGNN_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      MPGNN_CUDA_CHECK(ce);
      ce = cudaGraphInstantiate(exec, graph, 0);
      cudaGraphDestroy(graph);
      MPGNN_CUDA_CHECK(ce);
      return MPGNN_OK;
    };
    MPGNN_PROPAGATE(capture(true, &t->exec));
    if (validate_last && epochs > 1) MPGNN_PROPAGATE(capture(false, &t->exec_train));
    t->cap_lr = lr; t->cap_b1 = b1; t->cap_b2 = b2; t->cap_eps = eps; t->cap_wd = wd;
    for (int64_t e = 0; e < epochs; ++e)
      MPGNN_CUDA_CHECK(cudaGraphLaunch((validate_last && e + 1 < epochs) ? t->exec_train : t->exec, s));
  } else {
    for (int64_t e = 0; e < epochs; ++e)
      MPGNN_PROPAGATE(trainer_epoch(t, lr, b1, b2, eps, wd, !(validate_last && e + 1 < epochs), s));
  }
  if (s != caller) MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
  if (h_trace != nullptr || h_last_val_f1 != nullptr) {
    TrainerState hs;
    MPGNN_CUDA_CHECK(cudaMemcpyAsync(&hs, t->st, sizeof(hs), cudaMemcpyDeviceToHost, s));
    MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
    int64_t done = hs.epoch < t->trace_capacity ? hs.epoch : t->trace_capacity;
    if (h_trace != nullptr && done > 0)
      MPGNN_CUDA_CHECK(cudaMemcpy(h_trace, t->trace, (size_t)done * 4 * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_last_val_f1 != nullptr) MPGNN_CUDA_CHECK(cudaMemcpy(h_last_val_f1, t->f1_val, sizeof(double), cudaMemcpyDeviceToHost));
  }
  return MPGNN_OK;
}

// flat parameter exchange in state_dict order with torch layouts ([out,in] for fc weights);
// internally the fc weights are stored transposed ([in,out]) so the forward needs no repacking
int trainer_set_params(Trainer* t, const float* d_flat, cudaStream_t s) {
  MPGNN_REQUIRE(t && d_flat, MPGNN_EINVAL, "trainer_set_params: NULL argument");
  const int64_t H = t->hidden, C = t->classes, K1 = t->cat;
  MPGNN_CUDA_CHECK(cudaMemcpyAsync(t->params, d_flat, (size_t)t->n_params * 4, cudaMemcpyDeviceToDevice, s));
  MPGNN_PROPAGATE(launch_pack_b(t->params + t->off_w1, H, d_flat + t->off_w1, 1, K1, K1, H, s));   // W1t[i][o] = W1[o][i], W1: [H, K1]
  MPGNN_PROPAGATE(launch_pack_b(t->params + t->off_w2, C, d_flat + t->off_w2, 1, H, H, C, s));   // W2t[h][c] = W2[c][h]
  MPGNN_CUDA_CHECK(cudaMemsetAsync(t->adam_m, 0, (size_t)t->n_params * 4, s));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(t->adam_v, 0, (size_t)t->n_params * 4, s));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(t->st, 0, sizeof(TrainerState), s));
  t->params_set = true;
  return MPGNN_OK;
}

int trainer_get_params(const Trainer* t, float* d_flat, cudaStream_t s) {
  MPGNN_REQUIRE(t && d_flat, MPGNN_EINVAL, "trainer_get_params: NULL argument");
  const int64_t H = t->hidden, C = t->classes, K1 = t->cat;
  MPGNN_CUDA_CHECK(cudaMemcpyAsync(d_flat, t->params, (size_t)t->n_params * 4, cudaMemcpyDeviceToDevice, s));
  MPGNN_PROPAGATE(launch_pack_b(d_flat + t->off_w1, K1, t->params + t->off_w1, 1, H, H, K1, s));   // W1[o][i] = W1t[i][o]
  MPGNN_PROPAGATE(launch_pack_b(d_flat + t->off_w2, H, t->params + t->off_w2, 1, C, C, H, s));   // W2[c][h] = W2t[h][c]
  return MPGNN_OK;
}

int64_t trainer_num_params(const Trainer* t) { return t ? t->n_params : 0; }

// macro-F1 / nll of the CURRENT parameters on an arbitrary index set (mpgnn_test, main.py:1102-1115)
int trainer_evaluate(Trainer* t, const int64_t* d_idx, const int64_t* d_y, int64_t n_idx, cudaStream_t s, float* h_loss,
                     double* h_f1) {
  MPGNN_REQUIRE(t && d_idx && d_y && n_idx > 0, MPGNN_EINVAL, "trainer_evaluate: bad arguments");
  MPGNN_REQUIRE(t->params_set, MPGNN_EINVAL, "trainer_evaluate: call mpgnn_trainer_set_params first");
  MPGNN_PROPAGATE(trainer_prepare(t, s));
  const int64_t *vi = t->val_idx, *vy = t->val_y, vn = t->n_val;
  t->val_idx = d_idx; t->val_y = d_y; t->n_val = n_idx;
  int rc = trainer_forward(t, false, s);
  if (rc == MPGNN_OK) rc = launch_macro_f1(t->logp, t->classes, d_idx, d_y, n_idx, t->cm, t->f1_val, s);
  t->val_idx = vi; t->val_y = vy; t->n_val = vn;
  MPGNN_PROPAGATE(rc);
  MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
  if (h_loss) MPGNN_CUDA_CHECK(cudaMemcpy(h_loss, t->loss_val, sizeof(float), cudaMemcpyDeviceToHost));
  if (h_f1) MPGNN_CUDA_CHECK(cudaMemcpy(h_f1, t->f1_val, sizeof(double), cudaMemcpyDeviceToHost));
  return MPGNN_OK;
}

}  // namespace mpgnn
