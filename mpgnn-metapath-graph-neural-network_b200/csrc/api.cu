// extern "C" surface declared in include/mpgnn_b200.h.
#include <string.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace mpgnn {

static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// ---- launch counter + optional event timing ------------------------------------------------
static std::atomic<long long> g_launches{0};   // candidate trainers launch from several host threads
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

constexpr int kMaxTimerSlots = 32;
constexpr int kMaxTimerEvents = 4096;
static bool g_timing = false;
static int g_n_slots = 0;
static char g_slot_names[kMaxTimerSlots][48];
static double g_slot_ms[kMaxTimerSlots];
static long long g_slot_calls[kMaxTimerSlots];
static cudaEvent_t g_ev_start[kMaxTimerEvents], g_ev_stop[kMaxTimerEvents];
static int g_ev_slot[kMaxTimerEvents];
static int g_n_events = 0, g_n_events_created = 0;
static std::mutex g_timer_mu;   // candidate trainers record from several host threads

ScopedTimer::ScopedTimer(const char* name, cudaStream_t s) : slot(-1), stream(s) {
  if (!g_timing) return;
  // events recorded while a stream is being captured belong to the graph and cannot be queried: no timing there
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) { (void)cudaGetLastError(); return; }
  if (cap != cudaStreamCaptureStatusNone) return;
  std::lock_guard<std::mutex> lk(g_timer_mu);
  if (g_n_events >= kMaxTimerEvents) return;
  int k = 0;
  for (; k < g_n_slots; ++k)
    if (strncmp(g_slot_names[k], name, 47) == 0) break;
  if (k == g_n_slots) {
    if (g_n_slots >= kMaxTimerSlots) return;
    strncpy(g_slot_names[k], name, 47);
    g_slot_names[k][47] = 0;
    g_slot_ms[k] = 0.0;
    g_slot_calls[k] = 0;
    ++g_n_slots;
  }
  if (g_n_events >= g_n_events_created) {
    cudaEventCreate(&g_ev_start[g_n_events]);
    cudaEventCreate(&g_ev_stop[g_n_events]);
    ++g_n_events_created;
  }
  slot = g_n_events++;
  g_ev_slot[slot] = k;
  cudaEventRecord(g_ev_start[slot], stream);
}

ScopedTimer::~ScopedTimer() {
  if (slot >= 0) cudaEventRecord(g_ev_stop[slot], stream);
}

static void timing_flush() {
  std::lock_guard<std::mutex> lk(g_timer_mu);
  for (int i = 0; i < g_n_events; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(g_ev_stop[i]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, g_ev_start[i], g_ev_stop[i]) == cudaSuccess) {
      g_slot_ms[g_ev_slot[i]] += ms;
      g_slot_calls[g_ev_slot[i]] += 1;
    }
  }
  g_n_events = 0;
  (void)cudaGetLastError();   // a failed query must not surface as the next launch's error
}

int graph_build_device(const int64_t* d_edge_index, const int64_t* d_edge_type, int64_t e, int64_t n, int64_t r,
                       cudaStream_t s, mpgnn_graph_impl** out);
void graph_free(mpgnn_graph_impl* g);
int64_t hop_workspace_bytes(int64_t n, int64_t f_in, int64_t f_out);
void set_tc_cta_cap(int cap);
int hop_h_compact(const mpgnn_graph_impl* g, int64_t rel, int64_t f_in, int64_t f_out, uint32_t flags);
int hop_fwd(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t f_in, const float* w, const float* root,
            const float* bias, int64_t f_out, uint32_t flags, double p, uint64_t seed, uint64_t offset,
            const uint8_t* mask_bits, float* h, float* y, uint32_t* actmask, void* ws_ptr, int64_t ws_bytes,
            cudaStream_t s, const uint64_t* offset_ptr, bool h_precomputed);
int hop_bwd(const mpgnn_graph_impl* g, int64_t rel, const float* x, const float* h, const float* y,
            const uint32_t* actmask, const float* gy, int64_t f_in, const float* w, const float* root, int64_t f_out,
            uint32_t flags, double p, float* gx, float* gw, float* groot, float* gbias, void* ws_ptr, int64_t ws_bytes,
            cudaStream_t s);
int launch_logsoftmax_nll(const float* logits, int64_t n, int64_t c, const int64_t* idx, const int64_t* y,
                          int64_t n_idx, float* logp, float* loss, float* glogits, void* ws, int64_t ws_bytes,
                          cudaStream_t s);
int launch_macro_f1(const float* logp, int64_t c, const int64_t* idx, const int64_t* y, int64_t n_idx, int32_t* cm,
                    double* f1, cudaStream_t s);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, int64_t step, double lr, double beta1,
                double beta2, double eps, double wd, cudaStream_t s);

int64_t score_workspace_bytes(int64_t n);
int score_relation(const mpgnn_graph_impl* g, int64_t rel, float* w, const float* labels, const uint8_t* src_mask,
                   int64_t epochs, double lr,
                   float* m, float* v, float* loss_traj, int32_t* argmax_dst, void* ws_ptr, int64_t ws_bytes,
                   cudaStream_t s);

int64_t score_bags_workspace_bytes(int64_t n, int64_t n_bags, int64_t feat);
int score_bags(const mpgnn_graph_impl* g, int64_t rel, const int32_t* bag_ptr, const int32_t* bag_src, int64_t n_bags,
               const float* bag_labels, const float* x, int64_t feat, float* w, float* lin, const uint8_t* grad_mask,
               int use_mask, int64_t epochs, double lr, float* loss_traj, int32_t* best_dst, int32_t* best_src,
               float* diff, float* src_val, void* ws_ptr, int64_t ws_bytes, cudaStream_t s);
struct Trainer;
int trainer_create(const mpgnn_graph_impl* g, const float* x, int64_t f_in, int64_t hidden, int64_t classes,
                   const int64_t* h_rel, const int64_t* h_path_ptr, int64_t n_paths, const int64_t* train_idx,
                   const int64_t* train_y, int64_t n_train, const int64_t* val_idx, const int64_t* val_y, int64_t n_val,
                   double dropout_p, uint64_t seed, uint32_t flags, int64_t max_epochs, Trainer** out);
void trainer_free(Trainer* t);
int64_t trainer_num_params(const Trainer* t);
int trainer_set_params(Trainer* t, const float* d_flat, cudaStream_t s);
int trainer_get_params(const Trainer* t, float* d_flat, cudaStream_t s);
int trainer_run(Trainer* t, int64_t epochs, double lr, double b1, double b2, double eps, double wd, int use_graph,
                cudaStream_t s, double* h_trace, double* h_last_val_f1);
int trainer_evaluate(Trainer* t, const int64_t* d_idx, const int64_t* d_y, int64_t n_idx, cudaStream_t s, float* h_loss,
                     double* h_f1);

}  // namespace mpgnn

using namespace mpgnn;

static inline const mpgnn_graph_impl* impl(const mpgnn_graph* g) {
  return reinterpret_cast<const mpgnn_graph_impl*>(g);
}
static inline cudaStream_t stream_of(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

const char* mpgnn_last_error(void) { return g_error; }

int mpgnn_abi_version(void) { return 1; }

void mpgnn_set_tc_cta_cap(int max_ctas) { set_tc_cta_cap(max_ctas); }

long long mpgnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void mpgnn_timing_enable(int on) {
  timing_flush();
  g_timing = on != 0;
}

void mpgnn_timing_reset(void) {
  timing_flush();
  g_n_slots = 0;
}

int mpgnn_timing_collect(char* names, int64_t names_bytes, double* ms, int64_t* calls, int64_t capacity) {
  timing_flush();
  int64_t used = 0;
  int n = 0;
  for (int k = 0; k < g_n_slots && k < capacity; ++k) {
    const int64_t len = (int64_t)strlen(g_slot_names[k]);
    if (used + len + 2 > names_bytes) break;
    memcpy(names + used, g_slot_names[k], (size_t)len);
    used += len;
    names[used++] = ';';
    ms[k] = g_slot_ms[k];
    calls[k] = g_slot_calls[k];
    ++n;
  }
  if (names_bytes > 0) names[used < names_bytes ? used : names_bytes - 1] = 0;
  return n;
}

int mpgnn_graph_build(const int64_t* d_edge_index, const int64_t* d_edge_type, int64_t num_edges, int64_t num_nodes,
                      int64_t num_relations, void* stream, mpgnn_graph** out) {
  mpgnn_graph_impl* g = nullptr;
  int rc = graph_build_device(d_edge_index, d_edge_type, num_edges, num_nodes, num_relations, stream_of(stream), &g);
  if (rc == MPGNN_OK) *out = reinterpret_cast<mpgnn_graph*>(g);
  return rc;
}

int mpgnn_graph_build_host(const int64_t* h_edge_index, const int64_t* h_edge_type, int64_t num_edges,
                           int64_t num_nodes, int64_t num_relations, void* stream, mpgnn_graph** out) {
  MPGNN_REQUIRE(num_edges >= 0 && (num_edges == 0 || (h_edge_index && h_edge_type)), MPGNN_EINVAL,
                "graph_build_host: bad edge arrays");
  cudaStream_t s = stream_of(stream);
  int64_t *d_ei = nullptr, *d_et = nullptr;
  const int64_t e1 = num_edges > 0 ? num_edges : 1;
  MPGNN_CUDA_CHECK(cudaMalloc(&d_ei, (size_t)e1 * 16));
  cudaError_t ce = cudaMalloc(&d_et, (size_t)e1 * 8);
  if (ce != cudaSuccess) {
    cudaFree(d_ei);
    set_error("graph_build_host: cudaMalloc failed: %s", cudaGetErrorString(ce));
    return MPGNN_ECUDA;
  }
  int rc = MPGNN_OK;
  if (num_edges > 0) {
    ce = cudaMemcpyAsync(d_ei, h_edge_index, (size_t)num_edges * 16, cudaMemcpyHostToDevice, s);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_et, h_edge_type, (size_t)num_edges * 8, cudaMemcpyHostToDevice, s);
    if (ce != cudaSuccess) {
      set_error("graph_build_host: H2D copy failed: %s", cudaGetErrorString(ce));
      rc = MPGNN_ECUDA;
    }
  }
  if (rc == MPGNN_OK) rc = mpgnn_graph_build(d_ei, d_et, num_edges, num_nodes, num_relations, stream, out);
  cudaStreamSynchronize(s);
  cudaFree(d_ei);
  cudaFree(d_et);
  return rc;
}

void mpgnn_graph_free(mpgnn_graph* g) { graph_free(reinterpret_cast<mpgnn_graph_impl*>(g)); }

int mpgnn_graph_info(const mpgnn_graph* g, int64_t* num_nodes, int64_t* num_edges, int64_t* num_relations) {
  MPGNN_REQUIRE(g != nullptr, MPGNN_EINVAL, "graph_info: NULL graph");
  if (num_nodes) *num_nodes = impl(g)->n;
  if (num_edges) *num_edges = impl(g)->e;
  if (num_relations) *num_relations = impl(g)->r;
  return MPGNN_OK;
}

int mpgnn_graph_relation_view(const mpgnn_graph* g, int64_t relation, int transpose, const int32_t** d_ptr,
                              const int32_t** d_idx, const int32_t** d_eid, int64_t* num_rel_edges) {
  MPGNN_REQUIRE(g != nullptr, MPGNN_EINVAL, "relation_view: NULL graph");
  const mpgnn_graph_impl* gi = impl(g);
  MPGNN_REQUIRE(relation >= 0 && relation < gi->r, MPGNN_ERANGE, "relation_view: relation %lld outside [0,%lld)",
                (long long)relation, (long long)gi->r);
  if (d_ptr) *d_ptr = (transpose ? gi->csc_ptr : gi->csr_ptr) + relation * gi->n;
  if (d_idx) *d_idx = transpose ? gi->csc_idx : gi->csr_idx;
  if (d_eid) *d_eid = transpose ? gi->csc_eid : gi->csr_eid;
  if (num_rel_edges) *num_rel_edges = gi->rel_offsets_host[relation + 1] - gi->rel_offsets_host[relation];
  return MPGNN_OK;
}

int mpgnn_graph_relation_counts(const mpgnn_graph* g, int64_t* h_counts) {
  MPGNN_REQUIRE(g != nullptr && h_counts != nullptr, MPGNN_EINVAL, "relation_counts: NULL argument");
  const mpgnn_graph_impl* gi = impl(g);
  for (int64_t k = 0; k < gi->r; ++k) h_counts[k] = gi->rel_offsets_host[k + 1] - gi->rel_offsets_host[k];
  return MPGNN_OK;
}

int mpgnn_spmm(const mpgnn_graph* g, int64_t relation, int transpose, int mean, const float* d_x, int64_t ldx,
               int64_t feat, const float* d_init, int64_t ldinit, float* d_out, int64_t ldout, void* stream) {
  MPGNN_REQUIRE(g && d_x && d_out, MPGNN_EINVAL, "spmm: NULL argument");
  const mpgnn_graph_impl* gi = impl(g);
  MPGNN_REQUIRE(relation >= 0 && relation < gi->r, MPGNN_ERANGE, "spmm: relation %lld outside [0,%lld)",
                (long long)relation, (long long)gi->r);
  MPGNN_REQUIRE(feat >= 1 && ldx >= feat && ldout >= feat && (!d_init || ldinit >= feat), MPGNN_EINVAL,
                "spmm: bad strides");
  return launch_spmm_graph(gi, relation, transpose, mean, d_x, ldx, feat, d_init, ldinit, d_out, ldout, stream_of(stream));
}

int mpgnn_scale_rows_by_degree(const mpgnn_graph* g, int64_t relation, const float* d_x, int64_t ldx, int64_t feat,
                               float* d_out, int64_t ldout, void* stream) {
  MPGNN_REQUIRE(g && d_x && d_out, MPGNN_EINVAL, "scale_rows_by_degree: NULL argument");
  const mpgnn_graph_impl* gi = impl(g);
  MPGNN_REQUIRE(relation >= 0 && relation < gi->r, MPGNN_ERANGE, "scale_rows_by_degree: relation %lld outside [0,%lld)",
                (long long)relation, (long long)gi->r);
  MPGNN_REQUIRE(feat >= 1 && ldx >= feat && ldout >= feat, MPGNN_EINVAL, "scale_rows_by_degree: bad strides");
  return launch_scale_rows_by_degree(gi, relation, d_x, ldx, feat, d_out, ldout, stream_of(stream));
}

int mpgnn_hop_fwd(const mpgnn_graph* g, int64_t relation, const float* d_x, int64_t f_in, const float* d_w,
                  const float* d_root, const float* d_bias, int64_t f_out, uint32_t flags, double dropout_p,
                  uint64_t seed, uint64_t offset, const uint8_t* d_mask_bits, float* d_h, float* d_y,
                  uint32_t* d_actmask, void* d_workspace, int64_t workspace_bytes, void* stream) {
  return hop_fwd(impl(g), relation, d_x, f_in, d_w, d_root, d_bias, f_out, flags, dropout_p, seed, offset,
                 d_mask_bits, d_h, d_y, d_actmask, d_workspace, workspace_bytes, stream_of(stream), nullptr, false);
}

int mpgnn_hop_bwd(const mpgnn_graph* g, int64_t relation, const float* d_x, const float* d_h, const float* d_y,
                  const uint32_t* d_actmask, const float* d_gy, int64_t f_in, const float* d_w, const float* d_root,
                  int64_t f_out, uint32_t flags, double dropout_p, float* d_gx, float* d_gw, float* d_groot,
                  float* d_gbias, void* d_workspace, int64_t workspace_bytes, void* stream) {
  return hop_bwd(impl(g), relation, d_x, d_h, d_y, d_actmask, d_gy, f_in, d_w, d_root, f_out, flags, dropout_p, d_gx, d_gw,
                 d_groot, d_gbias, d_workspace, workspace_bytes, stream_of(stream));
}

int64_t mpgnn_hop_h_rows(const mpgnn_graph* g, int64_t relation, int64_t f_in, int64_t f_out, uint32_t flags) {
  const mpgnn_graph_impl* gi = impl(g);
  MPGNN_REQUIRE(gi != nullptr, MPGNN_EINVAL, "hop_h_rows: NULL graph");
  MPGNN_REQUIRE(relation >= 0 && relation < gi->r, MPGNN_ERANGE, "hop_h_rows: relation %lld outside [0,%lld)",
                (long long)relation, (long long)gi->r);
  return hop_h_compact(gi, relation, f_in, f_out, flags) ? graph_rel_nnz_rows(gi, relation) : gi->n;
}

int mpgnn_graph_relation_rows(const mpgnn_graph* g, int64_t relation, const int32_t** d_rows, int64_t* count) {
  const mpgnn_graph_impl* gi = impl(g);
  MPGNN_REQUIRE(gi && d_rows && count, MPGNN_EINVAL, "graph_relation_rows: NULL argument");
  MPGNN_REQUIRE(relation >= 0 && relation < gi->r, MPGNN_ERANGE, "graph_relation_rows: relation %lld outside [0,%lld)",
                (long long)relation, (long long)gi->r);
  *d_rows = gi->nz_rows + gi->rel_nz_host[relation];
  *count = graph_rel_nnz_rows(gi, relation);
  return MPGNN_OK;
}

int64_t mpgnn_hop_workspace_bytes(int64_t num_nodes, int64_t f_in, int64_t f_out) {
  return hop_workspace_bytes(num_nodes, f_in, f_out);
}

int64_t mpgnn_gemm_workspace_bytes(int64_t m, int64_t k, int64_t n) {
  return (align_up(k * n, 64) + align_up(gemm_tn_partial_floats(m, k + 1, n), 64)) * 4 + 4 * 256;
}

int mpgnn_gemm_rows(const float* d_a, int64_t lda, int64_t m, int64_t k, const float* d_b, int64_t ldb_k,
                    int64_t ldb_n, int64_t n, const float* d_bias, int relu, const float* d_gate, int64_t ldgate,
                    float* d_out, int64_t ldo, void* d_workspace, int64_t workspace_bytes, void* stream) {
  MPGNN_REQUIRE(d_a && d_b && d_out && m >= 0 && k >= 1 && n >= 1, MPGNN_EINVAL, "gemm_rows: bad arguments");
  cudaStream_t s = stream_of(stream);
  const float* b = d_b;
  if (!(ldb_n == 1 && ldb_k == n)) {
    Workspace ws(d_workspace, workspace_bytes);
    float* bp = ws.take<float>(align_up(k * n, 64));
    MPGNN_REQUIRE(bp != nullptr, MPGNN_EINVAL, "gemm_rows: workspace too small");
    MPGNN_PROPAGATE(launch_pack_b(bp, n, d_b, ldb_k, ldb_n, k, n, s));
    b = bp;
  }
  GemmRowsArgs a{};
  a.a1 = d_a; a.lda1 = lda; a.k1 = k;
  a.b = b; a.m = m; a.n = n;
  a.bias = d_bias; a.relu = relu;
  a.gate = d_gate; a.ldgate = ldgate;
  a.out = d_out; a.ldo = ldo;
  return launch_gemm_rows(a, s);
}

int mpgnn_gemm_tn(const float* d_a, int64_t lda, int64_t m, int64_t k, const float* d_b, int64_t ldb, int64_t n,
                  float* d_out, int64_t ldo, float* d_colsum, void* d_workspace, int64_t workspace_bytes,
                  void* stream) {
  MPGNN_REQUIRE(d_a && d_b && d_out && m >= 0 && k >= 1 && n >= 1, MPGNN_EINVAL, "gemm_tn: bad arguments");
  Workspace ws(d_workspace, workspace_bytes);
  (void)ws.take<float>(align_up(k * n, 64));
  const int64_t pf = gemm_tn_partial_floats(m, k + 1, n);
  float* partials = ws.take<float>(align_up(pf, 64));
  MPGNN_REQUIRE(partials != nullptr, MPGNN_EINVAL, "gemm_tn: workspace too small");
  GemmTnArgs t{};
  t.a1 = d_a; t.lda1 = lda; t.k1 = k;
  t.ones_row = 1;
  t.b = d_b; t.ldb = ldb; t.n = n; t.m = m;
  t.out1 = d_out; t.ldo1 = ldo;
  t.out_ones = d_colsum;
  t.partials = partials; t.partial_capacity_floats = pf;
  return launch_gemm_tn(t, stream_of(stream));
}

int mpgnn_logsoftmax_nll(const float* d_logits, int64_t num_nodes, int64_t num_classes, const int64_t* d_idx,
                         const int64_t* d_y, int64_t n_idx, float* d_logp, float* d_loss, float* d_glogits,
                         void* d_workspace, int64_t workspace_bytes, void* stream) {
  return launch_logsoftmax_nll(d_logits, num_nodes, num_classes, d_idx, d_y, n_idx, d_logp, d_loss, d_glogits,
                               d_workspace, workspace_bytes, stream_of(stream));
}

int mpgnn_macro_f1(const float* d_logp, int64_t num_classes, const int64_t* d_idx, const int64_t* d_y, int64_t n_idx,
                   int32_t* d_confusion, double* d_f1, void* stream) {
  return launch_macro_f1(d_logp, num_classes, d_idx, d_y, n_idx, d_confusion, d_f1, stream_of(stream));
}

int mpgnn_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                    int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                    void* stream) {
  return launch_adam(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, step, lr, beta1, beta2, eps, weight_decay,
                     stream_of(stream));
}

int64_t mpgnn_score_workspace_bytes(int64_t num_nodes) { return score_workspace_bytes(num_nodes); }

int mpgnn_score_relation(const mpgnn_graph* g, int64_t relation, float* d_w, const float* d_labels,
                         const uint8_t* d_source_mask, int64_t epochs, double lr, float* d_m, float* d_v, float* d_loss_traj, int32_t* d_argmax_dst,
                         void* d_workspace, int64_t workspace_bytes, void* stream) {
  return score_relation(impl(g), relation, d_w, d_labels, d_source_mask, epochs, lr, d_m, d_v, d_loss_traj, d_argmax_dst, d_workspace,
                        workspace_bytes, stream_of(stream));
}

int64_t mpgnn_score_bags_workspace_bytes(int64_t num_nodes, int64_t num_bags, int64_t feat) {
  return score_bags_workspace_bytes(num_nodes, num_bags, feat);
}

int mpgnn_score_bags(const mpgnn_graph* g, int64_t relation, const int32_t* d_bag_ptr, const int32_t* d_bag_src,
                     int64_t num_bags, const float* d_bag_labels, const float* d_x, int64_t feat, float* d_w,
                     float* d_lin, const uint8_t* d_grad_mask, int use_mask, int64_t epochs, double lr,
                     float* d_loss_traj, int32_t* d_best_dst, int32_t* d_best_src, float* d_diff, float* d_src_val,
                     void* d_workspace, int64_t workspace_bytes, void* stream) {
  return score_bags(impl(g), relation, d_bag_ptr, d_bag_src, num_bags, d_bag_labels, d_x, feat, d_w, d_lin,
                    d_grad_mask, use_mask, epochs, lr, d_loss_traj, d_best_dst, d_best_src, d_diff, d_src_val,
                    d_workspace, workspace_bytes, stream_of(stream));
}

int mpgnn_trainer_create(const mpgnn_graph* g, const float* d_x, int64_t f_in, int64_t hidden, int64_t num_classes,
                         const int64_t* h_relations, int64_t n_layers, const int64_t* d_train_idx,
                         const int64_t* d_train_y, int64_t n_train, const int64_t* d_val_idx, const int64_t* d_val_y,
                         int64_t n_val, double dropout_p, uint64_t seed, uint32_t flags, int64_t max_epochs,
                         mpgnn_trainer** out) {
  const int64_t path_ptr[2] = {0, n_layers};
  return mpgnn_trainer_create_multi(g, d_x, f_in, hidden, num_classes, h_relations, path_ptr, 1, d_train_idx, d_train_y,
                                    n_train, d_val_idx, d_val_y, n_val, dropout_p, seed, flags, max_epochs, out);
}
int mpgnn_trainer_create_multi(const mpgnn_graph* g, const float* d_x, int64_t f_in, int64_t hidden, int64_t num_classes,
                               const int64_t* h_relations, const int64_t* h_path_ptr, int64_t n_paths,
                               const int64_t* d_train_idx, const int64_t* d_train_y, int64_t n_train,
                               const int64_t* d_val_idx, const int64_t* d_val_y, int64_t n_val, double dropout_p,
                               uint64_t seed, uint32_t flags, int64_t max_epochs, mpgnn_trainer** out) {
  Trainer* t = nullptr;
  int rc = trainer_create(impl(g), d_x, f_in, hidden, num_classes, h_relations, h_path_ptr, n_paths, d_train_idx,
                          d_train_y, n_train, d_val_idx, d_val_y, n_val, dropout_p, seed, flags, max_epochs, &t);
  if (rc == MPGNN_OK) *out = reinterpret_cast<mpgnn_trainer*>(t);
  return rc;
}
void mpgnn_trainer_free(mpgnn_trainer* t) { trainer_free(reinterpret_cast<Trainer*>(t)); }
int64_t mpgnn_trainer_num_params(const mpgnn_trainer* t) { return trainer_num_params(reinterpret_cast<const Trainer*>(t)); }
int mpgnn_trainer_set_params(mpgnn_trainer* t, const float* d_flat, void* stream) {
  return trainer_set_params(reinterpret_cast<Trainer*>(t), d_flat, stream_of(stream));
}
int mpgnn_trainer_get_params(const mpgnn_trainer* t, float* d_flat, void* stream) {
  return trainer_get_params(reinterpret_cast<const Trainer*>(t), d_flat, stream_of(stream));
}
int mpgnn_trainer_run(mpgnn_trainer* t, int64_t epochs, double lr, double beta1, double beta2, double eps,
                      double weight_decay, int mode, void* stream, double* h_trace, double* h_last_val_f1) {
  return trainer_run(reinterpret_cast<Trainer*>(t), epochs, lr, beta1, beta2, eps, weight_decay, mode,
                     stream_of(stream), h_trace, h_last_val_f1);
}
int mpgnn_trainer_evaluate(mpgnn_trainer* t, const int64_t* d_idx, const int64_t* d_y, int64_t n_idx, void* stream,
                           float* h_loss, double* h_f1) {
  return trainer_evaluate(reinterpret_cast<Trainer*>(t), d_idx, d_y, n_idx, stream_of(stream), h_loss, h_f1);
}

}  // extern "C"
