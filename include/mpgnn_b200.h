/*
 * mpgnn_b200 -- C ABI of the B200-native MPGNN hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch types),
 * runs on the CUDA stream passed as `stream` (a cudaStream_t cast to void*), performs no
 * hidden device synchronisation unless stated, and returns 0 on success or a negative
 * MPGNN_E* code; the message is available from mpgnn_last_error() (thread local).
 * Pointers named d_* are DEVICE pointers, h_* are HOST pointers.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference repository root).  Orientation: edge_index row 0 is the aggregation TARGET
 * (CSR row), row 1 the message SOURCE (CSR column) -- every reference conv is built
 * with flow='target_to_source' (model.py:190,192).
 */
#ifndef MPGNN_B200_H
#define MPGNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPGNN_OK 0
#define MPGNN_EINVAL (-1)   /* bad argument (maps to ValueError in the Python host)      */
#define MPGNN_ECUDA (-2)    /* CUDA runtime error (RuntimeError)                          */
#define MPGNN_ERANGE (-3)   /* node id / relation id outside [0,N) / [0,R) (ValueError)   */
#define MPGNN_ENOTSUP (-4)  /* size outside what the kernels support (NotImplementedError) */

/* epilogue flags of mpgnn_hop_fwd / mpgnn_hop_bwd */
#define MPGNN_F_RELU 1u         /* y = relu(z)                 (model.py:210,213)            */
#define MPGNN_F_DROPOUT_SEED 2u /* keep-mask from the counter RNG (seed, offset)             */
#define MPGNN_F_DROPOUT_MASK 4u /* keep-mask supplied bit-packed, np.packbits(axis=1) layout */
#define MPGNN_F_NEED_GX 8u      /* bwd: also produce the input gradient (hidden layers)      */
#define MPGNN_F_TF32X3 16u      /* projection on tcgen05 with the 3xTF32 split (fp32 parity) */
#define MPGNN_F_BF16 32u        /* reserved: not built, every entry point rejects it (ENOTSUP) */
#define MPGNN_F_COMPACT_H 64u   /* force the compact hop (d_h = one row per non-empty bucket) where the shapes allow */
#define MPGNN_F_DENSE_H 128u    /* never use the compact hop (d_h = N rows)                                            */

typedef struct mpgnn_graph mpgnn_graph; /* relation-typed CSR + CSC, device resident */

const char* mpgnn_last_error(void);
int mpgnn_abi_version(void);

/* ---- measurement hooks (bench.py) ------------------------------------------------------
 * mpgnn_launch_count: kernels launched by this library so far in this process.
 * mpgnn_timing_*: optional CUDA-event timing of each kernel class on its launching stream;
 * collect() synchronises on the recorded events and returns the number of classes, their
 * ';'-separated names, accumulated milliseconds and call counts. */
long long mpgnn_launch_count(void);
/* Upper bound on the CTAs (= SMs) a tensor-core projection launched FROM THE CALLING THREAD may take; 0 = all.  For
 * callers that run several independent trainings side by side on their own streams (the candidate fan-out of
 * main.py:1444-1462 on one GPU): persistent one-CTA-per-SM kernels of different streams otherwise serialise.  Results
 * do not depend on it; CUDA graphs captured while it is set keep the grid they were captured with. */
void mpgnn_set_tc_cta_cap(int max_ctas);
void mpgnn_timing_enable(int on);
void mpgnn_timing_reset(void);
int mpgnn_timing_collect(char* names, int64_t names_bytes, double* ms, int64_t* calls, int64_t capacity);

/* ---- K1: relation-typed CSR/CSC construction ------------------------------------------
 * Replaces the per-call O(E) filter `masked_edge_index(edge_index, edge_type == relation)`
 * (mp_rgcn_layer.py:29-37, call site :231) and the scatter index PyG derives from it:
 * edges are bucketed ONCE by (relation,row) and by (relation,col) with a stable LSD radix
 * sort, so bucket (r,i) lists its edges in the original edge order, duplicates kept --
 * bit-exact with what the reference enumerates.  Synchronises the stream once at the end
 * (range check).  d_edge_index is int64 [2,E] row-major, d_edge_type int64 [E]. */
int mpgnn_graph_build(const int64_t* d_edge_index, const int64_t* d_edge_type, int64_t num_edges,
                      int64_t num_nodes, int64_t num_relations, void* stream, mpgnn_graph** out);
/* Same with HOST buffers (the reference-facing form: main.py:366-372 builds them on the
 * host); the host->device copies are part of the call. */
int mpgnn_graph_build_host(const int64_t* h_edge_index, const int64_t* h_edge_type, int64_t num_edges,
                           int64_t num_nodes, int64_t num_relations, void* stream, mpgnn_graph** out);
void mpgnn_graph_free(mpgnn_graph* g);
int mpgnn_graph_info(const mpgnn_graph* g, int64_t* num_nodes, int64_t* num_edges, int64_t* num_relations);
/* Device views of one relation: ptr has N+1 entries indexing the GLOBAL arrays idx/eid
 * (so idx[ptr[i]..ptr[i+1]) are the neighbours of node i); transpose=0 -> buckets by row
 * (idx = message sources), transpose=1 -> buckets by col (idx = targets).  eid = original
 * edge id.  *num_rel_edges = E_r. */
int mpgnn_graph_relation_view(const mpgnn_graph* g, int64_t relation, int transpose, const int32_t** d_ptr,
                              const int32_t** d_idx, const int32_t** d_eid, int64_t* num_rel_edges);
/* Host copy of the number of edges of every relation (int64 [R]). */
int mpgnn_graph_relation_counts(const mpgnn_graph* g, int64_t* h_counts);

/* ---- K2: per-hop aggregation ----------------------------------------------------------
 * PyG 2.3.1 MessagePassing.propagate(aggr='mean', flow='target_to_source') as called at
 * mp_rgcn_layer.py:236: h[i,:] = sum_{e in E_r,row(e)=i} x[col(e),:] / max(1,deg_r(i)),
 * fp32 sum in edge order.  transpose=1 gives the backward's un-normalised transpose
 * gather  out[j,:] = init[j,:] + sum_{e in E_r,col(e)=j} x[row(e),:].  x/out row strides
 * in floats; d_init may be NULL (zeros) or alias d_out.  With d_init == d_out (same stride)
 * the call accumulates in place and neither reads nor writes the rows that have no edge of
 * the relation -- the result is bit for bit the out-of-place one. */
int mpgnn_spmm(const mpgnn_graph* g, int64_t relation, int transpose, int mean, const float* d_x, int64_t ldx,
               int64_t feat, const float* d_init, int64_t ldinit, float* d_out, int64_t ldout, void* stream);

/* d_out[i,:] = d_x[i,:] / max(1, deg_r(i)) -- the mean's normalisation on its own (d_out may alias d_x).  With
 * mpgnn_spmm(transpose=1, mean=0) it gives the backward of the mean aggregation, g_x += A_r^T (D_r^-1 g_h), for callers
 * that compose the aggregation themselves: the all-relation RGCN baseline (model.py:132-151 `Net`, main_rgcn.py:452-472;
 * PyG 2.3.1 RGCNConv: out = sum_r mean_r(x) W_r + x root + bias). */
int mpgnn_scale_rows_by_degree(const mpgnn_graph* g, int64_t relation, const float* d_x, int64_t ldx, int64_t feat,
                               float* d_out, int64_t ldout, void* stream);

/* ---- K2+K3: one metapath hop, forward --------------------------------------------------
 * CustomRGCNConv.forward(layer_num, relation, x, edge_index, edge_type)
 * (mp_rgcn_layer.py:158-271) fused with the relu + dropout MPNetm applies to it
 * (model.py:210-214):   y = drop(relu( mean_r(x) @ W + x @ root + bias )).
 * d_h (N x f_in) receives the aggregated features kept for the backward.
 * d_bias may be NULL.  dropout: p in [0,1); MPGNN_F_DROPOUT_MASK reads d_mask_bits
 * ([N, ceil(f_out/8)] bytes, MSB first); MPGNN_F_DROPOUT_SEED draws from (seed, offset).
 * d_actmask (may be NULL; needs f_out % 32 == 0) receives the activation bitmask of y,
 * [N, f_out/32] 32-bit words with bit j of word c = [y(row, 32c+j) > 0]: all the backward
 * needs of y, 1/32 of its size, and written by the projection epilogue for free. */
int mpgnn_hop_fwd(const mpgnn_graph* g, int64_t relation, const float* d_x, int64_t f_in, const float* d_w,
                  const float* d_root, const float* d_bias, int64_t f_out, uint32_t flags, double dropout_p,
                  uint64_t seed, uint64_t offset, const uint8_t* d_mask_bits, float* d_h, float* d_y,
                  uint32_t* d_actmask, void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- K4: one metapath hop, backward ----------------------------------------------------
 * What autograd derives for the call above:
 *   g_z = g_y * [y>0] (* 1/(1-p) under dropout);  g_bias = colsum g_z;
 *   g_W = h^T g_z;  g_root = x^T g_z;  t = (g_z W^T)/deg;
 *   g_x[j] = (g_z root^T)[j] + sum_{e in E_r, col(e)=j} t[row(e)]   (only with NEED_GX).
 * Reductions over the N rows use a fixed split and a fixed summation order
 * (deterministic, run-to-run and across GPU counts).  Gradients are WRITTEN, not
 * accumulated.  d_gx may be NULL without MPGNN_F_NEED_GX.  [y>0] is taken from d_actmask
 * (the bitmask mpgnn_hop_fwd wrote) when it is non-NULL, else from d_y; one of the two is
 * required with MPGNN_F_RELU.  With the bitmask the tensor-core kernels gate g_y while
 * loading it and g_z is never written to memory.  d_gx must not overlap d_gy or any other
 * argument: the projection writes g_z root^T into it while g_y is still being read, and
 * the transposed aggregation then adds in place, touching only rows with incoming edges. */
int mpgnn_hop_bwd(const mpgnn_graph* g, int64_t relation, const float* d_x, const float* d_h, const float* d_y,
                  const uint32_t* d_actmask, const float* d_gy, int64_t f_in, const float* d_w, const float* d_root,
                  int64_t f_out, uint32_t flags, double dropout_p, float* d_gx, float* d_gw, float* d_groot,
                  float* d_gbias, void* d_workspace, int64_t workspace_bytes, void* stream);
/* Layout of d_h.  When most buckets of the relation are empty (and the shapes are the tensor-core ones) the hop keeps
 * the aggregated features COMPACT: d_h holds one row per node that has an edge of the relation (ascending node id), the
 * projection runs as  y = act(x root + b + scatter(h_c W))  and the backward contracts only those rows.  The choice is a
 * pure function of (graph, relation, f_in, f_out, flags & (TF32X3 | COMPACT_H | DENSE_H)); mpgnn_hop_fwd and
 * mpgnn_hop_bwd must be given the same values.  Returns the number of rows of d_h the pair uses: num_nodes (dense) or
 * the relation's number of non-empty buckets (compact); negative on error.  d_h always needs room for N x f_in. */
int64_t mpgnn_hop_h_rows(const mpgnn_graph* g, int64_t relation, int64_t f_in, int64_t f_out, uint32_t flags);
/* Non-empty rows of a relation, ascending (device pointer into the handle, *count of them). */
int mpgnn_graph_relation_rows(const mpgnn_graph* g, int64_t relation, const int32_t** d_rows, int64_t* count);
/* Scratch both hop calls need for (N, f_in, f_out). */
int64_t mpgnn_hop_workspace_bytes(int64_t num_nodes, int64_t f_in, int64_t f_out);

/* ---- dense helpers used by the MPNetm head (model.py:220-226) and its backward ---------
 * out[M,N] = epi( A[M,K] @ B + bias ), B(k,n) = d_b[k*ldb_k + n*ldb_n]; relu optional;
 * d_gate (may be NULL): out *= [gate > 0] (ReLU backward).  A row stride lda, out ldo. */
int mpgnn_gemm_rows(const float* d_a, int64_t lda, int64_t m, int64_t k, const float* d_b, int64_t ldb_k,
                    int64_t ldb_n, int64_t n, const float* d_bias, int relu, const float* d_gate, int64_t ldgate,
                    float* d_out, int64_t ldo, void* d_workspace, int64_t workspace_bytes, void* stream);
/* out[K,N] = A[M,K]^T @ B[M,N] (reduction over the M rows, deterministic split);
 * d_colsum (may be NULL) receives colsum(B) [N]. */
int mpgnn_gemm_tn(const float* d_a, int64_t lda, int64_t m, int64_t k, const float* d_b, int64_t ldb, int64_t n,
                  float* d_out, int64_t ldo, float* d_colsum, void* d_workspace, int64_t workspace_bytes,
                  void* stream);
int64_t mpgnn_gemm_workspace_bytes(int64_t m, int64_t k, int64_t n);

/* ---- head: log_softmax + nll on an index set (main.py:1065, 1088, 1106) ----------------
 * d_logits [N,C] -> d_logp [N,C]; *d_loss = -mean_i logp[idx[i], y[i]] (fixed-order sum);
 * d_glogits (may be NULL) [N,C] = d loss / d logits (zero outside idx). */
int mpgnn_logsoftmax_nll(const float* d_logits, int64_t num_nodes, int64_t num_classes, const int64_t* d_idx,
                         const int64_t* d_y, int64_t n_idx, float* d_logp, float* d_loss, float* d_glogits,
                         void* d_workspace, int64_t workspace_bytes, void* stream);
/* K6: macro-F1 of argmax(logp[idx]) vs y, sklearn f1_score(average='macro') semantics
 * (main.py:1090-1099, 1112): labels = union(pred, true); computed on device in double.
 * d_confusion: int32 [C*C] scratch (zeroed by the call); d_f1: double [1]. */
int mpgnn_macro_f1(const float* d_logp, int64_t num_classes, const int64_t* d_idx, const int64_t* d_y,
                   int64_t n_idx, int32_t* d_confusion, double* d_f1, void* stream);

/* ---- optimiser: torch.optim.Adam(lr, betas, eps, weight_decay) step (main.py:1119) ----
 * One fused pass over n floats; `step` is the 1-based step count. */
int mpgnn_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                    int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                    void* stream);

/* ---- K5: search-stage relation scorer, non-bag mode ------------------------------------
 * score_relation_parallel (main.py:727-760) = `epochs` x train() (main.py:641-673) of the Score
 * net (model.py:26-125): pred[src] = max over the relation's destinations of w[dst] (first
 * maximum in edge order), MSE(mean) against d_labels[src] (float 0/1 per node), Adam(lr) on w,
 * clamp to [0,1].  Sources = nodes with at least one edge of `relation` (first search iteration,
 * d_source_mask NULL) or the nodes flagged in d_source_mask (uint8 [N]); a flagged node without
 * such an edge predicts 0 and still counts in the mean, as in the reference (main.py:653-656).
 * d_w [N] in/out (initial weights: only destination entries matter); d_m, d_v [N] Adam state
 * (zeroed by the call); d_loss_traj [epochs] receives the loss computed BEFORE each update (the
 * reference returns the last one); d_argmax_dst int32 [N] receives the destination each source
 * selected in the last forward (-1 for non-sources).  Deterministic (no float atomics). */
int64_t mpgnn_score_workspace_bytes(int64_t num_nodes);
int mpgnn_score_relation(const mpgnn_graph* g, int64_t relation, float* d_w, const float* d_labels,
                         const uint8_t* d_source_mask, int64_t epochs, double lr, float* d_m, float* d_v, float* d_loss_traj, int32_t* d_argmax_dst,
                         void* d_workspace, int64_t workspace_bytes, void* stream);

/* K5, bag mode (score_relation_bags_parallel / retrain_bags, main.py:814-917; OutputLayer.forward
 * BAGS branch, model.py:45-72): ONE restart = `epochs` train() steps.  A bag is a list of source
 * nodes (d_bag_ptr int32 [B+1] into d_bag_src int32, already restricted to sources that have an
 * edge of `relation`); its prediction is max over its sources s of w[argmax_d w[d]*a_s]*a_s with
 * a_s = <x[s], lin>.  MSE(mean) against d_bag_labels [B]; Adam(lr) on d_w [N] (gradient zeroed where
 * d_grad_mask [N] is 0, when use_mask) and on d_lin [F]; both clamped to [0,1].  Outputs of the LAST
 * forward (taken before the last update, like the reference): d_best_dst / d_best_src int32 [B],
 * d_diff [B] (prediction - label), d_src_val [N] (value of every source that was visited);
 * d_loss_traj [epochs].  Gradient sums use integer atomics on 2^48-scaled values: deterministic. */
int64_t mpgnn_score_bags_workspace_bytes(int64_t num_nodes, int64_t num_bags, int64_t feat);
int mpgnn_score_bags(const mpgnn_graph* g, int64_t relation, const int32_t* d_bag_ptr, const int32_t* d_bag_src,
                     int64_t num_bags, const float* d_bag_labels, const float* d_x, int64_t feat, float* d_w,
                     float* d_lin, const uint8_t* d_grad_mask, int use_mask, int64_t epochs, double lr,
                     float* d_loss_traj, int32_t* d_best_dst, int32_t* d_best_src, float* d_diff, float* d_src_val,
                     void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- device-resident candidate trainer -------------------------------------------------
 * mpgnn_parallel_multiple / mpgnn_parallel_multiple_x (main.py:1117-1160): builds the MPNetm stack
 * (model.py:179-228: per metapath one conv per hop with relu + dropout; the metapaths' embeddings
 * concatenated; fc1, relu, fc2, log_softmax), then run() performs `epochs` x (mpgnn_train,
 * mpgnn_validation) -- forward, nll on the train index, backward, Adam, eval forward, validation
 * nll, macro-F1 on train and validation -- on the device, the whole epoch captured once in a CUDA
 * graph.  mpgnn_trainer_create takes ONE metapath (h_relations[n_layers], layer k consumes
 * h_relations[k]); mpgnn_trainer_create_multi takes n_paths of them, flattened, metapath i =
 * h_relations[h_path_ptr[i] .. h_path_ptr[i+1]).  Parameters are exchanged as one flat fp32 array in
 * state_dict order with torch layouts: per metapath, per hop weight[f_in,H], root[f_in,H], bias[H];
 * then fc1.weight[H, H*n_paths], fc1.bias[H], fc2.weight[C,H], fc2.bias[C].  set_params also resets
 * the optimiser state and the epoch counter and must precede run().  The first layer's aggregation
 * mean_r(x) does not depend on the parameters and is computed once per trainer, not per epoch.
 * run(): `mode` bit 0 = replay the epoch as a CUDA graph; bit 1 (MPGNN_TRAINER_VALIDATE_LAST) = run
 * the validation pass only in the last epoch of the call (the reference validates every epoch and
 * returns the last result; the pass has no side effects, so the returned number is the same).
 * h_trace (may be NULL) receives [epochs_done][4] doubles: train loss, validation loss, train
 * macro-F1, validation macro-F1 per epoch (NaN in the last three for epochs that skipped the
 * validation); *h_last_val_f1 is what the reference returns.  d_x and the index arrays must stay
 * alive while the trainer is used. */
#define MPGNN_TRAINER_GRAPH 1
#define MPGNN_TRAINER_VALIDATE_LAST 2
typedef struct mpgnn_trainer mpgnn_trainer;
int mpgnn_trainer_create(const mpgnn_graph* g, const float* d_x, int64_t f_in, int64_t hidden, int64_t num_classes,
                         const int64_t* h_relations, int64_t n_layers, const int64_t* d_train_idx,
                         const int64_t* d_train_y, int64_t n_train, const int64_t* d_val_idx, const int64_t* d_val_y,
                         int64_t n_val, double dropout_p, uint64_t seed, uint32_t flags, int64_t max_epochs,
                         mpgnn_trainer** out);
int mpgnn_trainer_create_multi(const mpgnn_graph* g, const float* d_x, int64_t f_in, int64_t hidden, int64_t num_classes,
                               const int64_t* h_relations, const int64_t* h_path_ptr, int64_t n_paths,
                               const int64_t* d_train_idx, const int64_t* d_train_y, int64_t n_train,
                               const int64_t* d_val_idx, const int64_t* d_val_y, int64_t n_val, double dropout_p,
                               uint64_t seed, uint32_t flags, int64_t max_epochs, mpgnn_trainer** out);
void mpgnn_trainer_free(mpgnn_trainer* t);
int64_t mpgnn_trainer_num_params(const mpgnn_trainer* t);
int mpgnn_trainer_set_params(mpgnn_trainer* t, const float* d_flat, void* stream);
int mpgnn_trainer_get_params(const mpgnn_trainer* t, float* d_flat, void* stream);
int mpgnn_trainer_run(mpgnn_trainer* t, int64_t epochs, double lr, double beta1, double beta2, double eps,
                      double weight_decay, int mode, void* stream, double* h_trace, double* h_last_val_f1);
/* nll + macro-F1 of the current parameters on another index set (mpgnn_test, main.py:1102-1115) */
int mpgnn_trainer_evaluate(mpgnn_trainer* t, const int64_t* d_idx, const int64_t* d_y, int64_t n_idx, void* stream,
                           float* h_loss, double* h_f1);

#ifdef __cplusplus
}
#endif
#endif /* MPGNN_B200_H */
