// One metapath hop = CustomRGCNConv.forward (mp_rgcn_layer.py:158-271) + the relu/dropout
// MPNetm applies to it (model.py:210-214), and the matching backward.
// Orchestrates K2 (spmm.cu), K3/K4 dense kernels (dense.cu / proj_tcgen05.cu).
#include "common.cuh"

namespace mpgnn {

int proj_tcgen05_supported(int64_t m, int64_t k1, int64_t k2, int64_t n, uint32_t flags);
int64_t proj_tcgen05_workspace_floats(int64_t k, int64_t n);
int launch_proj_tcgen05_ws(const GemmRowsArgs& a, uint32_t flags, float* b_img, cudaStream_t s);
int wgrad_tcgen05_supported(int64_t m, int64_t k1, int64_t k2, int64_t n, uint32_t flags);
int64_t wgrad_tcgen05_workspace_floats(int64_t m, int64_t n);
int launch_wgrad_tcgen05(const GemmTnArgs& a, float* ws, cudaStream_t s);

static int64_t wgrad_ws_floats(int64_t n, int64_t f_in, int64_t f_out) {
  int64_t a = gemm_tn_partial_floats(n, 2 * f_in + 1, f_out);
  int64_t b = wgrad_tcgen05_supported(n, f_in, f_in, f_out, MPGNN_F_TF32X3) ? wgrad_tcgen05_workspace_floats(n, f_out) : 0;
  return a > b ? a : b;
}

// packed [W;root] followed by its hi/lo UMMA images for the tensor-core path
static int64_t fwd_ws_floats(int64_t f_in, int64_t f_out) {
  return align_up(2 * f_in * f_out, 64) + align_up(proj_tcgen05_workspace_floats(2 * f_in, f_out), 64);
}

// ---- compact hop ---------------------------------------------------------------------------------------------------
// With E_r << N most rows of h = mean_r(x) are zero (73 % at C4).  The compact form keeps h as h_c [nnz_rows, f_in] (one
// row per non-empty bucket, graph_build.cu) and runs everything derived from it on those rows only:
//   forward   h_c = mean_r(x) (compact K2);  hw_c = h_c W;  y = act(x root + b + scatter(hw_c))   (epilogue add by rank)
//   backward  gz_c = g_z[rows with edges];  g_root, g_b = x^T g_z;  g_W = h_c^T gz_c;
//             g_x = g_z root^T;  t_c = (gz_c W^T)/deg;  g_x[j] += sum_{e: col(e)=j} t_c[rank(row(e))]
// i.e. the all-rows contractions run at K = f_in instead of 2 f_in and the h-side ones over nnz rows instead of N:
// ~36 % fewer MMA passes (the 3xTF32 kernels are tensor-bound) and ~12 GB less HBM traffic per hop at C4.
constexpr int64_t kCompactMinRows = 1 << 19;     // below this the extra launches cost more than the rows save

int hop_h_compact(const mpgnn_graph_impl* g, int64_t rel, int64_t f_in, int64_t f_out, uint32_t flags) {
  if ((flags & MPGNN_F_DENSE_H) || !(flags & MPGNN_F_TF32X3) || rel < 0 || rel >= g->r) return 0;
  const int64_t n = g->n;
  if (f_out % 32 != 0) return 0;
  if (!proj_tcgen05_supported(n, f_in, 0, f_out, flags) || !proj_tcgen05_supported(n, f_out, 0, f_in, flags)) return 0;
  if (!wgrad_tcgen05_supported(n, f_in, 0, f_out, flags)) return 0;
  if (flags & MPGNN_F_COMPACT_H) return 1;
  return n >= kCompactMinRows && 2 * graph_rel_nnz_rows(g, rel) <= n;
}

static int hop_fwd_compact(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t f_in, const float* w,
                           const float* root, GemmRowsArgs a, uint32_t flags, float* h_c, float* bp, float* hw_c,
                           cudaStream_t s, bool h_precomputed) {
  const int64_t nnz = graph_rel_nnz_rows(g, rel), f_out = a.n;
  float* img_w = bp + align_up(2 * f_in * f_out, 64);
  float* img_root = img_w + 2 * f_in * f_out;
  if (!h_precomputed) {
    ScopedTimer tm("spmm_mean_fwd", s);
    MPGNN_PROPAGATE(launch_spmm_graph_compact(g, rel, x, f_in, f_in, h_c, f_in, s));
  }
  if (nnz > 0) {
    ScopedTimer tm("proj_fwd_compact_tcgen05", s);
    GemmRowsArgs c{};
    c.a1 = h_c; c.lda1 = f_in; c.k1 = f_in; c.b = w; c.m = nnz; c.n = f_out; c.out = hw_c; c.ldo = f_out;
    MPGNN_PROPAGATE(launch_proj_tcgen05_ws(c, flags, img_w, s));
  }
  ScopedTimer tm("proj_fwd_tcgen05", s);
  a.a1 = x; a.lda1 = f_in; a.k1 = f_in; a.a2 = nullptr; a.lda2 = 0; a.k2 = 0;
  a.b = root;
  a.add_src = hw_c; a.ld_add = f_out;
  a.add_bits = g->grp_bits + rel * g->groups; a.add_rank = g->grp_rank + rel * g->groups;
  a.add_base = (uint32_t)g->rel_nz_host[rel];
  return launch_proj_tcgen05_ws(a, flags, img_root, s);
}

int64_t hop_workspace_bytes(int64_t n, int64_t f_in, int64_t f_out) {
  int64_t floats = 0;
  floats += fwd_ws_floats(f_in, f_out);                                // packed [W;root]
  floats += align_up(n * f_out, 64);                                   // g_z
  floats += align_up(wgrad_ws_floats(n, f_in, f_out), 64);             // split-K partials (SIMT or tcgen05)
  floats += align_up(f_out * 2 * f_in, 64);                            // packed [W^T | root^T]
  floats += align_up(proj_tcgen05_workspace_floats(f_out, 2 * f_in), 64);  // its hi/lo UMMA images
  floats += align_up(n * f_in, 64);                                    // t = g_z W^T / deg (g_z root^T goes straight to g_x)
  return floats * 4 + 8 * 256;
}

int hop_fwd(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t f_in, const float* w, const float* root,
            const float* bias, int64_t f_out, uint32_t flags, double p, uint64_t seed, uint64_t offset,
            const uint8_t* mask_bits, float* h, float* y, uint32_t* actmask, void* ws_ptr, int64_t ws_bytes,
            cudaStream_t s, const uint64_t* offset_ptr, bool h_precomputed) {
  MPGNN_REQUIRE(g && x && w && root && h && y, MPGNN_EINVAL, "hop_fwd: NULL argument");
  MPGNN_REQUIRE(rel >= 0 && rel < g->r, MPGNN_ERANGE, "hop_fwd: relation %lld outside [0,%lld)", (long long)rel,
                (long long)g->r);
  MPGNN_REQUIRE(f_in >= 1 && f_out >= 1, MPGNN_EINVAL, "hop_fwd: bad feature sizes");
  const bool drop_seed = flags & MPGNN_F_DROPOUT_SEED, drop_mask = flags & MPGNN_F_DROPOUT_MASK;
  MPGNN_REQUIRE(!(drop_seed && drop_mask), MPGNN_EINVAL, "hop_fwd: both dropout modes set");
  MPGNN_REQUIRE(!(drop_seed || drop_mask) || (p >= 0.0 && p < 1.0), MPGNN_EINVAL, "hop_fwd: dropout p=%g", p);
  MPGNN_REQUIRE(!drop_mask || mask_bits, MPGNN_EINVAL, "hop_fwd: mask mode without mask");
  MPGNN_REQUIRE(!(flags & MPGNN_F_BF16), MPGNN_ENOTSUP, "hop_fwd: MPGNN_F_BF16 is not built; use MPGNN_F_TF32X3");
  // the backward recovers the dropped elements from [y > 0] (a dropped element is 0, a kept one relu(z)/(1-p)); without
  // the relu y == 0 says nothing about the keep-mask, so that combination is refused instead of differentiated wrongly
  MPGNN_REQUIRE(!(drop_seed || drop_mask) || (flags & MPGNN_F_RELU), MPGNN_ENOTSUP,
                "hop_fwd: dropout without MPGNN_F_RELU is not built (model.py:210-214 always applies relu first)");
  MPGNN_REQUIRE(actmask == nullptr || f_out % 32 == 0, MPGNN_ENOTSUP,
                "hop_fwd: the activation bitmask needs f_out %% 32 == 0 (got %lld)", (long long)f_out);
  Workspace ws(ws_ptr, ws_bytes);
  float* bp = ws.take<float>(fwd_ws_floats(f_in, f_out));
  MPGNN_REQUIRE(bp != nullptr, MPGNN_EINVAL, "hop_fwd: workspace too small");

  if (hop_h_compact(g, rel, f_in, f_out, flags)) {
    float* hw_c = ws.take<float>(align_up(g->n * f_out, 64));        // the g_z slot of the backward, free during the forward
    MPGNN_REQUIRE(hw_c != nullptr, MPGNN_EINVAL, "hop_fwd: workspace too small");
    GemmRowsArgs a{};
    a.m = g->n; a.n = f_out;
    a.bias = bias;
    a.relu = (flags & MPGNN_F_RELU) ? 1 : 0;
    a.dropout_mode = drop_seed ? 1 : (drop_mask ? 2 : 0);
    a.dropout_p = (float)p; a.dropout_scale = (float)(1.0 / (1.0 - p));
    a.dropout_thr16 = dropout_threshold16(p); a.seed = seed; a.offset = offset; a.mask_bits = mask_bits;
    a.offset_ptr = offset_ptr;
    a.out = y; a.ldo = f_out;
    a.actmask_out = actmask;
    return hop_fwd_compact(g, rel, x, f_in, w, root, a, flags, h, bp, hw_c, s, h_precomputed);
  }

  if (!h_precomputed) {      // the caller may hold mean_r(x) already (first layer of a model: x is a constant)
    ScopedTimer tm("spmm_mean_fwd", s);
    MPGNN_PROPAGATE(launch_spmm_graph(g, rel, /*transpose=*/0, /*mean=*/1, x, f_in, f_in, nullptr, 0, h, f_in, s));
  }
  // [W; root] as one [2 f_in, f_out] operand: callers that keep root right behind W (the trainer's parameter block) need
  // no packing at all, the others get two device copies into the scratch
  const bool stacked = root == w + f_in * f_out;
  if (!stacked) {
    MPGNN_CUDA_CHECK(cudaMemcpyAsync(bp, w, (size_t)(f_in * f_out) * 4, cudaMemcpyDeviceToDevice, s));
    MPGNN_CUDA_CHECK(cudaMemcpyAsync(bp + f_in * f_out, root, (size_t)(f_in * f_out) * 4, cudaMemcpyDeviceToDevice, s));
  }

  GemmRowsArgs a{};
  a.a1 = h; a.lda1 = f_in; a.k1 = f_in;
  a.a2 = x; a.lda2 = f_in; a.k2 = f_in;
  a.b = stacked ? w : bp; a.m = g->n; a.n = f_out;
  a.bias = bias;
  a.relu = (flags & MPGNN_F_RELU) ? 1 : 0;
  a.dropout_mode = drop_seed ? 1 : (drop_mask ? 2 : 0);
  a.dropout_p = (float)p; a.dropout_scale = (float)(1.0 / (1.0 - p));
  a.dropout_thr16 = dropout_threshold16(p); a.seed = seed; a.offset = offset; a.mask_bits = mask_bits;
  a.offset_ptr = offset_ptr;
  a.out = y; a.ldo = f_out;
  if ((flags & MPGNN_F_TF32X3) && proj_tcgen05_supported(a.m, a.k1, a.k2, a.n, flags)) {
    ScopedTimer tm("proj_fwd_tcgen05", s);
    a.actmask_out = actmask;                           // written by the epilogue, no extra pass
    return launch_proj_tcgen05_ws(a, flags, bp + align_up(2 * f_in * f_out, 64), s);
  }
  {
    ScopedTimer tm("proj_fwd_simt", s);
    MPGNN_PROPAGATE(launch_gemm_rows(a, s));
  }
  if (actmask != nullptr) {
    ScopedTimer tm("pack_actmask", s);
    MPGNN_PROPAGATE(launch_pack_actmask(y, g->n, f_out, actmask, s));
  }
  return MPGNN_OK;
}

int hop_bwd(const mpgnn_graph_impl* g, int64_t rel, const float* x, const float* h, const float* y,
            const uint32_t* actmask, const float* gy, int64_t f_in, const float* w, const float* root, int64_t f_out,
            uint32_t flags, double p, float* gx, float* gw, float* groot, float* gbias, void* ws_ptr, int64_t ws_bytes,
            cudaStream_t s) {
  MPGNN_REQUIRE(g && x && h && gy && w && root && gw && groot, MPGNN_EINVAL, "hop_bwd: NULL argument");
  MPGNN_REQUIRE(!(flags & MPGNN_F_RELU) || y || actmask, MPGNN_EINVAL, "hop_bwd: RELU needs d_y or d_actmask");
  MPGNN_REQUIRE(actmask == nullptr || f_out % 32 == 0, MPGNN_ENOTSUP,
                "hop_bwd: the activation bitmask needs f_out %% 32 == 0 (got %lld)", (long long)f_out);
  MPGNN_REQUIRE(rel >= 0 && rel < g->r, MPGNN_ERANGE, "hop_bwd: relation %lld outside [0,%lld)", (long long)rel,
                (long long)g->r);
  const bool need_gx = flags & MPGNN_F_NEED_GX;
  MPGNN_REQUIRE(!need_gx || gx, MPGNN_EINVAL, "hop_bwd: NEED_GX without d_gx");
  const bool drop = flags & (MPGNN_F_DROPOUT_SEED | MPGNN_F_DROPOUT_MASK);
  MPGNN_REQUIRE(!(flags & MPGNN_F_BF16), MPGNN_ENOTSUP, "hop_bwd: MPGNN_F_BF16 is not built; use MPGNN_F_TF32X3");
  MPGNN_REQUIRE(!drop || (flags & MPGNN_F_RELU), MPGNN_ENOTSUP,
                "hop_bwd: dropout without MPGNN_F_RELU is not built (the keep-mask is recovered from [y > 0])");
  const int64_t n = g->n;
  Workspace ws(ws_ptr, ws_bytes);
  (void)ws.take<float>(fwd_ws_floats(f_in, f_out));
  float* gz = ws.take<float>(align_up(n * f_out, 64));
  const int64_t part_floats = gemm_tn_partial_floats(n, 2 * f_in + 1, f_out);
  float* partials = ws.take<float>(align_up(wgrad_ws_floats(n, f_in, f_out), 64));
  float* bp2 = ws.take<float>(align_up(f_out * 2 * f_in, 64));
  float* bp2_img = ws.take<float>(align_up(proj_tcgen05_workspace_floats(f_out, 2 * f_in), 64));
  float* t = need_gx ? ws.take<float>(align_up(n * f_in, 64)) : nullptr;
  MPGNN_REQUIRE(gz && partials && bp2 && bp2_img && (!need_gx || t), MPGNN_EINVAL, "hop_bwd: workspace too small");

  if (hop_h_compact(g, rel, f_in, f_out, flags)) {
    // d_h holds h_c [nnz, f_in] (what hop_fwd wrote under the same flags); see the comment above hop_h_compact
    MPGNN_REQUIRE(!(flags & MPGNN_F_RELU) || actmask != nullptr, MPGNN_EINVAL,
                  "hop_bwd: the compact hop takes [y > 0] from d_actmask (pass the bitmask mpgnn_hop_fwd wrote)");
    const int64_t nnz = graph_rel_nnz_rows(g, rel);
    const float sc = drop ? (float)(1.0 / (1.0 - p)) : 1.f;
    const uint32_t* mask = (flags & MPGNN_F_RELU) ? actmask : nullptr;
    float* gz_c = gz;
    if (nnz > 0) {
      ScopedTimer tm("gather_gz_compact", s);
      MPGNN_PROPAGATE(launch_gather_gated_rows(gy, f_out, mask, sc, g->nz_rows + g->rel_nz_host[rel], nnz, f_out, gz_c, s));
    }
    GemmTnArgs tn{};
    tn.a1 = x; tn.lda1 = f_in; tn.k1 = f_in; tn.k2 = 0;
    tn.b = gy; tn.ldb = f_out; tn.n = f_out; tn.m = n;
    tn.out1 = groot; tn.ldo1 = f_out; tn.out_ones = gbias;
    tn.b_actmask = mask; tn.b_scale = sc;
    {
      ScopedTimer tm("wgrad_tn_tcgen05", s);                          // g_root = x^T g_z, g_b = colsum g_z (all rows)
      MPGNN_PROPAGATE(launch_wgrad_tcgen05(tn, partials, s));
    }
    if (nnz > 0) {
      ScopedTimer tm("wgrad_tn_compact_tcgen05", s);                  // g_W = h_c^T gz_c (rows with edges)
      GemmTnArgs tc{};
      tc.a1 = h; tc.lda1 = f_in; tc.k1 = f_in; tc.k2 = 0;
      tc.b = gz_c; tc.ldb = f_out; tc.n = f_out; tc.m = nnz;
      tc.out1 = gw; tc.ldo1 = f_out; tc.out_ones = nullptr; tc.b_scale = 1.f;
      MPGNN_PROPAGATE(launch_wgrad_tcgen05(tc, partials, s));
    } else {
      MPGNN_CUDA_CHECK(cudaMemsetAsync(gw, 0, (size_t)(f_in * f_out) * 4, s));
    }
    if (need_gx) {
      float* wt = bp2;                        // W^T, root^T: [f_out, f_in] each
      float* roott = bp2 + f_out * f_in;
      MPGNN_PROPAGATE(launch_pack_b(wt, f_in, w, 1, f_out, f_out, f_in, s));
      MPGNN_PROPAGATE(launch_pack_b(roott, f_in, root, 1, f_out, f_out, f_in, s));
      {
        ScopedTimer tm("dgrad_nt_tcgen05", s);                        // g_x = g_z root^T (all rows)
        GemmRowsArgs a{};
        a.a1 = gy; a.lda1 = f_out; a.k1 = f_out; a.b = roott; a.m = n; a.n = f_in;
        a.out = gx; a.ldo = f_in;
        a.a1_actmask = mask; a.a1_scale = sc;
        if (mask == nullptr && sc != 1.f) return MPGNN_ENOTSUP;      // unreachable: dropout implies RELU (checked above)
        MPGNN_PROPAGATE(launch_proj_tcgen05_ws(a, flags, bp2_img + 2 * f_out * f_in, s));
      }
      if (nnz > 0) {
        {
          ScopedTimer tm("dgrad_nt_compact_tcgen05", s);              // t_c = (gz_c W^T) / deg (rows with edges)
          GemmRowsArgs a{};
          a.a1 = gz_c; a.lda1 = f_out; a.k1 = f_out; a.b = wt; a.m = nnz; a.n = f_in;
          a.deg_ptr = g->cptr + g->rel_nz_host[rel] + rel; a.deg_cols = f_in;
          a.out = t; a.ldo = f_in;
          MPGNN_PROPAGATE(launch_proj_tcgen05_ws(a, flags, bp2_img, s));
        }
        ScopedTimer tm("spmm_transpose_bwd", s);
        MPGNN_PROPAGATE(launch_spmm_graph_transpose_compact(g, rel, t, f_in, f_in, gx, f_in, s));
      }
    }
    return MPGNN_OK;
  }

  // g_z = g_y * [y > 0] * 1/(1-p).  With the activation bitmask and both tensor-core kernels available
  // the gating is fused into their operand loads (g_z is never materialised); otherwise one pass writes it.
  const bool tc_w = wgrad_tcgen05_supported(n, f_in, f_in, f_out, flags);
  const bool tc_d = (flags & MPGNN_F_TF32X3) && proj_tcgen05_supported(n, f_out, 0, 2 * f_in, flags);
  const float scale = drop ? (float)(1.0 / (1.0 - p)) : 1.f;
  const float* gz_src = gy;
  const uint32_t* fused_mask = nullptr;
  if (flags & MPGNN_F_RELU) {
    if (actmask != nullptr && tc_w && (!need_gx || tc_d)) {
      fused_mask = actmask;
    } else {
      ScopedTimer tm("relu_dropout_bwd", s);
      if (actmask != nullptr) MPGNN_PROPAGATE(launch_relu_dropout_bwd_mask(gy, actmask, scale, gz, n, f_out, s));
      else MPGNN_PROPAGATE(launch_relu_dropout_bwd(gy, y, scale, gz, n * f_out, s));
      gz_src = gz;
    }
  }
  GemmTnArgs tn{};
  tn.a1 = h; tn.lda1 = f_in; tn.k1 = f_in;
  tn.a2 = x; tn.lda2 = f_in; tn.k2 = f_in;
  tn.ones_row = 1;
  tn.b = gz_src; tn.ldb = f_out; tn.n = f_out; tn.m = n;
  tn.out1 = gw; tn.ldo1 = f_out;
  tn.out2 = groot; tn.ldo2 = f_out;
  tn.out_ones = gbias;
  tn.partials = partials; tn.partial_capacity_floats = part_floats;
  tn.b_actmask = fused_mask; tn.b_scale = scale;
  if (tc_w) {
    ScopedTimer tm("wgrad_tn_tcgen05", s);
    MPGNN_PROPAGATE(launch_wgrad_tcgen05(tn, partials, s));
  } else {
    ScopedTimer tm("wgrad_tn_simt", s);
    MPGNN_PROPAGATE(launch_gemm_tn(tn, s));
  }

  if (need_gx) {
    // [t | g_z root^T] = g_z @ [W^T | root^T], first f_in columns divided by deg_r(row).  The two halves go to two
    // tensors: t to the workspace, g_z root^T straight into g_x, so that the transposed aggregation runs in place
    // and touches only the rows of g_x that have incoming messages.
    if (root == w + f_in * f_out) {      // [W; root] contiguous: its transpose [W^T | root^T] in one launch
      MPGNN_PROPAGATE(launch_pack_b(bp2, 2 * f_in, w, 1, f_out, f_out, 2 * f_in, s));
    } else {
      MPGNN_PROPAGATE(launch_pack_b(bp2, 2 * f_in, w, 1, f_out, f_out, f_in, s));
      MPGNN_PROPAGATE(launch_pack_b(bp2 + f_in, 2 * f_in, root, 1, f_out, f_out, f_in, s));
    }
    GemmRowsArgs a{};
    a.a1 = gz_src; a.lda1 = f_out; a.k1 = f_out;
    a.a2 = nullptr; a.lda2 = 0; a.k2 = 0;
    a.b = bp2; a.m = n; a.n = 2 * f_in;
    a.deg_ptr = g->csr_ptr + rel * n; a.deg_cols = f_in;
    a.out = t; a.ldo = f_in;
    a.out2 = gx; a.ldo2 = f_in; a.out_split = f_in;
    a.a1_actmask = fused_mask; a.a1_scale = scale;
    if (tc_d) {
      ScopedTimer tm("dgrad_nt_tcgen05", s);
      MPGNN_PROPAGATE(launch_proj_tcgen05_ws(a, flags, bp2_img, s));
    } else {
      ScopedTimer tm("dgrad_nt_simt", s);
      MPGNN_PROPAGATE(launch_gemm_rows(a, s));
    }
    // g_x[j] += sum_{e: col(e)=j} t[row(e)]
    ScopedTimer tm("spmm_transpose_bwd", s);
    MPGNN_PROPAGATE(launch_spmm_graph(g, rel, /*transpose=*/1, /*mean=*/0, t, f_in, f_in, gx, f_in, gx, f_in, s));
  }
  return MPGNN_OK;
}

}  // namespace mpgnn
