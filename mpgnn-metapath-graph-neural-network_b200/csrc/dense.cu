// Dense fp32 SIMT kernels: the exact-fp32 projection path (K3 baseline), the head GEMMs and
// the deterministic weight-gradient reductions (K4).  Tall-skinny shapes: the node
// dimension is huge, the feature dimensions are <= a few hundred.
//
// These are the general-shape kernels (any M/N/K, e.g. F_in=2 on layer 0, C=2 logits).
// The tensor-core projection for the large aligned shapes lives in proj_tcgen05.cu.
#include "common.cuh"

namespace mpgnn {

// ----------------------------------------------------------------------------------------
// gemm_rows: out[M,N] = epi([A1|A2] @ B), B row-major [K,N]
// ----------------------------------------------------------------------------------------
// Epilogue of one row x 4 consecutive columns (col0 % 4 == 0): bias, 1/deg on the first deg_cols columns, relu,
// ReLU-backward gate, dropout (seeded counter stream or explicit mask bits); shared by every gemm_rows kernel.
__device__ __forceinline__ uint64_t gr_launch_key(const GemmRowsArgs& a) {
  if (a.dropout_mode != 1) return 0ull;
  return dropout_launch_key(a.seed, a.offset + (a.offset_ptr != nullptr ? *a.offset_ptr : 0ull));
}
__device__ __forceinline__ void gr_epilogue(const GemmRowsArgs& a, uint64_t launch_key, int64_t row, int64_t col0,
                                            const float (&acc)[4]) {
  const float scale = a.dropout_mode ? a.dropout_scale : 1.f;
  const int64_t mask_ld = (a.n + 7) / 8;
  float inv_deg_den = 1.f;
  if (a.deg_ptr != nullptr) {
    const int d = __ldg(a.deg_ptr + row + 1) - __ldg(a.deg_ptr + row);
    inv_deg_den = (float)max(d, 1);
  }
  // the 4 columns are one aligned block of the dropout stream: one block word per row
  uint2 dw = make_uint2(0u, 0u);
  if (a.dropout_mode == 1) dw = dropout_block(dropout_row_key(launch_key, (uint64_t)row), (uint32_t)(col0 >> 2));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t col = col0 + j;
    if (col >= a.n) continue;
    float v = acc[j];
    if (a.bias != nullptr) v += __ldg(a.bias + col);
    if (col < a.deg_cols) v = v / inv_deg_den;
    if (a.relu) v = fmaxf(v, 0.f);
    if (a.gate != nullptr) v = (__ldg(a.gate + row * a.ldgate + col) > 0.f) ? v : 0.f;
    if (a.dropout_mode == 1) {
      const uint32_t half = (j & 2) ? dw.y : dw.x;
      v = ((half >> (16 * (j & 1))) & 0xFFFFu) >= a.dropout_thr16 ? v * scale : 0.f;
    } else if (a.dropout_mode == 2) {
      const uint8_t byte = __ldg(a.mask_bits + row * mask_ld + (col >> 3));
      v = ((byte >> (7 - (col & 7))) & 1) ? v * scale : 0.f;
    }
    if (a.out_split > 0 && col >= a.out_split) a.out2[row * a.ldo2 + (col - a.out_split)] = v;
    else a.out[row * a.ldo + col] = v;
  }
}

constexpr int GR_BM = 128, GR_BN = 64, GR_BK = 16, GR_THREADS = 256, GR_PAD = 4;

__global__ void __launch_bounds__(GR_THREADS) gemm_rows_kernel(GemmRowsArgs a) {
  __shared__ __align__(16) float As[GR_BK][GR_BM + GR_PAD];
  __shared__ __align__(16) float Bs[GR_BK][GR_BN];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t row0 = (int64_t)blockIdx.x * GR_BM;
  const int64_t col0 = (int64_t)blockIdx.y * GR_BN;
  const int64_t ktot = a.k1 + a.k2;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lk = tid % GR_BK, lm0 = tid / GR_BK;   // A loader: k fastest (coalesced along the row)
  const int bn = tid % GR_BN, bk0 = tid / GR_BN;   // B loader
  for (int64_t k0 = 0; k0 < ktot; k0 += GR_BK) {
    const int64_t kk = k0 + lk;
#pragma unroll
    for (int p = 0; p < GR_BM / 16; ++p) {
      const int m = lm0 + 16 * p;
      const int64_t row = row0 + m;
      float v = 0.f;
      if (row < a.m) {
        if (kk < a.k1) v = __ldg(a.a1 + row * a.lda1 + kk);
        else if (kk < ktot) v = __ldg(a.a2 + row * a.lda2 + (kk - a.k1));
      }
      As[lk][m] = v;
    }
#pragma unroll
    for (int p = 0; p < GR_BK / 4; ++p) {
      const int k = bk0 + 4 * p;
      const int64_t kg = k0 + k, ng = col0 + bn;
      Bs[k][bn] = (kg < ktot && ng < a.n) ? __ldg(a.b + kg * a.n + ng) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GR_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const uint64_t launch_key = gr_launch_key(a);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = row0 + ty * 8 + i;
    if (row >= a.m) continue;
    gr_epilogue(a, launch_key, row, col0 + tx * 4, acc[i]);
  }
}

// Thin shapes of the candidate models (2-d one-hot input layer, 2-class head): K <= 8 or N <= 8.  The tiled kernel
// above spends its time on tile bookkeeping there; these stream the wide operand once.
// (a) K <= 8: one thread = one row x 4 columns, the <= 8 A values of the row are broadcast loads.
__global__ void __launch_bounds__(256) gemm_rows_smallk_kernel(GemmRowsArgs a) {
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int64_t row = (int64_t)blockIdx.x * 16 + ty;
  const int64_t col0 = (int64_t)blockIdx.y * 64 + tx * 4;
  if (row >= a.m || col0 >= a.n) return;
  const int ktot = (int)(a.k1 + a.k2);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < ktot; ++k) {
    const float av = k < a.k1 ? __ldg(a.a1 + row * a.lda1 + k) : __ldg(a.a2 + row * a.lda2 + (k - a.k1));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col0 + j < a.n) acc[j] = fmaf(av, __ldg(a.b + (int64_t)k * a.n + col0 + j), acc[j]);
  }
  gr_epilogue(a, gr_launch_key(a), row, col0, acc);
}

// (b) N <= 4: 8 lanes per row split K in 16-byte pieces, B (K x N <= 2048 floats) staged in shared memory, the 8
// partial dots are combined with a fixed xor tree, lane 0 of the group runs the epilogue.
__global__ void __launch_bounds__(256) gemm_rows_smalln_kernel(GemmRowsArgs a) {
  __shared__ float Bs[2048];
  const int ktot = (int)(a.k1 + a.k2), n = (int)a.n;
  for (int i = threadIdx.x; i < ktot * n; i += 256) Bs[i] = __ldg(a.b + i);
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const int64_t row = (int64_t)blockIdx.x * 32 + (threadIdx.x >> 3);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (row < a.m) {
    for (int k4 = sub * 4; k4 < ktot; k4 += 32) {
      const float4 v = k4 < a.k1 ? __ldg(reinterpret_cast<const float4*>(a.a1 + row * a.lda1 + k4))
                                 : __ldg(reinterpret_cast<const float4*>(a.a2 + row * a.lda2 + (k4 - a.k1)));
      const float av[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < n) acc[j] = fmaf(av[e], Bs[(k4 + e) * n + j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 4);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
  }
  if (row < a.m && sub == 0) gr_epilogue(a, gr_launch_key(a), row, 0, acc);
}

int launch_gemm_rows(const GemmRowsArgs& a, cudaStream_t s) {
  if (a.m <= 0 || a.n <= 0) return MPGNN_OK;
  MPGNN_REQUIRE(a.k1 >= 0 && a.k2 >= 0 && (a.k1 == 0 || a.a1) && (a.k2 == 0 || a.a2) && a.b && a.out, MPGNN_EINVAL,
                "gemm_rows: bad arguments");
  const int64_t ktot = a.k1 + a.k2;
  if (ktot >= 1 && ktot <= 8 && ceil_div(a.n, 64) <= 65535) {
    dim3 grid((unsigned)ceil_div(a.m, 16), (unsigned)ceil_div(a.n, 64));
    gemm_rows_smallk_kernel<<<grid, 256, 0, s>>>(a);
    MPGNN_LAUNCH_CHECK();
    return MPGNN_OK;
  }
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (a.n <= 4 && ktot * a.n <= 2048 && a.k1 % 4 == 0 && a.k2 % 4 == 0 && (a.k1 == 0 || (al16(a.a1) && a.lda1 % 4 == 0)) &&
      (a.k2 == 0 || (al16(a.a2) && a.lda2 % 4 == 0))) {
    gemm_rows_smalln_kernel<<<(unsigned)ceil_div(a.m, 32), 256, 0, s>>>(a);
    MPGNN_LAUNCH_CHECK();
    return MPGNN_OK;
  }
  const int64_t gx = ceil_div(a.m, GR_BM), gy = ceil_div(a.n, GR_BN);
  MPGNN_REQUIRE(gy <= 65535, MPGNN_ENOTSUP, "gemm_rows: N=%lld too wide", (long long)a.n);
  dim3 grid((unsigned)gx, (unsigned)gy);
  gemm_rows_kernel<<<grid, GR_THREADS, 0, s>>>(a);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

// ----------------------------------------------------------------------------------------
// gemm_tn: out[K,N] = [A1|A2|1]^T @ B over the M rows; fixed split + fixed-order reduction
// ----------------------------------------------------------------------------------------
constexpr int TN_BK = 64, TN_BN = 64, TN_BR = 16, TN_THREADS = 256;

static void tn_split(int64_t m, int64_t ktot, int64_t n, int64_t* splits, int64_t* rows_per_split) {
  // a pure function of the shape => the same summation tree on every run and every GPU count
  const int64_t tiles = ceil_div(ktot, TN_BK) * ceil_div(n, TN_BN);
  int64_t target = (int64_t)kNumSMs * 8 / (tiles > 0 ? tiles : 1);
  if (target < 1) target = 1;
  // thin outputs (<= 8 rows or <= 4 columns) have one tile: give the streaming kernels more, shorter row ranges
  const bool thin = ktot <= 8 || n <= 4;
  int64_t sp = ceil_div(m, thin ? 128 : 512);
  if (thin) target = 2048;
  if (sp > target) sp = target;
  if (sp < 1) sp = 1;
  int64_t rps = align_up(ceil_div(m, sp), TN_BR);
  sp = ceil_div(m, rps);
  if (sp < 1) sp = 1;
  *splits = sp;
  *rows_per_split = rps;
}

int64_t gemm_tn_partial_floats(int64_t m, int64_t ktot, int64_t n) {
  int64_t sp, rps;
  tn_split(m, ktot, n, &sp, &rps);
  return sp * ktot * n;
}

__global__ void __launch_bounds__(TN_THREADS) gemm_tn_kernel(GemmTnArgs a, int64_t ktot, int64_t rows_per_split,
                                                              int tiles_n) {
  __shared__ __align__(16) float As[TN_BR][TN_BK];
  __shared__ __align__(16) float Bs[TN_BR][TN_BN];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int tile_k = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
  const int64_t k0 = (int64_t)tile_k * TN_BK, n0 = (int64_t)tile_n * TN_BN;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(a.m, r_begin + rows_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lc = tid % 64, lr0 = tid / 64;
  const int64_t kk = k0 + lc;
  const int64_t nn = n0 + lc;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += TN_BR) {
#pragma unroll
    for (int p = 0; p < TN_BR / 4; ++p) {
      const int rr = lr0 + 4 * p;
      const int64_t row = r0 + rr;
      float av = 0.f, bv = 0.f;
      if (row < r_end) {
        if (kk < a.k1) av = __ldg(a.a1 + row * a.lda1 + kk);
        else if (kk < a.k1 + a.k2) av = __ldg(a.a2 + row * a.lda2 + (kk - a.k1));
        else if (kk < ktot) av = 1.f;  // the ones column: row ktot-1 of the output = colsum(B)
        if (nn < a.n) bv = __ldg(a.b + row * a.ldb + nn);
      }
      As[rr][lc] = av;
      Bs[rr][lc] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < TN_BR; ++rr) {
      const float4 av = *reinterpret_cast<const float4*>(&As[rr][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[rr][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w};
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* part = a.partials + (int64_t)blockIdx.y * ktot * a.n;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t kr = k0 + ty * 4 + i;
    if (kr >= ktot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t nc = n0 + tx * 4 + j;
      if (nc < a.n) part[kr * a.n + nc] = acc[i][j];
    }
  }
}

// Second pass: one warp per 32 consecutive outputs and kTnRedSeg warps per block, each summing a contiguous segment of
// the splits (coalesced, independent loads in flight); the segment sums are combined in segment order, so the summation
// tree is a function of (splits) only -- the same on every run.
constexpr int kTnRedSeg = 8;
__global__ void __launch_bounds__(32 * kTnRedSeg) gemm_tn_reduce_kernel(GemmTnArgs a, int64_t ktot, int64_t splits) {
  __shared__ float seg_sum[kTnRedSeg][32];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  const int64_t total = ktot * a.n;
  const int64_t s0 = splits * seg / kTnRedSeg, s1 = splits * (seg + 1) / kTnRedSeg;
  float part = 0.f;
  if (i < total)
    for (int64_t sp = s0; sp < s1; ++sp) part += a.partials[sp * total + i];
  seg_sum[seg][lane] = part;
  __syncthreads();
  if (seg != 0 || i >= total) return;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kTnRedSeg; ++k) s += seg_sum[k][lane];
  const int64_t kr = i / a.n, nc = i % a.n;
  if (kr < a.k1) {
    if (a.out1) a.out1[kr * a.ldo1 + nc] = s;
  } else if (kr < a.k1 + a.k2) {
    if (a.out2) a.out2[(kr - a.k1) * a.ldo2 + nc] = s;
  } else if (a.out_ones) {
    a.out_ones[nc] = s;
  }
}

// Thin gemm_tn shapes, same [split][ktot][n] partial layout and the same fixed-order second pass as the tiled kernel.
// (a) ktot <= 8 (input-layer gradients: [h | x | 1]^T g_z with 2-d features): thread = column, 4 row lanes per
//     block; B is read once, coalesced; the A values of a row are broadcast loads.
__global__ void __launch_bounds__(256) gemm_tn_smallk_kernel(GemmTnArgs a, int ktot, int64_t rows_per_split) {
  __shared__ float red[4][8][64];
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int64_t col = (int64_t)blockIdx.x * 64 + c;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(a.m, r_begin + rows_per_split);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (col < a.n) {
    for (int64_t row = r_begin + rl; row < r_end; row += 4) {
      const float bv = __ldg(a.b + row * a.ldb + col);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k >= ktot) break;
        float av;
        if (k < a.k1) av = __ldg(a.a1 + row * a.lda1 + k);
        else if (k < a.k1 + a.k2) av = __ldg(a.a2 + row * a.lda2 + (k - a.k1));
        else av = 1.f;
        acc[k] = fmaf(av, bv, acc[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][k][c] = acc[k];
  __syncthreads();
  if (rl == 0 && col < a.n) {
    float* part = a.partials + (int64_t)blockIdx.y * ktot * a.n;
    for (int k = 0; k < ktot; ++k) part[(int64_t)k * a.n + col] = ((red[0][k][c] + red[1][k][c]) + red[2][k][c]) + red[3][k][c];
  }
}

// (b) n <= 4 (2-class head gradients: [a1 | 1]^T g_logits): thread = row of the output (k < ktot <= 128), 2 row
//     lanes per block; A is read once, coalesced; the <= 4 B values of a row are broadcast loads.
__global__ void __launch_bounds__(256) gemm_tn_smalln_kernel(GemmTnArgs a, int ktot, int64_t rows_per_split) {
  __shared__ float red[2][4][128];
  const int k = threadIdx.x & 127, rl = threadIdx.x >> 7;
  const int n = (int)a.n;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r_end = min(a.m, r_begin + rows_per_split);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < ktot) {
    for (int64_t row = r_begin + rl; row < r_end; row += 2) {
      float av;
      if (k < a.k1) av = __ldg(a.a1 + row * a.lda1 + k);
      else if (k < a.k1 + a.k2) av = __ldg(a.a2 + row * a.lda2 + (k - a.k1));
      else av = 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < n) acc[j] = fmaf(av, __ldg(a.b + row * a.ldb + j), acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[rl][j][k] = acc[j];
  __syncthreads();
  if (rl == 0 && k < ktot) {
    float* part = a.partials + (int64_t)blockIdx.y * ktot * a.n;
    for (int j = 0; j < n; ++j) part[(int64_t)k * a.n + j] = red[0][j][k] + red[1][j][k];
  }
}

int launch_gemm_tn(const GemmTnArgs& a, cudaStream_t s) {
  const int64_t ktot = a.k1 + a.k2 + (a.ones_row ? 1 : 0);
  if (ktot <= 0 || a.n <= 0) return MPGNN_OK;
  MPGNN_REQUIRE(a.m >= 0 && a.b && a.partials, MPGNN_EINVAL, "gemm_tn: bad arguments");
  int64_t splits, rps;
  tn_split(a.m > 0 ? a.m : 1, ktot, a.n, &splits, &rps);
  MPGNN_REQUIRE(splits * ktot * a.n <= a.partial_capacity_floats, MPGNN_EINVAL, "gemm_tn: workspace too small");
  MPGNN_REQUIRE(splits <= 65535, MPGNN_ENOTSUP, "gemm_tn: too many splits");
  const int tiles_n = (int)ceil_div(a.n, TN_BN);
  const int64_t tiles = ceil_div(ktot, TN_BK) * tiles_n;
  if (ktot <= 8) {
    dim3 grid((unsigned)ceil_div(a.n, 64), (unsigned)splits);
    gemm_tn_smallk_kernel<<<grid, 256, 0, s>>>(a, (int)ktot, rps);
  } else if (a.n <= 4 && ktot <= 128) {
    dim3 grid(1, (unsigned)splits);
    gemm_tn_smalln_kernel<<<grid, 256, 0, s>>>(a, (int)ktot, rps);
  } else {
    dim3 grid((unsigned)tiles, (unsigned)splits);
    gemm_tn_kernel<<<grid, TN_THREADS, 0, s>>>(a, ktot, rps, tiles_n);
  }
  MPGNN_LAUNCH_CHECK();
  const int64_t total = ktot * a.n;
  gemm_tn_reduce_kernel<<<(unsigned)ceil_div(total, 32), 32 * kTnRedSeg, 0, s>>>(a, ktot, splits);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

// ----------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------
__global__ void pack_b_kernel(float* dst, int64_t ldd, const float* __restrict__ src, int64_t sk, int64_t sn,
                              int64_t k, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * n) return;
  const int64_t kk = i / n, nn = i % n;
  dst[kk * ldd + nn] = src[kk * sk + nn * sn];
}

int launch_pack_b(float* dst, int64_t ldd, const float* src, int64_t src_ld_k, int64_t src_ld_n, int64_t k,
                  int64_t n, cudaStream_t s) {
  if (k <= 0 || n <= 0) return MPGNN_OK;
  pack_b_kernel<<<(unsigned)ceil_div(k * n, 256), 256, 0, s>>>(dst, ldd, src, src_ld_k, src_ld_n, k, n);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

__global__ void relu_dropout_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, float scale,
                                        float* __restrict__ gz, int64_t count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    gz[i] = (y[i] > 0.f) ? gy[i] * scale : 0.f;
}

__global__ void relu_dropout_bwd_kernel_v4(const float4* __restrict__ gy, const float4* __restrict__ y, float scale,
                                           float4* __restrict__ gz, int64_t count4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride) {
    const float4 g = gy[i], v = y[i];
    float4 o;
    o.x = v.x > 0.f ? g.x * scale : 0.f;
    o.y = v.y > 0.f ? g.y * scale : 0.f;
    o.z = v.z > 0.f ? g.z * scale : 0.f;
    o.w = v.w > 0.f ? g.w * scale : 0.f;
    gz[i] = o;
  }
}

int launch_relu_dropout_bwd(const float* gy, const float* y, float scale, float* gz, int64_t count, cudaStream_t s) {
  if (count <= 0) return MPGNN_OK;
  const bool v4 = count % 4 == 0 && ((uintptr_t)gy % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)gz % 16 == 0);
  const int64_t work = v4 ? count / 4 : count;
  int64_t blocks = ceil_div(work, 256);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  if (v4)
    relu_dropout_bwd_kernel_v4<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(gy),
                                                               reinterpret_cast<const float4*>(y), scale,
                                                               reinterpret_cast<float4*>(gz), work);
  else
    relu_dropout_bwd_kernel<<<(unsigned)blocks, 256, 0, s>>>(gy, y, scale, gz, count);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

// activation bitmask of a contiguous [m, n] matrix (n % 32 == 0): word i bit j = [y[32 i + j] > 0]
__global__ void pack_actmask_kernel(const float* __restrict__ y, int64_t count, uint32_t* __restrict__ mask) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {   // count % 32 == 0
    const uint32_t w = __ballot_sync(0xFFFFFFFFu, y[i] > 0.f);
    if ((threadIdx.x & 31) == 0) mask[i >> 5] = w;
  }
}

int launch_pack_actmask(const float* y, int64_t m, int64_t n, uint32_t* mask, cudaStream_t s) {
  MPGNN_REQUIRE(n % 32 == 0, MPGNN_ENOTSUP, "actmask: the feature width must be a multiple of 32 (got %lld)", (long long)n);
  const int64_t count = m * n;
  if (count <= 0) return MPGNN_OK;
  int64_t blocks = ceil_div(count, 256);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  pack_actmask_kernel<<<(unsigned)blocks, 256, 0, s>>>(y, count, mask);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

__global__ void relu_dropout_bwd_mask_kernel(const float4* __restrict__ gy, const uint32_t* __restrict__ mask, float scale,
                                             float4* __restrict__ gz, int64_t count4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride) {
    const float4 g = gy[i];
    const uint32_t nib = (__ldg(mask + (i >> 3)) >> ((i & 7) * 4)) & 0xFu;
    float4 o;
    o.x = (nib & 1u) ? g.x * scale : 0.f;
    o.y = (nib & 2u) ? g.y * scale : 0.f;
    o.z = (nib & 4u) ? g.z * scale : 0.f;
    o.w = (nib & 8u) ? g.w * scale : 0.f;
    gz[i] = o;
  }
}

int launch_relu_dropout_bwd_mask(const float* gy, const uint32_t* mask, float scale, float* gz, int64_t m, int64_t n,
                                 cudaStream_t s) {
  MPGNN_REQUIRE(n % 32 == 0, MPGNN_ENOTSUP, "actmask: the feature width must be a multiple of 32 (got %lld)", (long long)n);
  MPGNN_REQUIRE(((uintptr_t)gy % 16 == 0) && ((uintptr_t)gz % 16 == 0), MPGNN_EINVAL, "actmask: unaligned gradient");
  const int64_t work = m * n / 4;
  if (work <= 0) return MPGNN_OK;
  int64_t blocks = ceil_div(work, 256);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  relu_dropout_bwd_mask_kernel<<<(unsigned)blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(gy), mask, scale,
                                                                 reinterpret_cast<float4*>(gz), work);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

// g_z restricted to a list of rows, compact: out[k,:] = gate(g_y[rows[k],:]) * scale with gate = the activation bitmask of
// y (null = no gate).  One float4 per thread, a row's 128-bit pieces on consecutive lanes (full 512-byte rows at F = 128).
__global__ void __launch_bounds__(256) gather_gated_rows_kernel(const float* __restrict__ gy, int64_t ldgy,
                                                                const uint32_t* __restrict__ mask, float scale,
                                                                const int32_t* __restrict__ rows, int64_t n_rows, int units,
                                                                float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t work = n_rows * units;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < work; i += stride) {
    const int64_t k = i / units;
    const int u = (int)(i - k * units);
    const int64_t row = __ldg(rows + k);
    float4 v = __ldg(reinterpret_cast<const float4*>(gy + row * ldgy) + u);
    if (mask != nullptr) {
      const uint32_t nib = __ldg(mask + row * (units >> 3) + (u >> 3)) >> ((u & 7) * 4);
      v.x = (nib & 1u) ? v.x * scale : 0.f;
      v.y = (nib & 2u) ? v.y * scale : 0.f;
      v.z = (nib & 4u) ? v.z * scale : 0.f;
      v.w = (nib & 8u) ? v.w * scale : 0.f;
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

int launch_gather_gated_rows(const float* gy, int64_t ldgy, const uint32_t* mask, float scale, const int32_t* rows,
                             int64_t n_rows, int64_t n, float* out, cudaStream_t s) {
  MPGNN_REQUIRE(n % 32 == 0 && ldgy % 4 == 0, MPGNN_ENOTSUP, "gather_gated_rows: the feature width must be a multiple of 32");
  MPGNN_REQUIRE(((uintptr_t)gy % 16 == 0) && ((uintptr_t)out % 16 == 0), MPGNN_EINVAL, "gather_gated_rows: unaligned operand");
  const int units = (int)(n / 4);
  const int64_t work = n_rows * units;
  if (work <= 0) return MPGNN_OK;
  int64_t blocks = ceil_div(work, 256);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  gather_gated_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(gy, ldgy, mask, scale, rows, n_rows, units, out);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

}  // namespace mpgnn
