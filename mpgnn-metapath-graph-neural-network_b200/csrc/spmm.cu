// K2 -- per-hop aggregation along one relation (CSR gather, optional mean).
//
// Restates PyG 2.3.1 propagate(aggr='mean', flow='target_to_source') as the reference calls
// it (mp_rgcn_layer.py:236): out[i,:] = sum_{p in bucket(r,i)} x[idx[p],:] / max(1,deg).
// The per-row sum runs in bucket order = original edge order, i.e. the order the
// reference's CPU scatter_add_ uses, starting from 0 like its zero-initialised buffer.
//
// HBM bound.  A group of LPR lanes owns one output row and covers its features with
// VEC-wide (128-bit for VEC=4) loads; a warp therefore streams 32/LPR consecutive rows per
// iteration (512 B per row at F=128).  Neighbour indices of a row are fetched with one
// coalesced load per 32-edge batch and broadcast with shuffles, so high-degree rows cost
// one index transaction per 32 gathers; gathers are issued 4 deep before the dependent adds.
// Grid = a multiple of the SM count, each warp walks row-groups with a grid stride.
#include "common.cuh"

namespace mpgnn {

template <int VEC>
struct VecT;
template <>
struct VecT<4> { using type = float4; };
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<1> { using type = float; };

template <int VEC>
__device__ __forceinline__ void vadd(typename VecT<VEC>::type& a, const typename VecT<VEC>::type& b);
template <>
__device__ __forceinline__ void vadd<4>(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
template <>
__device__ __forceinline__ void vadd<2>(float2& a, const float2& b) { a.x += b.x; a.y += b.y; }
template <>
__device__ __forceinline__ void vadd<1>(float& a, const float& b) { a += b; }

template <int VEC>
__device__ __forceinline__ void vdiv(typename VecT<VEC>::type& a, float d);
template <>
__device__ __forceinline__ void vdiv<4>(float4& a, float d) { a.x /= d; a.y /= d; a.z /= d; a.w /= d; }
template <>
__device__ __forceinline__ void vdiv<2>(float2& a, float d) { a.x /= d; a.y /= d; }
template <>
__device__ __forceinline__ void vdiv<1>(float& a, float d) { a /= d; }

template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type vzero();
template <>
__device__ __forceinline__ float4 vzero<4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <>
__device__ __forceinline__ float2 vzero<2>() { return make_float2(0.f, 0.f); }
template <>
__device__ __forceinline__ float vzero<1>() { return 0.f; }

// units = feat / VEC vector slots per row; lane u of the group covers slots u, u+LPR, ...
template <int VEC, int LPR>
__global__ void __launch_bounds__(256) spmm_gather_kernel(const int32_t* __restrict__ ptr,
                                                          const int32_t* __restrict__ idx, int64_t n_rows,
                                                          int mean, const float* __restrict__ x, int64_t ldx,
                                                          int units, const float* __restrict__ init,
                                                          int64_t ldinit, float* __restrict__ out, int64_t ldout,
                                                          int skip_deg) {
  using V = typename VecT<VEC>::type;
  constexpr int ROWS_PER_WARP = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;          // lane inside the row group
  const int grp = lane / LPR;          // which of the warp's rows
  const unsigned grp_mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (grp * LPR));
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_groups = (n_rows + ROWS_PER_WARP - 1) / ROWS_PER_WARP;

  for (int64_t gi = warp_id; gi < n_groups; gi += n_warps) {
    const int64_t row = gi * ROWS_PER_WARP + grp;
    const bool row_ok = row < n_rows;
    int32_t beg = 0, end = 0;
    if (row_ok) {
      beg = __ldg(ptr + row);
      end = __ldg(ptr + row + 1);
    }
    const int deg = end - beg;
    if (skip_deg > 0 && deg > skip_deg) continue;   // hub bucket: left to the chunked kernels (uniform per row group)
    for (int u0 = 0; u0 < units; u0 += LPR) {   // feature chunks (one pass when feat <= 32*VEC)
      const int u = u0 + sub;
      const bool u_ok = row_ok && u < units;
      V acc = vzero<VEC>();
      if (init != nullptr && u_ok) acc = *reinterpret_cast<const V*>(init + row * ldinit + (int64_t)u * VEC);
      for (int32_t b = beg; b < end; b += LPR) {
        // one coalesced index load per LPR edges, broadcast inside the row group
        const int32_t my = (b + sub < end) ? __ldg(idx + b + sub) : 0;
        const int cnt = min(LPR, end - b);
        int k = 0;
        for (; k + 4 <= cnt; k += 4) {
          const int32_t c0 = __shfl_sync(grp_mask, my, grp * LPR + k + 0);
          const int32_t c1 = __shfl_sync(grp_mask, my, grp * LPR + k + 1);
          const int32_t c2 = __shfl_sync(grp_mask, my, grp * LPR + k + 2);
          const int32_t c3 = __shfl_sync(grp_mask, my, grp * LPR + k + 3);
          if (u_ok) {
            const V v0 = __ldg(reinterpret_cast<const V*>(x + (int64_t)c0 * ldx + (int64_t)u * VEC));
            const V v1 = __ldg(reinterpret_cast<const V*>(x + (int64_t)c1 * ldx + (int64_t)u * VEC));
            const V v2 = __ldg(reinterpret_cast<const V*>(x + (int64_t)c2 * ldx + (int64_t)u * VEC));
            const V v3 = __ldg(reinterpret_cast<const V*>(x + (int64_t)c3 * ldx + (int64_t)u * VEC));
            vadd<VEC>(acc, v0);
            vadd<VEC>(acc, v1);
            vadd<VEC>(acc, v2);
            vadd<VEC>(acc, v3);
          }
        }
        for (; k < cnt; ++k) {
          const int32_t c = __shfl_sync(grp_mask, my, grp * LPR + k);
          if (u_ok) {
            const V v = __ldg(reinterpret_cast<const V*>(x + (int64_t)c * ldx + (int64_t)u * VEC));
            vadd<VEC>(acc, v);
          }
        }
      }
      if (u_ok) {
        if (mean && deg > 1) vdiv<VEC>(acc, (float)deg);
        *reinterpret_cast<V*>(out + row * ldout + (int64_t)u * VEC) = acc;
      }
    }
  }
}

// In-place accumulation, out[j,:] += sum over bucket j (the transposed aggregation of the backward: g_x already
// holds g_z root^T).  A bucket without edges needs neither a read nor a write, and with E_r << N that is most rows
// (73 % at C4), so this variant never touches them: lane l of a warp fetches the pointers of row base + l (one
// coalesced load per 32 rows instead of a dependent broadcast load per row), a ballot marks the rows with edges,
// and the warp then walks only those, kRowsInFlight per LPR-lane group at a time with their accumulator, index
// and first-gather loads issued together.  Per row the sum starts from the stored value and adds the bucket in
// edge order, exactly like the general kernel with init == out.
template <int LPR>
__global__ void __launch_bounds__(256) spmm_accumulate_kernel(const int32_t* __restrict__ ptr,
                                                              const int32_t* __restrict__ idx, int64_t n_rows, int mean,
                                                              const float* __restrict__ x, int64_t ldx, int units,
                                                              float* out, int64_t ldout, int skip_deg) {
  constexpr int G = 32 / LPR;            // rows a warp works on side by side
  constexpr int J = 2;                   // rows in flight per group (kRowsInFlight)
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int grp = lane / LPR;
  const unsigned grp_mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (grp * LPR));
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;

  for (int64_t base = warp_id * 32; base < n_rows; base += n_warps * 32) {
    const int64_t my_row = base + lane;
    int32_t beg = 0, end = 0;
    if (my_row < n_rows) {
      beg = __ldg(ptr + my_row);
      end = __ldg(ptr + my_row + 1);
    }
    const int my_deg = end - beg;
    unsigned todo = __ballot_sync(0xffffffffu, my_deg > 0 && !(skip_deg > 0 && my_deg > skip_deg));
    while (todo != 0u) {                 // warp uniform
      int src[J];
#pragma unroll
      for (int j = 0; j < J; ++j) src[j] = -1;
#pragma unroll
      for (int t = 0; t < J * G; ++t) {  // hand the next J*G rows with edges to the groups
        const int b = todo != 0u ? __ffs(todo) - 1 : -1;
        todo &= todo - 1u;               // 0 stays 0
        if (t % G == grp) src[t / G] = b;
      }
      int32_t rb[J], re[J];
      int64_t row[J];
      bool ok[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        ok[j] = src[j] >= 0;             // uniform inside the group
        rb[j] = __shfl_sync(0xffffffffu, beg, src[j] & 31);
        re[j] = __shfl_sync(0xffffffffu, end, src[j] & 31);
        row[j] = base + (src[j] & 31);
      }
      for (int u0 = 0; u0 < units; u0 += LPR) {
        const int u = u0 + sub;
        const bool u_ok = u < units;
        float4 acc[J], first[J];
        int32_t my[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok[j] && u_ok) acc[j] = *reinterpret_cast<const float4*>(out + row[j] * ldout + (int64_t)u * 4);
          my[j] = (ok[j] && rb[j] + sub < re[j]) ? __ldg(idx + rb[j] + sub) : 0;
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {    // every row's first gather in flight before the first add
          const int32_t c = __shfl_sync(0xffffffffu, my[j], grp * LPR);
          first[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok[j] && u_ok) first[j] = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * ldx + (int64_t)u * 4));
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
          if (!ok[j]) continue;
          vadd<4>(acc[j], first[j]);
          int k = 1;                     // edge 0 of the first batch is in
          for (int32_t b = rb[j]; b < re[j]; b += LPR) {
            const int32_t mine = (b == rb[j]) ? my[j] : ((b + sub < re[j]) ? __ldg(idx + b + sub) : 0);
            const int cnt = min(LPR, re[j] - b);
            for (; k < cnt; ++k) {
              const int32_t c = __shfl_sync(grp_mask, mine, grp * LPR + k);
              if (u_ok) vadd<4>(acc[j], __ldg(reinterpret_cast<const float4*>(x + (int64_t)c * ldx + (int64_t)u * 4)));
            }
            k = 0;
          }
          if (u_ok) {
            const int deg = re[j] - rb[j];
            if (mean && deg > 1) vdiv<4>(acc[j], (float)deg);
            *reinterpret_cast<float4*>(out + row[j] * ldout + (int64_t)u * 4) = acc[j];
          }
        }
      }
    }
  }
}

template <int LPR>
static int launch_accumulate(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                             int units, float* out, int64_t ldout, int skip_deg, cudaStream_t s) {
  int64_t blocks = ceil_div(ceil_div(n_rows, 32), 8);      // 8 warps per block, 32 rows per warp and pass
  const int64_t cap = (int64_t)kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  if (blocks > kNumSMs) blocks = (blocks / kNumSMs) * kNumSMs;
  if (blocks < 1) blocks = 1;
  spmm_accumulate_kernel<LPR><<<(unsigned)blocks, 256, 0, s>>>(ptr, idx, n_rows, mean, x, ldx, units, out, ldout, skip_deg);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

template <int VEC, int LPR>
static int launch_one(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                      int units, const float* init, int64_t ldinit, float* out, int64_t ldout, int skip_deg, cudaStream_t s) {
  constexpr int ROWS_PER_WARP = 32 / LPR;
  const int64_t n_groups = ceil_div(n_rows, ROWS_PER_WARP);
  const int64_t blocks_needed = ceil_div(n_groups, 8);  // 8 warps per block
  // several row-groups per warp once the graph is large; grid a multiple of the SM count
  int64_t blocks = blocks_needed;
  const int64_t cap = (int64_t)kNumSMs * 8 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks > kNumSMs) blocks = (blocks / kNumSMs) * kNumSMs;
  if (blocks < 1) blocks = 1;
  spmm_gather_kernel<VEC, LPR><<<(unsigned)blocks, 256, 0, s>>>(ptr, idx, n_rows, mean, x, ldx, units, init, ldinit,
                                                             out, ldout, skip_deg);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

template <int VEC>
static int launch_vec(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                      int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, int skip_deg,
                      cudaStream_t s) {
  const int units = (int)(feat / VEC);
#define MPGNN_SPMM_CASE(L) \
  return launch_one<VEC, L>(ptr, idx, n_rows, mean, x, ldx, units, init, ldinit, out, ldout, skip_deg, s)
  if (units <= 1) MPGNN_SPMM_CASE(1);
  if (units <= 2) MPGNN_SPMM_CASE(2);
  if (units <= 4) MPGNN_SPMM_CASE(4);
  if (units <= 8) MPGNN_SPMM_CASE(8);
  if (units <= 16) MPGNN_SPMM_CASE(16);
  MPGNN_SPMM_CASE(32);
#undef MPGNN_SPMM_CASE
}


static bool aligned_to(const void* p, int64_t bytes) { return (reinterpret_cast<uintptr_t>(p) % bytes) == 0; }

static int launch_spmm_skip(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                            int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, int skip_deg,
                            cudaStream_t s);

int launch_spmm(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, cudaStream_t s) {
  return launch_spmm_skip(ptr, idx, n_rows, mean, x, ldx, feat, init, ldinit, out, ldout, 0, s);
}

// ---- hub buckets: chunked partial sums + fixed-order finish -----------------------------------------------------
// One warp per chunk of kHeavyDeg edges (lane u covers 128-bit slots u, u+32, ...): partial[c][:] = sum of the chunk's
// gathered rows in bucket order.
__global__ void __launch_bounds__(256) spmm_heavy_chunk_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                               const int32_t* __restrict__ rows,
                                                               const int32_t* __restrict__ chunk_ptr, int n_heavy,
                                                               int64_t n_chunks, const float* __restrict__ x, int64_t ldx,
                                                               int units, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp_id; c < n_chunks; c += n_warps) {
    int lo = 0, hi = n_heavy - 1;                  // last heavy bucket whose first chunk is <= c
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(chunk_ptr + mid) <= c) lo = mid; else hi = mid - 1;
    }
    const int32_t row = __ldg(rows + lo);
    const int32_t beg = __ldg(ptr + row) + (int32_t)(c - __ldg(chunk_ptr + lo)) * kHeavyDeg;
    const int32_t end = min(__ldg(ptr + row + 1), beg + kHeavyDeg);
    for (int u0 = 0; u0 < units; u0 += 32) {
      const int u = u0 + lane;
      const bool u_ok = u < units;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int32_t b = beg; b < end; b += 32) {
        const int32_t my = (b + lane < end) ? __ldg(idx + b + lane) : 0;
        const int cnt = min(32, end - b);
        int k = 0;
        for (; k + 4 <= cnt; k += 4) {
          const int32_t c0 = __shfl_sync(0xffffffffu, my, k), c1 = __shfl_sync(0xffffffffu, my, k + 1);
          const int32_t c2 = __shfl_sync(0xffffffffu, my, k + 2), c3 = __shfl_sync(0xffffffffu, my, k + 3);
          if (u_ok) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c0 * ldx + (int64_t)u * 4));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c1 * ldx + (int64_t)u * 4));
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c2 * ldx + (int64_t)u * 4));
            const float4 v3 = __ldg(reinterpret_cast<const float4*>(x + (int64_t)c3 * ldx + (int64_t)u * 4));
            vadd<4>(acc, v0); vadd<4>(acc, v1); vadd<4>(acc, v2); vadd<4>(acc, v3);
          }
        }
        for (; k < cnt; ++k) {
          const int32_t c0 = __shfl_sync(0xffffffffu, my, k);
          if (u_ok) vadd<4>(acc, __ldg(reinterpret_cast<const float4*>(x + (int64_t)c0 * ldx + (int64_t)u * 4)));
        }
      }
      if (u_ok) *reinterpret_cast<float4*>(partial + (c * units + u) * 4) = acc;
    }
  }
}

// One warp per hub bucket: init + its chunk partials in chunk order, mean, store.
__global__ void __launch_bounds__(256) spmm_heavy_finish_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ rows,
                                                                const int32_t* __restrict__ chunk_ptr, int n_heavy, int mean,
                                                                int units, const float* __restrict__ partial,
                                                                const float* __restrict__ init, int64_t ldinit,
                                                                float* __restrict__ out, int64_t ldout) {
  const int lane = threadIdx.x & 31;
  const int j = (int)((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5));
  if (j >= n_heavy) return;
  const int32_t row = __ldg(rows + j);
  const int deg = __ldg(ptr + row + 1) - __ldg(ptr + row);
  const int64_t c0 = __ldg(chunk_ptr + j), nc = (deg + kHeavyDeg - 1) / kHeavyDeg;
  for (int u = lane; u < units; u += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (init != nullptr) acc = *reinterpret_cast<const float4*>(init + (int64_t)row * ldinit + (int64_t)u * 4);
    for (int64_t c = c0; c < c0 + nc; ++c) vadd<4>(acc, *reinterpret_cast<const float4*>(partial + (c * units + u) * 4));
    if (mean && deg > 1) vdiv<4>(acc, (float)deg);
    *reinterpret_cast<float4*>(out + (int64_t)row * ldout + (int64_t)u * 4) = acc;
  }
}

static int launch_heavy(const int32_t* ptr, const int32_t* idx, const int32_t* rows, const int32_t* chunk_ptr, int n_heavy,
                        int64_t n_chunks, int mean, const float* x, int64_t ldx, int64_t feat, const float* init, int64_t ldinit,
                        float* out, int64_t ldout, cudaStream_t s);

int launch_spmm_graph(const mpgnn_graph_impl* g, int64_t rel, int transpose, int mean, const float* x, int64_t ldx,
                      int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, cudaStream_t s) {
  const int32_t* ptr = (transpose ? g->csc_ptr : g->csr_ptr) + rel * g->n;
  const int32_t* idx = transpose ? g->csc_idx : g->csr_idx;
  const HeavyRows& hv = g->heavy[transpose ? 1 : 0];
  const int64_t h0 = hv.count > 0 ? hv.rel_ptr_host[rel] : 0, h1 = hv.count > 0 ? hv.rel_ptr_host[rel + 1] : 0;
  const bool vec4 = feat % 4 == 0 && ldx % 4 == 0 && ldout % 4 == 0 && (init == nullptr || ldinit % 4 == 0) &&
                    aligned_to(x, 16) && aligned_to(out, 16) && (init == nullptr || aligned_to(init, 16));
  if (h1 == h0 || !vec4)      // no hub bucket in this relation (or a layout only the general kernel takes)
    return launch_spmm_skip(ptr, idx, g->n, mean, x, ldx, feat, init, ldinit, out, ldout, 0, s);
  MPGNN_PROPAGATE(launch_spmm_skip(ptr, idx, g->n, mean, x, ldx, feat, init, ldinit, out, ldout, kHeavyDeg, s));
  return launch_heavy(ptr, idx, hv.rows + h0, hv.chunk_ptr + h0, (int)(h1 - h0), hv.rel_chunks_host[rel], mean, x, ldx, feat, init,
                      ldinit, out, ldout, s);
}

// hub buckets of one relation through the chunked kernels; `ptr`/`rows` name the buckets in the index space `out` uses
static int launch_heavy(const int32_t* ptr, const int32_t* idx, const int32_t* rows, const int32_t* chunk_ptr, int n_heavy,
                        int64_t n_chunks, int mean, const float* x, int64_t ldx, int64_t feat, const float* init, int64_t ldinit,
                        float* out, int64_t ldout, cudaStream_t s) {
  const int units = (int)(feat / 4);
  float* partial = nullptr;
  MPGNN_CUDA_CHECK(cudaMallocAsync(&partial, (size_t)n_chunks * feat * sizeof(float), s));   // stream ordered, hub relations only
  int64_t blocks = ceil_div(n_chunks, 8);
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  spmm_heavy_chunk_kernel<<<(unsigned)blocks, 256, 0, s>>>(ptr, idx, rows, chunk_ptr, n_heavy, n_chunks, x, ldx, units, partial);
  MPGNN_LAUNCH_CHECK();
  spmm_heavy_finish_kernel<<<(unsigned)ceil_div(n_heavy, 8), 256, 0, s>>>(ptr, rows, chunk_ptr, n_heavy, mean, units, partial, init,
                                                                          ldinit, out, ldout);
  MPGNN_LAUNCH_CHECK();
  MPGNN_CUDA_CHECK(cudaFreeAsync(partial, s));
  return MPGNN_OK;
}

// K2 on the compact (DCSR) view: row k of `out` ([nnz_rel, feat]) = mean over the bucket of the relation's k-th non-empty
// row -- the same per-bucket sums in the same order as the dense form, without the 73 % (C4) of all-zero rows.
int launch_spmm_graph_compact(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t ldx, int64_t feat, float* out,
                              int64_t ldout, cudaStream_t s) {
  const int64_t nnz = graph_rel_nnz_rows(g, rel);
  if (nnz == 0) return MPGNN_OK;
  const int32_t* ptr = g->cptr + g->rel_nz_host[rel] + rel;
  const HeavyRows& hv = g->heavy[0];
  const int64_t h0 = hv.count > 0 ? hv.rel_ptr_host[rel] : 0, h1 = hv.count > 0 ? hv.rel_ptr_host[rel + 1] : 0;
  const bool vec4 = feat % 4 == 0 && ldx % 4 == 0 && ldout % 4 == 0 && aligned_to(x, 16) && aligned_to(out, 16);
  if (h1 == h0 || !vec4) return launch_spmm_skip(ptr, g->csr_idx, nnz, 1, x, ldx, feat, nullptr, 0, out, ldout, 0, s);
  MPGNN_PROPAGATE(launch_spmm_skip(ptr, g->csr_idx, nnz, 1, x, ldx, feat, nullptr, 0, out, ldout, kHeavyDeg, s));
  return launch_heavy(ptr, g->csr_idx, hv.rows_compact + h0, hv.chunk_ptr + h0, (int)(h1 - h0), hv.rel_chunks_host[rel], 1, x, ldx,
                      feat, nullptr, 0, out, ldout, s);
}

// K2^T with a compact operand: out[j,:] += sum_{e in E_r, col(e) = j} t_c[rank(row(e)),:], in place; the buckets are the CSC
// ones (dense over the nodes), only the gathered matrix is indexed by rank (csc_cidx).
int launch_spmm_graph_transpose_compact(const mpgnn_graph_impl* g, int64_t rel, const float* t_c, int64_t ldt, int64_t feat,
                                        float* out, int64_t ldout, cudaStream_t s) {
  const int32_t* ptr = g->csc_ptr + rel * g->n;
  const HeavyRows& hv = g->heavy[1];
  const int64_t h0 = hv.count > 0 ? hv.rel_ptr_host[rel] : 0, h1 = hv.count > 0 ? hv.rel_ptr_host[rel + 1] : 0;
  const bool vec4 = feat % 4 == 0 && ldt % 4 == 0 && ldout % 4 == 0 && aligned_to(t_c, 16) && aligned_to(out, 16);
  if (h1 == h0 || !vec4) return launch_spmm_skip(ptr, g->csc_cidx, g->n, 0, t_c, ldt, feat, out, ldout, out, ldout, 0, s);
  MPGNN_PROPAGATE(launch_spmm_skip(ptr, g->csc_cidx, g->n, 0, t_c, ldt, feat, out, ldout, out, ldout, kHeavyDeg, s));
  return launch_heavy(ptr, g->csc_cidx, hv.rows + h0, hv.chunk_ptr + h0, (int)(h1 - h0), hv.rel_chunks_host[rel], 0, t_c, ldt, feat,
                      out, ldout, out, ldout, s);
}

static int launch_spmm_skip(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                            int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, int skip_deg,
                            cudaStream_t s) {
  if (n_rows <= 0 || feat <= 0) return MPGNN_OK;
  MPGNN_REQUIRE(feat <= (1 << 20), MPGNN_ENOTSUP, "spmm: feature width %lld too large", (long long)feat);
  const bool has_init = init != nullptr;
  auto ok = [&](int64_t v) {
    return feat % v == 0 && ldx % v == 0 && ldout % v == 0 && (!has_init || ldinit % v == 0) &&
           aligned_to(x, v * 4) && aligned_to(out, v * 4) && (!has_init || aligned_to(init, v * 4));
  };
  if (ok(4) && init == out && ldinit == ldout && feat >= 32) {   // in place: only rows with edges are touched
    const int units = (int)(feat / 4);
    if (units <= 8) return launch_accumulate<8>(ptr, idx, n_rows, mean, x, ldx, units, out, ldout, skip_deg, s);
    if (units <= 16) return launch_accumulate<16>(ptr, idx, n_rows, mean, x, ldx, units, out, ldout, skip_deg, s);
    return launch_accumulate<32>(ptr, idx, n_rows, mean, x, ldx, units, out, ldout, skip_deg, s);
  }
  if (ok(4)) return launch_vec<4>(ptr, idx, n_rows, mean, x, ldx, feat, init, ldinit, out, ldout, skip_deg, s);
  if (ok(2)) return launch_vec<2>(ptr, idx, n_rows, mean, x, ldx, feat, init, ldinit, out, ldout, skip_deg, s);
  return launch_vec<1>(ptr, idx, n_rows, mean, x, ldx, feat, init, ldinit, out, ldout, skip_deg, s);
}

// out[i,:] = x[i,:] / max(1, deg_r(i)): the mean's normalisation on its own, for callers that aggregate the transposed
// way themselves (backward of the all-relation RGCN baseline: g_x += A_r^T (D_r^-1 g_h)).
__global__ void __launch_bounds__(256) scale_rows_by_degree_kernel(const int32_t* __restrict__ ptr, int64_t n_rows,
                                                                   const float* __restrict__ x, int64_t ldx, int64_t feat,
                                                                   float* __restrict__ out, int64_t ldout) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * feat; i += stride) {
    const int64_t row = i / feat, c = i - row * feat;
    const int deg = __ldg(ptr + row + 1) - __ldg(ptr + row);
    const float v = x[row * ldx + c];
    out[row * ldout + c] = deg > 1 ? v / (float)deg : v;
  }
}

int launch_scale_rows_by_degree(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t ldx, int64_t feat, float* out,
                                int64_t ldout, cudaStream_t s) {
  const int64_t work = g->n * feat;
  if (work <= 0) return MPGNN_OK;
  int64_t blocks = ceil_div(work, 256);
  if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
  scale_rows_by_degree_kernel<<<(unsigned)blocks, 256, 0, s>>>(g->csr_ptr + rel * g->n, g->n, x, ldx, feat, out, ldout);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

}  // namespace mpgnn
