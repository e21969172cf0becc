// K1 -- relation-typed CSR / CSC construction.
//
// Replaces the reference's per-call `masked_edge_index(edge_index, edge_type == relation)`
// (mp_rgcn_layer.py:29-37, :231): the whole edge list is bucketed ONCE by
// key = relation*N + node with a stable LSD radix sort (8-bit digits, per-block digit
// histograms, one global exclusive scan, stable in-block ranking), so bucket (r,i) lists
// its edges in original edge order, duplicates kept.  Integer work, HBM bound, bit exact.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace mpgnn {

// ----------------------------------------------------------------------------------------
// device-wide exclusive scan (reduce -> scan of block sums -> downsweep; no inter-block waits)
// ----------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* warp_sums) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t ws = lane < (SCAN_THREADS / 32) ? warp_sums[lane] : 0u;
    uint32_t winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (SCAN_THREADS / 32)) warp_sums[lane] = winc - ws;  // exclusive warp offsets
    if (lane == (SCAN_THREADS / 32) - 1) warp_sums[SCAN_THREADS / 32] = winc;
  }
  __syncthreads();
  *total = warp_sums[SCAN_THREADS / 32];
  return warp_sums[w] + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                                    uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t p = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (p < n) s += in[p];
  }
  uint32_t total;
  block_exclusive_scan(s, &total, warp_sums);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_downsweep_kernel(const uint32_t* __restrict__ in,
                                                                       uint32_t* __restrict__ out, int64_t n,
                                                                       const uint32_t* __restrict__ block_offsets) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0u;
    s += v[i];
  }
  uint32_t total;
  uint32_t run = block_exclusive_scan(s, &total, warp_sums) + (block_offsets ? block_offsets[blockIdx.x] : 0u);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

int64_t exclusive_scan_tmp_bytes(int64_t n) {
  int64_t bytes = 0;
  while (n > 1) {
    n = ceil_div(n, SCAN_TILE);
    bytes += align_up(n * 4, 256);
  }
  return bytes + 256;
}

int exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, int64_t n, void* d_tmp, int64_t tmp_bytes,
                       cudaStream_t s) {
  if (n <= 0) return MPGNN_OK;
  MPGNN_REQUIRE(tmp_bytes >= exclusive_scan_tmp_bytes(n), MPGNN_EINVAL, "exclusive_scan: workspace too small");
  const int64_t nb = ceil_div(n, SCAN_TILE);
  if (nb == 1) {
    scan_downsweep_kernel<<<1, SCAN_THREADS, 0, s>>>(d_in, d_out, n, nullptr);
    MPGNN_LAUNCH_CHECK();
    return MPGNN_OK;
  }
  uint32_t* sums = static_cast<uint32_t*>(d_tmp);
  scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(d_in, n, sums);
  MPGNN_LAUNCH_CHECK();
  char* next = static_cast<char*>(d_tmp) + align_up(nb * 4, 256);
  MPGNN_PROPAGATE(exclusive_scan_u32(sums, sums, nb, next, tmp_bytes - align_up(nb * 4, 256), s));
  scan_downsweep_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(d_in, d_out, n, sums);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

// ----------------------------------------------------------------------------------------
// keys
// ----------------------------------------------------------------------------------------
__global__ void make_keys_kernel(const int64_t* __restrict__ edge_index, const int64_t* __restrict__ edge_type,
                                 int64_t e, int64_t n, int64_t r, int which, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals, int* __restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e) return;
  int64_t rel = edge_type[i];
  int64_t row = edge_index[i];
  int64_t col = edge_index[e + i];
  if (rel < 0 || rel >= r || row < 0 || row >= n || col < 0 || col >= n) {
    *err = 1;
    rel = 0;
    row = 0;
    col = 0;
  }
  keys[i] = (uint32_t)(rel * n + (which == 0 ? row : col));
  vals[i] = (uint32_t)i;
}

__global__ void count_keys_kernel(const uint32_t* __restrict__ keys, int64_t e, uint32_t* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < e) atomicAdd(&counts[keys[i]], 1u);  // integer counts: order independent
}

__global__ void gather_other_kernel(const int64_t* __restrict__ edge_index, int64_t e, int which,
                                    const uint32_t* __restrict__ eid, int32_t* __restrict__ idx,
                                    int32_t* __restrict__ eid_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= e) return;
  uint32_t src = eid[i];
  // CSR (which=0) stores the message source = edge_index[1]; CSC stores the target = edge_index[0]
  idx[i] = (int32_t)edge_index[(which == 0 ? e : 0) + (int64_t)src];
  eid_out[i] = (int32_t)src;
}

// ----------------------------------------------------------------------------------------
// stable LSD radix sort pass (8-bit digit)
// ----------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_WARP_SPAN = 32 * RS_ITEMS;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n,
                                                                 int shift, uint32_t* __restrict__ hist,
                                                                 int64_t nb) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    int64_t p = base + (int64_t)i * RS_THREADS + threadIdx.x;
    if (p < n) atomicAdd(&h[(keys[p] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];  // digit-major for the global scan
}

// Order of the keys inside a tile: warp-major, then iteration, then lane.  The rank of a
// key among equal digits follows exactly that order, which keeps the pass stable.
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
    uint32_t* __restrict__ vals_out, int64_t n, int shift, const uint32_t* __restrict__ offs, int64_t nb) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * RS_WARP_SPAN;
  uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int64_t p = wbase + it * 32 + lane;
    const bool valid = p < n;
    key[it] = valid ? keys_in[p] : 0u;
    val[it] = valid ? vals_in[p] : 0u;
    rank[it] = 0;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const uint32_t d = (key[it] >> shift) & 255u;
      const unsigned peers = __match_any_sync(act, d);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (lane == leader) {
        old = cnt[w][d];
        cnt[w][d] = old + __popc(peers);
      }
      old = __shfl_sync(peers, old, leader);
      rank[it] = old + __popc(peers & ((1u << lane) - 1u));
    }
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // one thread per digit: warp-exclusive offsets + global base
    uint32_t run = offs[(int64_t)d * nb + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      uint32_t c = cnt[ww][d];
      cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    const int64_t p = wbase + it * 32 + lane;
    if (p < n) {
      const uint32_t d = (key[it] >> shift) & 255u;
      const uint32_t pos = cnt[w][d] + rank[it];
      keys_out[pos] = key[it];
      vals_out[pos] = val[it];
    }
  }
}

static int radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int64_t n,
                            int key_bits, uint32_t* hist, void* scan_tmp, int64_t scan_tmp_bytes, cudaStream_t s,
                            uint32_t** keys_sorted, uint32_t** vals_sorted) {
  const int64_t nb = ceil_div(n, RS_TILE);
  uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  for (int shift = 0; shift < key_bits; shift += 8) {
    radix_hist_kernel<<<(unsigned)nb, RS_THREADS, 0, s>>>(ki, n, shift, hist, nb);
    MPGNN_LAUNCH_CHECK();
    MPGNN_PROPAGATE(exclusive_scan_u32(hist, hist, 256 * nb, scan_tmp, scan_tmp_bytes, s));
    radix_scatter_kernel<<<(unsigned)nb, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, shift, hist, nb);
    MPGNN_LAUNCH_CHECK();
    uint32_t* t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  *keys_sorted = ki;
  *vals_sorted = vi;
  return MPGNN_OK;
}

// ---- hub buckets -------------------------------------------------------------------------------
__global__ void count_heavy_kernel(const int32_t* __restrict__ ptr, int64_t buckets, unsigned long long* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < buckets && ptr[i + 1] - ptr[i] > kHeavyDeg) atomicAdd(count, 1ull);
}
__global__ void collect_heavy_kernel(const int32_t* __restrict__ ptr, int64_t buckets, unsigned long long* __restrict__ slot,
                                     uint32_t* __restrict__ bucket_id, int32_t* __restrict__ deg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= buckets) return;
  const int d = ptr[i + 1] - ptr[i];
  if (d > kHeavyDeg) {
    const unsigned long long k = atomicAdd(slot, 1ull);       // order fixed afterwards by a host sort
    bucket_id[k] = (uint32_t)i;
    deg[k] = d;
  }
}

static void free_heavy(HeavyRows* h) {
  cudaFree(h->rows);
  cudaFree(h->chunk_ptr);
  cudaFree(h->rows_compact);
  free(h->rel_ptr_host);
  free(h->rel_chunks_host);
  memset(h, 0, sizeof(*h));
}

// Lists the buckets of `ptr` ([r*n+1]) with more than kHeavyDeg edges: sorted by (relation, node), with the chunk
// each one starts at inside its relation.  Called once per direction after the build (synchronises the stream).
static int find_heavy(const int32_t* ptr, int64_t n, int64_t r, cudaStream_t s, HeavyRows* out) {
  memset(out, 0, sizeof(*out));
  out->rel_ptr_host = static_cast<int64_t*>(calloc(r + 1, sizeof(int64_t)));
  out->rel_chunks_host = static_cast<int64_t*>(calloc(r, sizeof(int64_t)));
  MPGNN_REQUIRE(out->rel_ptr_host && out->rel_chunks_host, MPGNN_ECUDA, "graph_build: host allocation failed");
  const int64_t buckets = r * n;
  unsigned long long* d_count = nullptr;
  MPGNN_CUDA_CHECK(cudaMalloc(&d_count, sizeof(unsigned long long)));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s));
  const unsigned blocks = (unsigned)ceil_div(buckets, 256);
  count_heavy_kernel<<<blocks, 256, 0, s>>>(ptr, buckets, d_count);
  unsigned long long h_count = 0;
  MPGNN_CUDA_CHECK(cudaMemcpyAsync(&h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, s));
  MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
  if (h_count == 0) {
    cudaFree(d_count);
    return MPGNN_OK;
  }
  uint32_t* d_id = nullptr;
  int32_t* d_deg = nullptr;
  MPGNN_CUDA_CHECK(cudaMalloc(&d_id, h_count * 4));
  MPGNN_CUDA_CHECK(cudaMalloc(&d_deg, h_count * 4));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s));
  collect_heavy_kernel<<<blocks, 256, 0, s>>>(ptr, buckets, d_count, d_id, d_deg);
  std::vector<uint32_t> ids(h_count);
  std::vector<int32_t> degs(h_count);
  MPGNN_CUDA_CHECK(cudaMemcpyAsync(ids.data(), d_id, h_count * 4, cudaMemcpyDeviceToHost, s));
  MPGNN_CUDA_CHECK(cudaMemcpyAsync(degs.data(), d_deg, h_count * 4, cudaMemcpyDeviceToHost, s));
  MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
  cudaFree(d_count); cudaFree(d_id); cudaFree(d_deg);
  std::vector<size_t> order(h_count);
  for (size_t i = 0; i < h_count; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ids[a] < ids[b]; });
  std::vector<int32_t> rows(h_count), chunk0(h_count);
  int64_t cur_rel = -1, chunks = 0;
  for (size_t k = 0; k < h_count; ++k) {
    const uint32_t id = ids[order[k]];
    const int64_t rel = id / n;
    while (cur_rel < rel) {                       // close the relations up to `rel`
      if (cur_rel >= 0) out->rel_chunks_host[cur_rel] = chunks;
      ++cur_rel;
      out->rel_ptr_host[cur_rel] = (int64_t)k;
      chunks = 0;
    }
    rows[k] = (int32_t)(id % n);
    chunk0[k] = (int32_t)chunks;
    chunks += ceil_div((int64_t)degs[order[k]], kHeavyDeg);
  }
  if (cur_rel >= 0) out->rel_chunks_host[cur_rel] = chunks;
  for (int64_t q = cur_rel + 1; q <= r; ++q) out->rel_ptr_host[q] = (int64_t)h_count;
  MPGNN_CUDA_CHECK(cudaMalloc(&out->rows, h_count * 4));
  MPGNN_CUDA_CHECK(cudaMalloc(&out->chunk_ptr, h_count * 4));
  MPGNN_CUDA_CHECK(cudaMemcpyAsync(out->rows, rows.data(), h_count * 4, cudaMemcpyHostToDevice, s));
  MPGNN_CUDA_CHECK(cudaMemcpyAsync(out->chunk_ptr, chunk0.data(), h_count * 4, cudaMemcpyHostToDevice, s));
  MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
  out->count = (int64_t)h_count;
  return MPGNN_OK;
}

// ---- compact (DCSR) view of the CSR side ------------------------------------------------------------------------
// one warp per 32-row group of a relation: which of its rows have edges
__global__ void __launch_bounds__(256) group_bits_kernel(const int32_t* __restrict__ ptr, int64_t n, int64_t r, int64_t groups,
                                                         uint32_t* __restrict__ bits, uint32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= r * groups) return;
  const int64_t rel = w / groups, row = (w % groups) * 32 + lane;
  bool nz = false;
  if (row < n) {
    const int64_t b = rel * n + row;
    nz = ptr[b + 1] > ptr[b];
  }
  const uint32_t m = __ballot_sync(0xffffffffu, nz);
  if (lane == 0) {
    bits[w] = m;
    counts[w] = (uint32_t)__popc(m);
  }
}
// the non-empty rows and their row pointers, in rank order: relation `rel` owns nz_rows[base(rel) ..) and
// cptr[base(rel) + rel ..] (one more entry than rows: the end of its last bucket)
__global__ void __launch_bounds__(256) compact_rows_kernel(const int32_t* __restrict__ ptr, int64_t n, int64_t r, int64_t groups,
                                                           const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                                                           int32_t* __restrict__ nz_rows, int32_t* __restrict__ cptr) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= r * groups) return;
  const int64_t rel = w / groups, row = (w % groups) * 32 + lane;
  const uint32_t m = bits[w];
  if ((m >> lane) & 1u) {
    const int64_t pos = (int64_t)rank[w] + __popc(m & ((1u << lane) - 1u));
    nz_rows[pos] = (int32_t)row;
    cptr[pos + rel] = ptr[rel * n + row];
  }
  if (w % groups == 0 && lane == 0) cptr[(int64_t)rank[w + groups] + rel] = ptr[(rel + 1) * n];   // closes the relation
}
// rank of the target row of every edge in CSC order (binary search of the edge's relation in the r+1 offsets)
__global__ void __launch_bounds__(256) compact_index_kernel(const int32_t* __restrict__ csc_idx, int64_t e, int64_t r, int64_t groups,
                                                            const int32_t* __restrict__ csc_ptr, int64_t n,
                                                            const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                                                            int32_t* __restrict__ cidx) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= e) return;
  int64_t lo = 0, hi = r - 1;                      // last relation whose first position is <= p
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if ((int64_t)csc_ptr[mid * n] <= p) lo = mid; else hi = mid - 1;
  }
  const int32_t row = csc_idx[p];
  const int64_t w = lo * groups + (row >> 5);
  cidx[p] = (int32_t)(rank[w] - rank[lo * groups] + __popc(bits[w] & ((1u << (row & 31)) - 1u)));
}
__global__ void heavy_rank_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ rel_of, int64_t count,
                                  int64_t groups, const uint32_t* __restrict__ bits, const uint32_t* __restrict__ rank,
                                  int32_t* __restrict__ rows_compact) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int64_t rel = rel_of[k];
  const int32_t row = rows[k];
  const int64_t w = rel * groups + (row >> 5);
  rows_compact[k] = (int32_t)(rank[w] - rank[rel * groups] + __popc(bits[w] & ((1u << (row & 31)) - 1u)));
}

static int build_compact(mpgnn_graph_impl* g, cudaStream_t s) {
  const int64_t n = g->n, r = g->r, e = g->e;
  const int64_t G = ceil_div(n, 32), words = r * G;
  g->groups = G;
  g->rel_nz_host = static_cast<int64_t*>(calloc(r + 1, sizeof(int64_t)));
  MPGNN_REQUIRE(g->rel_nz_host != nullptr, MPGNN_ECUDA, "graph_build: host allocation failed");
  MPGNN_REQUIRE(words + 1 < (1LL << 31), MPGNN_ENOTSUP, "graph_build: R*N/32 does not fit the group index");
  MPGNN_CUDA_CHECK(cudaMalloc(&g->grp_bits, (size_t)words * 4));
  MPGNN_CUDA_CHECK(cudaMalloc(&g->grp_rank, (size_t)(words + 1) * 4));
  const int64_t scan_bytes = exclusive_scan_tmp_bytes(words + 1);
  void* scan_tmp = nullptr;
  MPGNN_CUDA_CHECK(cudaMalloc(&scan_tmp, (size_t)scan_bytes));
  MPGNN_CUDA_CHECK(cudaMemsetAsync(g->grp_rank, 0, (size_t)(words + 1) * 4, s));
  const unsigned wb = (unsigned)ceil_div(words * 32, 256);
  group_bits_kernel<<<wb, 256, 0, s>>>(g->csr_ptr, n, r, G, g->grp_bits, g->grp_rank);
  MPGNN_LAUNCH_CHECK();
  int rc = exclusive_scan_u32(g->grp_rank, g->grp_rank, words + 1, scan_tmp, scan_bytes, s);
  std::vector<uint32_t> base(r + 1);
  if (rc == MPGNN_OK) {
    cudaError_t ce = cudaMemcpy2DAsync(base.data(), 4, g->grp_rank, (size_t)G * 4, 4, (size_t)(r + 1), cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) { set_error("graph_build: %s", cudaGetErrorString(ce)); rc = MPGNN_ECUDA; }
  }
  cudaFree(scan_tmp);
  MPGNN_PROPAGATE(rc);
  for (int64_t k = 0; k <= r; ++k) g->rel_nz_host[k] = base[k];
  const int64_t nz = g->rel_nz_host[r];
  MPGNN_CUDA_CHECK(cudaMalloc(&g->nz_rows, (size_t)(nz > 0 ? nz : 1) * 4));
  MPGNN_CUDA_CHECK(cudaMalloc(&g->cptr, (size_t)(nz + r) * 4));
  MPGNN_CUDA_CHECK(cudaMalloc(&g->csc_cidx, (size_t)(e > 0 ? e : 1) * 4));
  compact_rows_kernel<<<wb, 256, 0, s>>>(g->csr_ptr, n, r, G, g->grp_bits, g->grp_rank, g->nz_rows, g->cptr);
  MPGNN_LAUNCH_CHECK();
  if (e > 0) {
    compact_index_kernel<<<(unsigned)ceil_div(e, 256), 256, 0, s>>>(g->csc_idx, e, r, G, g->csc_ptr, n, g->grp_bits, g->grp_rank,
                                                                     g->csc_cidx);
    MPGNN_LAUNCH_CHECK();
  }
  HeavyRows& hv = g->heavy[0];
  if (hv.count > 0) {
    std::vector<int32_t> rel_of(hv.count);
    for (int64_t rel = 0; rel < r; ++rel)
      for (int64_t k = hv.rel_ptr_host[rel]; k < hv.rel_ptr_host[rel + 1]; ++k) rel_of[k] = (int32_t)rel;
    int32_t* d_rel = nullptr;
    MPGNN_CUDA_CHECK(cudaMalloc(&d_rel, (size_t)hv.count * 4));
    MPGNN_CUDA_CHECK(cudaMalloc(&hv.rows_compact, (size_t)hv.count * 4));
    MPGNN_CUDA_CHECK(cudaMemcpyAsync(d_rel, rel_of.data(), (size_t)hv.count * 4, cudaMemcpyHostToDevice, s));
    heavy_rank_kernel<<<(unsigned)ceil_div(hv.count, 256), 256, 0, s>>>(hv.rows, d_rel, hv.count, G, g->grp_bits, g->grp_rank,
                                                                        hv.rows_compact);
    MPGNN_LAUNCH_CHECK();
    MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFree(d_rel);
  }
  MPGNN_CUDA_CHECK(cudaStreamSynchronize(s));
  return MPGNN_OK;
}

static void free_graph(mpgnn_graph_impl* g) {
  if (!g) return;
  cudaFree(g->grp_bits);
  cudaFree(g->grp_rank);
  cudaFree(g->nz_rows);
  cudaFree(g->cptr);
  cudaFree(g->csc_cidx);
  free(g->rel_nz_host);
  free_heavy(&g->heavy[0]);
  free_heavy(&g->heavy[1]);
  cudaFree(g->csr_ptr);
  cudaFree(g->csr_idx);
  cudaFree(g->csr_eid);
  cudaFree(g->csc_ptr);
  cudaFree(g->csc_idx);
  cudaFree(g->csc_eid);
  free(g->rel_offsets_host);
  delete g;
}

int graph_build_device(const int64_t* d_edge_index, const int64_t* d_edge_type, int64_t e, int64_t n, int64_t r,
                       cudaStream_t s, mpgnn_graph_impl** out) {
  MPGNN_REQUIRE(out != nullptr, MPGNN_EINVAL, "graph_build: out is NULL");
  MPGNN_REQUIRE(n >= 1 && r >= 1 && e >= 0, MPGNN_EINVAL, "graph_build: need N>=1, R>=1, E>=0 (N=%lld R=%lld E=%lld)",
                (long long)n, (long long)r, (long long)e);
  MPGNN_REQUIRE(e < (1LL << 31), MPGNN_ENOTSUP, "graph_build: E=%lld does not fit int32 positions", (long long)e);
  MPGNN_REQUIRE(n < (1LL << 31) && r * n < (1LL << 32) - 1, MPGNN_ENOTSUP,
                "graph_build: R*N=%lld does not fit the 32-bit (relation,node) key", (long long)(r * n));
  MPGNN_REQUIRE(e == 0 || (d_edge_index && d_edge_type), MPGNN_EINVAL, "graph_build: NULL edge arrays");
  const int64_t nk = r * n + 1;
  int key_bits = 1;
  while ((1LL << key_bits) < r * n) ++key_bits;

  mpgnn_graph_impl* g = new mpgnn_graph_impl();
  memset(g, 0, sizeof(*g));
  g->n = n; g->e = e; g->r = r;
  const int64_t e1 = e > 0 ? e : 1;
  cudaError_t ce = cudaSuccess;
  if (ce == cudaSuccess) ce = cudaMalloc(&g->csr_ptr, nk * 4);
  if (ce == cudaSuccess) ce = cudaMalloc(&g->csc_ptr, nk * 4);
  if (ce == cudaSuccess) ce = cudaMalloc(&g->csr_idx, e1 * 4);
  if (ce == cudaSuccess) ce = cudaMalloc(&g->csr_eid, e1 * 4);
  if (ce == cudaSuccess) ce = cudaMalloc(&g->csc_idx, e1 * 4);
  if (ce == cudaSuccess) ce = cudaMalloc(&g->csc_eid, e1 * 4);
  g->rel_offsets_host = static_cast<int64_t*>(calloc(r + 1, sizeof(int64_t)));
  if (ce != cudaSuccess || !g->rel_offsets_host) {
    set_error("graph_build: allocation failed (%s)", cudaGetErrorString(ce));
    free_graph(g);
    return MPGNN_ECUDA;
  }

  // temporaries (stream ordered)
  const int64_t nb = ceil_div(e1, RS_TILE);
  const int64_t scan_n = (256 * nb > nk) ? 256 * nb : nk;
  const int64_t scan_bytes = exclusive_scan_tmp_bytes(scan_n);
  uint32_t *keys_a = nullptr, *vals_a = nullptr, *keys_b = nullptr, *vals_b = nullptr, *hist = nullptr;
  void* scan_tmp = nullptr;
  int* d_err = nullptr;
  int32_t* h_ptr_samples = nullptr;
  int rc = MPGNN_OK;
#define BUILD_CHECK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess && rc == MPGNN_OK) {                                                    \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));            \
      rc = MPGNN_ECUDA;                                                                           \
    }                                                                                             \
  } while (0)
  BUILD_CHECK(cudaMalloc(&keys_a, e1 * 4));
  BUILD_CHECK(cudaMalloc(&vals_a, e1 * 4));
  BUILD_CHECK(cudaMalloc(&keys_b, e1 * 4));
  BUILD_CHECK(cudaMalloc(&vals_b, e1 * 4));
  BUILD_CHECK(cudaMalloc(&hist, 256 * nb * 4));
  BUILD_CHECK(cudaMalloc(&scan_tmp, scan_bytes));
  BUILD_CHECK(cudaMalloc(&d_err, sizeof(int)));
  h_ptr_samples = static_cast<int32_t*>(malloc((r + 1) * sizeof(int32_t)));
  if (rc == MPGNN_OK) BUILD_CHECK(cudaMemsetAsync(d_err, 0, sizeof(int), s));

  for (int which = 0; which < 2 && rc == MPGNN_OK; ++which) {
    int32_t* ptr = which == 0 ? g->csr_ptr : g->csc_ptr;
    int32_t* idx = which == 0 ? g->csr_idx : g->csc_idx;
    int32_t* eid = which == 0 ? g->csr_eid : g->csc_eid;
    BUILD_CHECK(cudaMemsetAsync(ptr, 0, nk * 4, s));
    if (e > 0 && rc == MPGNN_OK) {
      const unsigned eb = (unsigned)ceil_div(e, 256);
      make_keys_kernel<<<eb, 256, 0, s>>>(d_edge_index, d_edge_type, e, n, r, which, keys_a, vals_a, d_err);
      BUILD_CHECK(cudaGetLastError());
      count_keys_kernel<<<eb, 256, 0, s>>>(keys_a, e, reinterpret_cast<uint32_t*>(ptr));
      BUILD_CHECK(cudaGetLastError());
      uint32_t *ks = nullptr, *vs = nullptr;
      if (rc == MPGNN_OK)
        rc = radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, e, key_bits, hist, scan_tmp, scan_bytes, s, &ks, &vs);
      if (rc == MPGNN_OK) {
        gather_other_kernel<<<eb, 256, 0, s>>>(d_edge_index, e, which, vs, idx, eid);
        BUILD_CHECK(cudaGetLastError());
      }
    }
    if (rc == MPGNN_OK)
      rc = exclusive_scan_u32(reinterpret_cast<uint32_t*>(ptr), reinterpret_cast<uint32_t*>(ptr), nk, scan_tmp,
                              scan_bytes, s);
  }
  int h_err = 0;
  if (rc == MPGNN_OK) {
    BUILD_CHECK(cudaMemcpy2DAsync(h_ptr_samples, 4, g->csr_ptr, (size_t)n * 4, 4, (size_t)(r + 1),
                                  cudaMemcpyDeviceToHost, s));
    BUILD_CHECK(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  BUILD_CHECK(cudaStreamSynchronize(s));
#undef BUILD_CHECK
  cudaFree(keys_a); cudaFree(vals_a); cudaFree(keys_b); cudaFree(vals_b);
  cudaFree(hist); cudaFree(scan_tmp); cudaFree(d_err);
  if (rc == MPGNN_OK && h_err) {
    set_error("graph_build: an edge names a node outside [0,%lld) or a relation outside [0,%lld)", (long long)n,
              (long long)r);
    rc = MPGNN_ERANGE;
  }
  if (rc == MPGNN_OK)
    for (int64_t k = 0; k <= r; ++k) g->rel_offsets_host[k] = h_ptr_samples[k];
  free(h_ptr_samples);
  if (rc == MPGNN_OK) rc = find_heavy(g->csr_ptr, n, r, s, &g->heavy[0]);
  if (rc == MPGNN_OK) rc = find_heavy(g->csc_ptr, n, r, s, &g->heavy[1]);
  if (rc == MPGNN_OK) rc = build_compact(g, s);
  if (rc != MPGNN_OK) {
    free_graph(g);
    return rc;
  }
  *out = g;
  return MPGNN_OK;
}

void graph_free(mpgnn_graph_impl* g) { free_graph(g); }

}  // namespace mpgnn
