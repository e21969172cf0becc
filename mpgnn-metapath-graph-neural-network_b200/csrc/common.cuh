// Shared helpers of the mpgnn_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/mpgnn_b200.h"

namespace mpgnn {

void set_error(const char* fmt, ...);
void count_launch();

// Optional per-kernel-class CUDA-event timing on the launching stream (mpgnn_timing_*):
// off by default; bench.py enables it for a separate breakdown pass.
struct ScopedTimer {
  int slot;
  cudaStream_t stream;
  ScopedTimer(const char* name, cudaStream_t s);
  ~ScopedTimer();
};

#define MPGNN_CUDA_CHECK(expr)                                                                      \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      mpgnn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return MPGNN_ECUDA;                                                                           \
    }                                                                                               \
  } while (0)

// every kernel launch site goes through this: counts the launch (bench.py's gpu_launches)
#define MPGNN_LAUNCH_CHECK()                 \
  do {                                       \
    mpgnn::count_launch();                   \
    MPGNN_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

#define MPGNN_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      mpgnn::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define MPGNN_PROPAGATE(expr)     \
  do {                            \
    int _rc = (expr);             \
    if (_rc != MPGNN_OK) return _rc; \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kMaxDynSmem = 232448;   // 227 KB: opt-in dynamic shared memory per CTA on sm_100

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Carves sub-buffers out of a caller-provided workspace (256-byte aligned).
struct Workspace {
  char* base;
  int64_t size;
  int64_t used;
  Workspace(void* p, int64_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(int64_t count) {
    int64_t bytes = align_up(count * (int64_t)sizeof(T), 256);
    if (base == nullptr || used + bytes > size) return nullptr;
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
};

// Buckets with more than kHeavyDeg edges ("hub" rows of a relation, e.g. popular message sources in the transposed
// view of a power-law graph).  One warp walking such a bucket alone sets the duration of the whole aggregation, so
// they are cut into chunks of kHeavyDeg edges summed by separate warps and combined in chunk order (fixed tree).
// Found once per graph: the graph is a run constant.
constexpr int kHeavyDeg = 256;
struct HeavyRows {
  int64_t count;             // heavy buckets over all relations
  int32_t* rows;             // [count] node id of the bucket inside its relation, sorted by (relation, node)  (device)
  int32_t* chunk_ptr;        // [count] first chunk of the bucket, counted inside its relation               (device)
  int64_t* rel_ptr_host;     // [r+1] slice of rows / chunk_ptr belonging to each relation
  int64_t* rel_chunks_host;  // [r]   chunks in each relation
  int32_t* rows_compact;     // [count] rank of the bucket among the non-empty buckets of its relation (CSR side only)
};

struct mpgnn_graph_impl {
  int64_t n, e, r;
  HeavyRows heavy[2];        // [0] CSR buckets (rel,row), [1] CSC buckets (rel,col)
  int32_t* csr_ptr;  // [r*n+1] bucket (rel,row) -> positions in csr_idx
  int32_t* csr_idx;  // [e] message source (col) of each edge, stable order
  int32_t* csr_eid;  // [e] original edge id
  int32_t* csc_ptr;  // [r*n+1] bucket (rel,col)
  int32_t* csc_idx;  // [e] target (row) of each edge
  int32_t* csc_eid;  // [e]
  int64_t* rel_offsets_host;  // [r+1] csr_ptr[k*n] copied to the host
  // ---- compact (DCSR) view of the CSR side: with E_r << N most buckets of a relation are empty (73 % at C4), and
  // everything the hop computes FROM the aggregated features -- h W, h^T g_z, (g_z W^T)/deg -- only exists on the
  // rows with edges.  Rank = position of a row among the non-empty rows of its relation.
  int64_t groups;             // G = ceil(n / 32) 32-row groups per relation
  uint32_t* grp_bits;         // [r*G]   bit l of word (rel, g) = row 32g+l of the relation has edges
  uint32_t* grp_rank;         // [r*G+1] exclusive prefix of popc(grp_bits) over ALL relations (subtract rel_nz_host[rel])
  int32_t* nz_rows;           // [nz]    non-empty rows, relation after relation, ascending inside a relation
  int32_t* cptr;              // [nz+r]  compact row pointers: relation `rel` owns entries rel_nz_host[rel]+rel .. (nnz_rel+1 of them)
  int32_t* csc_cidx;          // [e]     like csc_idx, but the RANK of the target row (what a compact [nnz, F] matrix is indexed by)
  int64_t* rel_nz_host;       // [r+1]   prefix of the number of non-empty rows per relation
};

// ---- internal launchers shared between translation units --------------------------------
int exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, int64_t n, void* d_tmp, int64_t tmp_bytes,
                       cudaStream_t s);
int64_t exclusive_scan_tmp_bytes(int64_t n);

struct GemmRowsArgs {
  // out[M,N] = epi( [A1 | A2][M,K1+K2] @ B[K1+K2,N] )   (B row-major contiguous, ld = N)
  const float* a1; int64_t lda1; int64_t k1;
  const float* a2; int64_t lda2; int64_t k2;
  const float* b;
  int64_t m, n;
  const float* bias;           // [N] or null
  int relu;
  const float* gate; int64_t ldgate;   // out *= [gate>0]
  const int32_t* deg_ptr;      // CSR ptr of the relation (N+1) or null
  int64_t deg_cols;            // out[:, :deg_cols] /= max(1,deg)
  int dropout_mode;            // 0 none, 1 seed, 2 mask bits
  float dropout_p; float dropout_scale;   // scale = (float)(1/(1-p)) formed in double on the host
  uint32_t dropout_thr16;                 // keep iff 16-bit random lane >= thr16 = round(p*65536)
  uint64_t seed; uint64_t offset; const uint8_t* mask_bits;
  const uint64_t* offset_ptr;             // optional device word added to `offset` (CUDA-graph replays)
  float* out; int64_t ldo;
  // optional second destination: columns >= out_split go to out2[:, col - out_split] (0 = everything to `out`;
  // the tcgen05 path needs a multiple of 32)
  float* out2; int64_t ldo2; int64_t out_split;
  // activation bitmask (tcgen05 path only): word (row, c) bit j = [out(row, 32c+j) > 0]
  uint32_t* actmask_out;                  // written by the epilogue when non-null (needs n % 32 == 0)
  const uint32_t* a1_actmask; float a1_scale;   // A1(r,k) := bit(r,k) ? A1(r,k)*a1_scale : 0 (k2 must be 0)
  // tcgen05 path only: out(row,:) += add_src[rank(row),:] (before bias / activation) for the rows flagged in add_bits
  // (32 rows per word); rank(row) = add_rank[row/32] - add_base + number of flagged rows below it in its word
  const float* add_src; int64_t ld_add; const uint32_t* add_bits; const uint32_t* add_rank; uint32_t add_base;
};
int launch_gemm_rows(const GemmRowsArgs& a, cudaStream_t s);

struct GemmTnArgs {
  // out[K1+K2(+1), N] = [A1 | A2 | 1]^T @ B   over the M rows; deterministic split-K.
  const float* a1; int64_t lda1; int64_t k1;
  const float* a2; int64_t lda2; int64_t k2;
  int ones_row;                // append a row of column sums of B
  const float* b; int64_t ldb; int64_t n;
  int64_t m;
  float* out1; int64_t ldo1;   // rows [0,k1)
  float* out2; int64_t ldo2;   // rows [k1,k1+k2)
  float* out_ones;             // [N] (row k1+k2) or null
  float* partials; int64_t partial_capacity_floats;
  const uint32_t* b_actmask; float b_scale;     // tcgen05 path: B(r,c) := bit(r,c) ? B(r,c)*b_scale : 0
};
int launch_gemm_tn(const GemmTnArgs& a, cudaStream_t s);
int64_t gemm_tn_partial_floats(int64_t m, int64_t ktot, int64_t n);

int launch_pack_b(float* dst, int64_t ldd, const float* src, int64_t src_ld_k, int64_t src_ld_n, int64_t k,
                  int64_t n, cudaStream_t s);
int launch_relu_dropout_bwd(const float* gy, const float* y, float scale, float* gz, int64_t count, cudaStream_t s);
int launch_pack_actmask(const float* y, int64_t m, int64_t n, uint32_t* mask, cudaStream_t s);
int launch_relu_dropout_bwd_mask(const float* gy, const uint32_t* mask, float scale, float* gz, int64_t m, int64_t n,
                                 cudaStream_t s);
int launch_gather_gated_rows(const float* gy, int64_t ldgy, const uint32_t* mask, float scale, const int32_t* rows,
                             int64_t n_rows, int64_t n, float* out, cudaStream_t s);

// aggregation over relation `rel` of a graph handle (transpose = CSC view), hub buckets included
int launch_spmm_graph(const mpgnn_graph_impl* g, int64_t rel, int transpose, int mean, const float* x, int64_t ldx,
                      int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, cudaStream_t s);
int launch_spmm(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, cudaStream_t s);
// compact forms: out[k,:] = mean of the bucket of the k-th non-empty row (k < nnz_rel);
// out[j,:] += sum_{e: col(e)=j} t_c[rank(row(e)),:] in place (t_c indexed by rank)
int launch_spmm_graph_compact(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t ldx, int64_t feat, float* out,
                              int64_t ldout, cudaStream_t s);
int launch_spmm_graph_transpose_compact(const mpgnn_graph_impl* g, int64_t rel, const float* t_c, int64_t ldt, int64_t feat,
                                        float* out, int64_t ldout, cudaStream_t s);
int launch_scale_rows_by_degree(const mpgnn_graph_impl* g, int64_t rel, const float* x, int64_t ldx, int64_t feat, float* out,
                                int64_t ldout, cudaStream_t s);
static inline int64_t graph_rel_nnz_rows(const mpgnn_graph_impl* g, int64_t rel) {
  return g->rel_nz_host[rel + 1] - g->rel_nz_host[rel];
}

// ---- device helpers ---------------------------------------------------------------------
// Counter-based dropout randomness, stateless (the stream does not depend on the launch geometry,
// so the SIMT and tcgen05 projections draw identical masks) and cheap where it is hot:
//   launch key = splitmix64(seed ^ splitmix64(offset))              once per launch / thread
//   row key    = splitmix64(launch key + row * GOLDEN)              once per output row
//   block word = 4 Philox-style rounds (32x32->64 multiply, xor) of (lo32(row key) + blk*GOLDEN32,
//                hi32(row key))                                      once per 4 consecutive columns
// i.e. 9 integer instructions per 4 elements in the epilogue.  Element j of block blk = col/4 keeps
// its value iff 16-bit lane j of the 64-bit block word (.x = lanes 0,1; .y = lanes 2,3) is
// >= thr16 = round(p * 65536).
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t dropout_launch_key(uint64_t seed, uint64_t offset) {
  return splitmix64(seed ^ splitmix64(offset));
}
__host__ __device__ __forceinline__ uint64_t dropout_row_key(uint64_t launch_key, uint64_t row) {
  return splitmix64(launch_key + row * 0x9E3779B97F4A7C15ull);
}
__host__ __device__ __forceinline__ uint2 dropout_block(uint64_t row_key, uint32_t blk) {
  uint32_t c0 = (uint32_t)row_key + blk * 0x9E3779B9u, c1 = (uint32_t)(row_key >> 32);
  const uint32_t rk[4] = {0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu};
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint64_t prod = (uint64_t)c0 * 0xD2511F53ull;
    c0 = (uint32_t)(prod >> 32) ^ c1 ^ rk[r];
    c1 = (uint32_t)prod;
  }
  uint2 w;
  w.x = c1; w.y = c0;
  return w;
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, uint64_t row, uint64_t col, uint32_t thr16) {
  const uint2 w = dropout_block(dropout_row_key(dropout_launch_key(seed, offset), row), (uint32_t)(col >> 2));
  const uint32_t half = (col & 2) ? w.y : w.x;
  return ((half >> (16 * (col & 1))) & 0xFFFFu) >= thr16;
}

static inline uint32_t dropout_threshold16(double p) {
  double t = p * 65536.0 + 0.5;
  if (t < 0) t = 0;
  if (t > 65535.0) t = 65535.0;
  return (uint32_t)t;
}

}  // namespace mpgnn
