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

struct mpgnn_graph_impl {
  int64_t n, e, r;
  int32_t* csr_ptr;  // [r*n+1] bucket (rel,row) -> positions in csr_idx
  int32_t* csr_idx;  // [e] message source (col) of each edge, stable order
  int32_t* csr_eid;  // [e] original edge id
  int32_t* csc_ptr;  // [r*n+1] bucket (rel,col)
  int32_t* csc_idx;  // [e] target (row) of each edge
  int32_t* csc_eid;  // [e]
  int64_t* rel_offsets_host;  // [r+1] csr_ptr[k*n] copied to the host
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
  uint64_t seed; uint64_t offset; const uint8_t* mask_bits;
  float* out; int64_t ldo;
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
};
int launch_gemm_tn(const GemmTnArgs& a, cudaStream_t s);
int64_t gemm_tn_partial_floats(int64_t m, int64_t ktot, int64_t n);

int launch_pack_b(float* dst, int64_t ldd, const float* src, int64_t src_ld_k, int64_t src_ld_n, int64_t k,
                  int64_t n, cudaStream_t s);
int launch_relu_dropout_bwd(const float* gy, const float* y, float scale, float* gz, int64_t count, cudaStream_t s);

int launch_spmm(const int32_t* ptr, const int32_t* idx, int64_t n_rows, int mean, const float* x, int64_t ldx,
                int64_t feat, const float* init, int64_t ldinit, float* out, int64_t ldout, cudaStream_t s);

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  *hi = __umulhi(a, b);
  return a * b;
}

// Philox4x32-10 (Salmon et al. 2011): counter-based, so the backward never needs the mask
// stored and the stream does not depend on the launch geometry.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0, hi1;
    uint32_t lo0 = mulhilo32(M0, ctr.x, &hi0);
    uint32_t lo1 = mulhilo32(M1, ctr.z, &hi1);
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// keep-decision of output element `elem` (row*F+col) under (seed, offset): one Philox call
// covers 4 consecutive elements; u in [0,1) from the top 24 bits, keep iff u >= p.
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, uint64_t elem, float p) {
  uint64_t blk = elem >> 2;
  uint4 ctr = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)offset, (uint32_t)(offset >> 32));
  uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t lane = (uint32_t)(elem & 3);
  uint32_t v = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  return (float)(v >> 8) * (1.0f / 16777216.0f) >= p;
}

}  // namespace mpgnn
