// K4 -- weight gradients on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   out1[128,N] = A1[M,128]^T @ B[M,N]     (g_W    = h^T g_z)
//   out2[128,N] = A2[M,128]^T @ B[M,N]     (g_root = x^T g_z)
//   colsum[N]   = sum_rows B               (g_bias)
//
// The reduction runs over the M node rows, so the node dimension is the MMA K dimension and
// both operands are "MN-major" as they sit in HBM (features contiguous).  Same 3xTF32 split as
// the projection kernel (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, fp32 accumulate in TMEM).
//
// Narrow inputs (A1, A2 64 wide -- the hidden-64 models of the candidate trainer): the two operands are
// STACKED along the MMA M dimension into one 128-wide tile ([h | x] per node row), so one accumulator
// and three MMAs per K-step produce both gradients (rows 0-63 = A1^T B, rows 64-127 = A2^T B).
//
// Wide outputs (N = 128, A1/A2 128 wide): the roles are SWAPPED -- g_z^T (128 features) is the M operand and
// [h | x] one 256-wide N operand, D^T[128 x 256] = g_z^T [h | x].  Three M128xN256xK8 MMAs per K-step instead of six
// M128xN128xK8: the same tensor work with 36 KB instead of 48 KB of shared-memory operand reads per K-step, which is
// what bounds this kernel (SS-mode operands).  The epilogue writes the transposed accumulator into the same
// [2][128][N] partial layout.
//
// One persistent CTA per SM owns a contiguous range of rows (a fixed function of M, so the
// summation tree is the same on every run and GPU count) and keeps BOTH 128xN accumulators in
// TMEM for its whole range; at the end it writes one partial per CTA and a second kernel adds
// the partials in CTA order (deterministic split-K).
//  * 16 producer warps stream 32-row chunks of A1, A2 and B: 128-bit loads issued 2 chunks
//    ahead (one fully coalesced 512-byte node row per warp instruction), hi/lo split, conflict-free
//    stores into a 2-stage ring of MN-major UMMA tiles in the SWIZZLE_128B_BASE32B layout -- the only
//    shared-memory layout the hardware accepts for MN-major tf32 operands; they also keep the
//    running column sums of B in registers.
//  * one warp issues the 24 MMAs of a chunk (4 K-steps x 2 accumulators x 3 terms).
//
// ONE 128-wide operand with N = 128 (the compact hop: x^T g_z over all rows, h_c^T g_z over the rows with edges) runs
// wgrad_tma_kernel further down instead: operands streamed by TMA into a raw ring, g_z^T transposed into TENSOR MEMORY
// by the converter warps, only x staged as MMA tiles in shared memory -- 0.86 / 0.99 of the HBM floor at C4 where the
// register-staged producers above reach 0.61.
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace mpgnn {

// proj_tcgen05.cu: fp32 [rows, cols] row-major tensor map, box = box_cols x box_rows (32 columns: SWIZZLE_128B)
int make_tensor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);

namespace tcw {

using namespace tc;

constexpr int kRowsPerChunk = 32;                 // node rows (MMA K) per pipeline stage
constexpr int kFeat = 128;                        // feature width of A1 / A2 (MMA M)
constexpr int kStagesPair = 2;                     // two A operands: 96 KB per stage
constexpr int kProducerWarpsW = 16;
constexpr int kEpiWarpsW = 4;
constexpr int kMmaWarpW = kProducerWarpsW + kEpiWarpsW;           // 20
constexpr int kThreadsW = (kMmaWarpW + 1) * 32;                   // 672
constexpr int kProducerThreadsW = kProducerWarpsW * 32;
constexpr int kATileBytes = kRowsPerChunk * kFeat * 4;            // 16 KB (one of hi / lo)

struct ParamsW {
  const float* a1; int64_t lda1;
  const float* a2; int64_t lda2;
  const float* b; int64_t ldb; int n;
  int64_t m; int64_t rows_per_cta;
  float* partials;       // [grid][2][128][n]
  float* colsum_part;    // [grid][32][n]
  const uint32_t* b_actmask; float b_scale;   // B(r,c) := bit(r,c) ? B(r,c)*b_scale : 0 ([m][n/32] words) or null
};

// instruction descriptor: fp32 accumulate, tf32 x tf32, A and B both MN-major
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
// MN-major tf32 operand tile, SWIZZLE_128B_BASE32B (cute::UMMA::Layout_MN_SW128_32B_Atom):
// atom = 4 K-rows x 32 MN-elements (4 x 128 bytes); inside a row the four 32-byte chunks are
// XOR-swizzled with the K-row index (Swizzle<2,5,2> on the byte address).  Element (mn, k) of a
// tile with `atoms` = width/32 MN-atoms lives at byte
//   (k/4)*atoms*512 + (mn/32)*512 + (k%4)*128 + ((((mn%32)/8) ^ (k%4))*32) + (mn%8)*4.
// Descriptor: LBO = 512 (next MN atom), SBO = atoms*512 (next group of 4 K-rows), layout type 1.
__device__ __forceinline__ uint32_t mn_tile_offset16(int i16, int k, int atoms) {   // i16 = 16-byte unit in the row
  return (uint32_t)((k >> 2) * atoms * 512 + (i16 >> 3) * 512 + (k & 3) * 128 + ((((i16 & 7) >> 1) ^ (k & 3)) << 5) +
                    (i16 & 1) * 16);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((512u >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;   // descriptor version
  d |= 1ull << 61;   // layout type SWIZZLE_128B_BASE32B
  return d;
}

// N = width of B (64 or 128); kMasked: B is gated by the activation bitmask; kStacked: A1, A2 are 64 wide and share
// one 128-wide operand tile / one accumulator
// (ONE 128-wide A operand -- the compact hop -- has its own kernel below: wgrad_tma_kernel)
template <int N, bool kMasked, bool kStacked>
__global__ void __launch_bounds__(kThreadsW, 1) wgrad_tc_kernel(const ParamsW p) {
  constexpr bool kSwap = (N == 128) && !kStacked;      // g_z^T as the M operand, [h | x] as one N = 256 operand
  constexpr int kHxWidth = 2 * kFeat;                  // width of the [h | x] N operand
  constexpr int kStagesW = kStagesPair;
  constexpr int kLoOff = 2 * kATileBytes;              // swapped roles: lo image of the [h | x] tile
  constexpr int kBOff = 4 * kATileBytes;               // B tiles follow the A tiles
  extern __shared__ __align__(1024) uint8_t smem[];
  const int b_tile_bytes = kRowsPerChunk * N * 4;                    // one of hi / lo
  const int stage_bytes = kBOff + 2 * b_tile_bytes;                  // a1 hi/lo, (a2 hi/lo,) b hi/lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStagesW * stage_bytes);
  // bars: full[kStagesW], empty[kStagesW], done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStagesW + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStagesW);
  const uint32_t bar_done = smem_u32(bars + 2 * kStagesW);

  // chunk c of this CTA covers rows (c * gridDim.x + blockIdx.x) * 32 ..: at any moment the CTAs of the grid read
  // one contiguous window of each operand (DRAM page locality; per-CTA contiguous ranges gave 444 scattered
  // streams and 4.0 TB/s).  The assignment is a function of the shape only, so the summation order is fixed.
  const int64_t total_chunks = (p.m + kRowsPerChunk - 1) / kRowsPerChunk;
  const int n_chunks = (int)(total_chunks > blockIdx.x ? (total_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  const int64_t r_end = p.m;

  if (tid == 0) {
    for (int s = 0; s < kStagesW; ++s) {
      mbar_init(bar_full + 8 * s, kProducerWarpsW);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarpW) {  // TMEM: D1 and D2, N fp32 columns each
    const uint32_t ncols = 2 * N;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProducerWarpsW) {
    // ================================ producers ==========================================
    // a warp instruction moves whole node rows: 32 lanes x 16 bytes = one 512-byte row (N = 128) or two
    // 256-byte rows (N = 64): fully coalesced loads; the swizzled destination keeps each quarter-warp
    // inside one 128-byte line, so the stores are conflict free.
    constexpr int n_units_b = kRowsPerChunk * N / 4 / kProducerThreadsW;   // 2 (N=128) or 1 (N=64)
    constexpr int b_atoms = N / 32;
    int a_row[2], a_col[2], a_soff[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      a_row[u] = warp * 2 + u;
      a_col[u] = lane * 4;
      a_soff[u] = (int)mn_tile_offset16(lane, a_row[u], kSwap ? kHxWidth / 32 : kFeat / 32);
    }
    // swapped roles: one [h | x] tile of 8 MN atoms (hi @0, lo @32K); x goes to atoms 4-7 = 16-byte units 32..63
    int a2_soff[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) a2_soff[u] = kSwap ? (int)mn_tile_offset16(lane + 32, a_row[u], 2 * kFeat / 32) : a_soff[u];
    int b_row[2], b_col[2], b_soff[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (N == 128) {
        b_row[u] = warp * 2 + u;
        b_col[u] = lane * 4;
        b_soff[u] = (int)mn_tile_offset16(lane, b_row[u], b_atoms);
      } else {                                          // N == 64: one instruction, two rows per warp
        b_row[u] = warp * 2 + (lane >> 4);
        b_col[u] = (lane & 15) * 4;
        b_soff[u] = (int)mn_tile_offset16(lane & 15, b_row[u], b_atoms);
      }
    }
    float4 csum[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    float4 buf[2][6];
    uint32_t mbuf[2][2] = {{0u, 0u}, {0u, 0u}};            // activation-mask words of the B units in flight
    constexpr bool masked = kMasked;
    constexpr int b_words = N / 32;
    int pf = 0;                                            // next chunk to prefetch
    auto issue = [&](float4 (&dst)[6], uint32_t (&mw)[2]) {
      const int64_t row0 = ((int64_t)pf * gridDim.x + blockIdx.x) * kRowsPerChunk;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t ra = row0 + a_row[u];
        const bool ok = ra < r_end;
        if (kStacked) {      // [A1 row | A2 row]: lanes 0-15 fetch the 64 floats of A1, lanes 16-31 those of A2
          const float* src = lane < 16 ? p.a1 + ra * p.lda1 + a_col[u] : p.a2 + ra * p.lda2 + (a_col[u] - 64);
          dst[u] = (ok && (lane < 16 || p.a2 != nullptr)) ? __ldg(reinterpret_cast<const float4*>(src)) : z;   // a2 may be absent
          dst[2 + u] = z;
        } else {
          dst[u] = ok ? __ldg(reinterpret_cast<const float4*>(p.a1 + ra * p.lda1 + a_col[u])) : z;
          dst[2 + u] = ok ? __ldg(reinterpret_cast<const float4*>(p.a2 + ra * p.lda2 + a_col[u])) : z;
        }
        const int64_t rb = row0 + b_row[u];
        dst[4 + u] = (u < n_units_b && rb < r_end) ? __ldg(reinterpret_cast<const float4*>(p.b + rb * p.ldb + b_col[u])) : z;
        if (masked) mw[u] = (u < n_units_b && rb < r_end) ? __ldg(p.b_actmask + rb * b_words + (b_col[u] >> 5)) : 0u;
      }
      ++pf;
    };
    auto split_store = [&](uint8_t* hi_ptr, int lo_delta, const float4& v) {
      const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      *reinterpret_cast<float4*>(hi_ptr) = h;
      *reinterpret_cast<float4*>(hi_ptr + lo_delta) = l;
    };
    auto store = [&](int s, const float4 (&src)[6], const uint32_t (&mw)[2]) {
      uint8_t* st = smem + (size_t)s * stage_bytes;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (kSwap) {
          split_store(st + a_soff[u], kLoOff, src[u]);                             // [h | x] tile: hi @0, lo @32K
          split_store(st + a2_soff[u], kLoOff, src[2 + u]);
        } else {
          split_store(st + a_soff[u], kATileBytes, src[u]);                        // a1: hi @0, lo @16K
          if (!kStacked) split_store(st + 2 * kATileBytes + a_soff[u], kATileBytes, src[2 + u]);   // a2: hi @32K, lo @48K
        }
        if (u < n_units_b) {
          float4 b = src[4 + u];
          if (masked) {                                   // fused ReLU/dropout backward: g_z = g_y gated by [y > 0]
            const uint32_t nib = mw[u] >> (b_col[u] & 31);
            b.x = (nib & 1u) ? b.x * p.b_scale : 0.f;
            b.y = (nib & 2u) ? b.y * p.b_scale : 0.f;
            b.z = (nib & 4u) ? b.z * p.b_scale : 0.f;
            b.w = (nib & 8u) ? b.w * p.b_scale : 0.f;
          }
          split_store(st + kBOff + b_soff[u], b_tile_bytes, b);                    // b: hi, lo
          csum[u].x += b.x; csum[u].y += b.y; csum[u].z += b.z; csum[u].w += b.w;
        }
      }
    };
    if (0 < n_chunks) issue(buf[0], mbuf[0]);
    if (1 < n_chunks) issue(buf[1], mbuf[1]);
    int s = 0;
    uint32_t sph = 0;
    for (int it0 = 0; it0 < n_chunks; it0 += 2) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int it = it0 + j;
        if (it < n_chunks) {
          mbar_wait(bar_empty + 8 * s, sph ^ 1u);
#ifndef MPGNN_WGRAD_EXP_NOFILL
          store(s, buf[j], mbuf[j]);
          if (it + 2 < n_chunks) issue(buf[j], mbuf[j]);     // registers are free again: next loads go out before the fence
#endif
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full + 8 * s);
          if (++s == kStagesW) { s = 0; sph ^= 1u; }
        }
      }
    }
    // per-thread column sums of B: a [32][N] slab per CTA, reduced in a fixed order afterwards
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (u < n_units_b)
        *reinterpret_cast<float4*>(p.colsum_part + ((int64_t)blockIdx.x * kRowsPerChunk + b_row[u]) * N + b_col[u]) = csum[u];
  } else if (warp == kMmaWarpW) {
    // ================================ MMA issuer ==========================================
    const uint32_t idesc = make_idesc_mn(kFeat, N);
    const uint32_t a_sbo = (kFeat / 32) * 512, b_sbo = (uint32_t)(N / 32) * 512;   // next group of 4 K-rows
    const uint32_t s0 = smem_u32(smem);
    int s = 0;
    uint32_t sph = 0;
    if (elect_one())          // one thread issues for the whole kernel (no per-chunk elect / reconvergence)
    for (int it = 0; it < n_chunks; ++it) {
      mbar_wait(bar_full + 8 * s, sph);
      tc_fence_after();
      {
        const uint32_t st = s0 + (uint32_t)(s * stage_bytes);
#pragma unroll
        for (int kg = 0; kg < kRowsPerChunk / 8; ++kg) {
#ifdef MPGNN_WGRAD_EXP_NOMMA
          if (kg > 0 || it > 0) continue;
#endif
          // one MMA consumes K = 8 node rows = two groups of 4 K-rows
          if (kSwap) {
            const uint32_t hx_sbo = (kHxWidth / 32) * 512;
            const uint32_t go = kg * 2 * b_sbo, ho = kg * 2 * hx_sbo;
            const uint64_t gh = desc_mn(st + kBOff + go, b_sbo), gl = desc_mn(st + kBOff + b_tile_bytes + go, b_sbo);
            const uint64_t hh = desc_mn(st + ho, hx_sbo), hl = desc_mn(st + kLoOff + ho, hx_sbo);
            const uint32_t idesc_t = make_idesc_mn(N, kHxWidth);
            const uint32_t acc_t = (it | kg) != 0 ? 1u : 0u;
            umma_tf32(tmem_base, gh, hh, idesc_t, acc_t);
            umma_tf32(tmem_base, gl, hh, idesc_t, 1u);
            umma_tf32(tmem_base, gh, hl, idesc_t, 1u);
            continue;
          }
          const uint32_t ao = kg * 2 * a_sbo, bo = kg * 2 * b_sbo;
          const uint64_t a1h = desc_mn(st + ao, a_sbo), a1l = desc_mn(st + kATileBytes + ao, a_sbo);
          const uint64_t a2h = desc_mn(st + 2 * kATileBytes + ao, a_sbo);
          const uint64_t a2l = desc_mn(st + 3 * kATileBytes + ao, a_sbo);
          const uint64_t bh = desc_mn(st + kBOff + bo, b_sbo);
          const uint64_t bl = desc_mn(st + kBOff + b_tile_bytes + bo, b_sbo);
          const uint32_t acc = (it | kg) != 0 ? 1u : 0u;
          umma_tf32(tmem_base, a1h, bh, idesc, acc);
          umma_tf32(tmem_base, a1l, bh, idesc, 1u);
          umma_tf32(tmem_base, a1h, bl, idesc, 1u);
          if (!kStacked) {
            umma_tf32(tmem_base + (uint32_t)N, a2h, bh, idesc, acc);
            umma_tf32(tmem_base + (uint32_t)N, a2l, bh, idesc, 1u);
            umma_tf32(tmem_base + (uint32_t)N, a2h, bl, idesc, 1u);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (it == n_chunks - 1) umma_commit(bar_done);
      }
      if (++s == kStagesW) { s = 0; sph ^= 1u; }
    }
    __syncwarp();             // the idle lanes wait here, not at the teardown barrier
  } else {
    // ================================ final epilogue: TMEM -> per-CTA partial =============
    const int ew = warp - kProducerWarpsW;          // == warp % 4: TMEM lane quarter
    float* part = p.partials + (int64_t)blockIdx.x * 2 * kFeat * N;
    const int mrow = ew * 32 + lane;                // feature row of A1 / A2
    if (n_chunks > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    if (kSwap) {      // D^T: lane = output column n, TMEM column j = row of [g_W ; g_root]
      for (int cc = 0; cc < kHxWidth / 32; ++cc) {
        uint32_t v[32];
        if (n_chunks > 0) {
          tmem_ld32(tmem_base + (uint32_t)(cc * 32) + ((uint32_t)(ew * 32) << 16), v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) part[(int64_t)(cc * 32 + q) * N + mrow] = __uint_as_float(v[q]);
      }
    } else
    for (int d = 0; d < (kStacked ? 1 : 2); ++d) {     // stacked: the one accumulator holds [out1 ; out2] = 128 rows
      for (int cc = 0; cc < N / 32; ++cc) {
        uint32_t v[32];
        if (n_chunks > 0) {
          tmem_ld32(tmem_base + (uint32_t)(d * N + cc * 32) + ((uint32_t)(ew * 32) << 16), v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        float* dst = part + ((int64_t)d * kFeat + mrow) * N + cc * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(dst + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpW) {
    tc_fence_after();
    const uint32_t ncols = 2 * N;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ---- one 128-wide operand, TMA-fed, g_z^T in TENSOR MEMORY ---------------------------------------------------------
//   D^T[128 x 128] = g_z^T x      (x = the lone A operand: node features, or the compact h_c)
// The two weight gradients of the compact hop, built like the projection kernel (the one kernel family of this library
// that streams at the HBM rate): nobody holds global loads in registers.
//  * one thread streams 32-row chunks of g_z and x with TMA (4 + 4 boxes of 32 columns x 32 rows, SWIZZLE_128B, rows
//    past M zero filled) and the chunk's activation-mask words with one bulk copy into a 4-deep raw ring (128 KB in
//    flight per SM, twice what the register-staged producers of the kernel above could hold);
//  * 16 converter warps read the raw tiles conflict-free: g_z transposed on the way (warp = 32 features x 8 rows, lane
//    = feature: LDS.32 along the rows), gated by the mask, split hi/lo and tcgen05.st'ed into a ring of 64-column TMEM
//    stages (lane = feature, column = node row); x row-wise (LDS.128), split into hi/lo MN-major UMMA tiles in shared
//    memory.  They keep the column sums of g_z (one feature per thread);
//  * one thread issues tcgen05.mma.kind::tf32 with A = g_z^T from TMEM, B = the x tiles: 3 MMAs per 8 rows, the same
//    terms in the same order as the kernel above; same chunk-to-CTA assignment and partial layout, so the summation
//    tree is still a function of the shape only.
constexpr int kRawStagesT = 4;
constexpr int kOpStagesT = 3;
constexpr int kRawBytesT = 2 * kATileBytes;        // g_z boxes (16 KB) then x boxes (16 KB)
constexpr int kOpBytesT = 2 * kATileBytes;         // x hi | x lo
constexpr int kMaskBytesT = kRowsPerChunk * (kFeat / 32) * 4;      // 512
constexpr int kTmaWarpT = kMmaWarpW + 1;           // 21
constexpr int kThreadsT = (kTmaWarpT + 1) * 32;    // 704
constexpr int kAColsT = 2 * kRowsPerChunk;         // TMEM columns per stage: hi | lo
constexpr int kTmemColsT = 512;                    // 128 (accumulator) + 3 x 64, next power of two
constexpr int kBarsT = 2 * kRawStagesT + 2 * kOpStagesT + 1;
constexpr size_t kSmemT = (size_t)kRawStagesT * (kRawBytesT + kMaskBytesT) + (size_t)kOpStagesT * kOpBytesT + kBarsT * 8 + 16;
static_assert(kSmemT <= (size_t)kMaxDynSmem, "raw ring + operand ring exceed the shared memory of a CTA");

template <bool kMasked>
__global__ void __launch_bounds__(kThreadsT, 1) wgrad_tma_kernel(const ParamsW p, const __grid_constant__ CUtensorMap tmap_x,
                                                                 const __grid_constant__ CUtensorMap tmap_g) {
  constexpr int N = kFeat;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sm_raw = smem;
  uint8_t* sm_op = smem + (size_t)kRawStagesT * kRawBytesT;
  uint8_t* sm_mask = sm_op + (size_t)kOpStagesT * kOpBytesT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_mask + kRawStagesT * kMaskBytesT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBarsT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_rfull = smem_u32(bars), bar_rempty = bar_rfull + 8 * kRawStagesT;
  const uint32_t bar_full = bar_rempty + 8 * kRawStagesT, bar_empty = bar_full + 8 * kOpStagesT;
  const uint32_t bar_done = bar_empty + 8 * kOpStagesT;
  const int64_t total_chunks = (p.m + kRowsPerChunk - 1) / kRowsPerChunk;
  const int n_chunks = (int)(total_chunks > blockIdx.x ? (total_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);

  if (tid == 0) {
    for (int r = 0; r < kRawStagesT; ++r) {
      mbar_init(bar_rfull + 8 * r, 1);                  // one arrive.expect_tx + the byte count of the copies
      mbar_init(bar_rempty + 8 * r, kProducerWarpsW);
    }
    for (int s = 0; s < kOpStagesT; ++s) {
      mbar_init(bar_full + 8 * s, kProducerWarpsW);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarpW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemColsT)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + (uint32_t)N;      // stage ring behind the accumulator

  if (warp < kProducerWarpsW) {
    // ================================ converters ==========================================
    const int q = warp & 3, o = warp >> 2;               // TMEM lane quarter (32 features of g_z), 8-row slice of the chunk
    // g_z: element (row o*8+j, feature q*32+lane) of box q; SWIZZLE_128B: 16-byte unit c of row r sits at c ^ (r % 8)
    int g_off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g_off[j] = q * 4096 + (o * 8 + j) * 128 + (((lane >> 2) ^ j) << 4) + (lane & 3) * 4;
    // x: rows warp*2+u, lane = 16-byte unit of the 512-byte row = (box lane/8, unit lane%8)
    int x_off[2], x_soff[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = warp * 2 + u;
      x_off[u] = kATileBytes + (lane >> 3) * 4096 + r * 128 + (((lane & 7) ^ (r & 7)) << 4);
      x_soff[u] = (int)mn_tile_offset16(lane, r, kFeat / 32);
    }
    const uint32_t my_ta = tmem_a + (uint32_t)(o * 8) + ((uint32_t)(q * 32) << 16);
    float csum = 0.f;
    int rs = 0, s = 0;
    uint32_t rph = 0, sph = 0;
    for (int it = 0; it < n_chunks; ++it) {
      const uint8_t* raw = sm_raw + (size_t)rs * kRawBytesT;
      const uint32_t* mrow = reinterpret_cast<const uint32_t*>(sm_mask + rs * kMaskBytesT) + (o * 8) * (N / 32) + q;
      mbar_wait(bar_rfull + 8 * rs, rph);                 // the chunk's bytes have landed
      float g[8];
      uint32_t mw[8];
      float4 xv[2];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = *reinterpret_cast<const float*>(raw + g_off[j]);
#pragma unroll
      for (int u = 0; u < 2; ++u) xv[u] = *reinterpret_cast<const float4*>(raw + x_off[u]);
      uint32_t dep = 0u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mw[j] = kMasked ? mrow[j * (N / 32)] : 0xFFFFFFFFu;
        dep |= __float_as_uint(g[j]) | mw[j];
      }
      dep |= __float_as_uint(xv[0].x) | __float_as_uint(xv[1].x);
      // hand the raw stage back to the TMA thread only once every load above has delivered (see proj_tcgen05.cu: a
      // plain arrive may overtake generic-proxy loads in flight and the refill would land under them)
      dep = __reduce_or_sync(0xFFFFFFFFu, dep);
      if (lane == 0) mbar_arrive_after(bar_rempty + 8 * rs, dep);
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = g[j];
        if (kMasked) v = ((mw[j] >> lane) & 1u) ? v * p.b_scale : 0.f;   // fused ReLU/dropout backward: g_z = g_y gated by [y > 0]
        csum += v;
        const float h = tf32_hi(v);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(v - h);
      }
      mbar_wait(bar_empty + 8 * s, sph ^ 1u);             // the MMAs that read operand stage s (x tiles, TMEM columns) are done
      tc_fence_after();
      uint8_t* st = sm_op + (size_t)s * kOpBytesT;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 v = xv[u];
        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        *reinterpret_cast<float4*>(st + x_soff[u]) = h;
        *reinterpret_cast<float4*>(st + kATileBytes + x_soff[u]) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
      }
      const uint32_t ta = my_ta + (uint32_t)(s * kAColsT);
      tmem_st8(ta, hi);
      tmem_st8(ta + kRowsPerChunk, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * s);
      if (++rs == kRawStagesT) { rs = 0; rph ^= 1u; }
      if (++s == kOpStagesT) { s = 0; sph ^= 1u; }
    }
    // column sums of g_z: slab row o of this CTA's [32][N] slab holds the 8-row slices' sums, the other rows are zero
    float* slab = p.colsum_part + (int64_t)blockIdx.x * kRowsPerChunk * N;
    const int gcol = q * 32 + lane;
    slab[o * N + gcol] = csum;
#pragma unroll
    for (int r = 0; r < 7; ++r) slab[(4 + o * 7 + r) * N + gcol] = 0.f;
  } else if (warp == kTmaWarpT) {
    // ================================ TMA producer (one lane) ============================
    if (lane == 0) {
      int rs = 0;
      uint32_t rph = 0;
      for (int it = 0; it < n_chunks; ++it) {
        mbar_wait(bar_rempty + 8 * rs, rph ^ 1u);         // converters are done with this raw stage
        const int64_t row0 = ((int64_t)it * gridDim.x + blockIdx.x) * kRowsPerChunk;
        const int64_t rows = p.m - row0 < kRowsPerChunk ? p.m - row0 : kRowsPerChunk;
        const uint32_t dst = smem_u32(sm_raw + (size_t)rs * kRawBytesT);
        const uint32_t bar = bar_rfull + 8 * rs;
        const uint32_t mask_bytes = kMasked ? (uint32_t)rows * (N / 32) * 4 : 0u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kRawBytesT + mask_bytes) : "memory");
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(dst + b * 4096), "l"(reinterpret_cast<uint64_t>(&tmap_g)), "r"(bar), "r"(b * 32), "r"((int)row0)
              : "memory");
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(dst + kATileBytes + b * 4096), "l"(reinterpret_cast<uint64_t>(&tmap_x)), "r"(bar), "r"(b * 32), "r"((int)row0)
              : "memory");
        }
        if (kMasked)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(sm_mask + rs * kMaskBytesT)), "l"(p.b_actmask + row0 * (N / 32)), "r"(mask_bytes), "r"(bar)
                       : "memory");
        if (++rs == kRawStagesT) { rs = 0; rph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarpW) {
    // ================================ MMA issuer ==========================================
    // fp32 accumulate, tf32 x tf32, A from TMEM (K-major by construction), B MN-major
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                               ((uint32_t)(kFeat >> 4) << 24);
    constexpr uint32_t x_sbo = (kFeat / 32) * 512;       // next group of 4 K-rows
    const uint32_t s0 = smem_u32(sm_op);
    int s = 0;
    uint32_t sph = 0;
    if (elect_one())
    for (int it = 0; it < n_chunks; ++it) {
      mbar_wait(bar_full + 8 * s, sph);
      tc_fence_after();
      const uint32_t st = s0 + (uint32_t)(s * kOpBytesT);
      const uint32_t ta = tmem_a + (uint32_t)(s * kAColsT);
#pragma unroll
      for (int kg = 0; kg < kRowsPerChunk / 8; ++kg) {
        const uint32_t xo = kg * 2 * x_sbo;
        const uint64_t xh = desc_mn(st + xo, x_sbo), xl = desc_mn(st + kATileBytes + xo, x_sbo);
        umma_tf32_ts(tmem_base, ta + kg * 8, xh, idesc, (it | kg) != 0 ? 1u : 0u);
        umma_tf32_ts(tmem_base, ta + kRowsPerChunk + kg * 8, xh, idesc, 1u);
        umma_tf32_ts(tmem_base, ta + kg * 8, xl, idesc, 1u);
      }
      umma_commit(bar_empty + 8 * s);
      if (it == n_chunks - 1) umma_commit(bar_done);
      if (++s == kOpStagesT) { s = 0; sph ^= 1u; }
    }
    __syncwarp();
  } else {
    // ================================ final epilogue: TMEM -> per-CTA partial =============
    // D^T: lane = output column n (feature of g_z), TMEM column j = row of the gradient (feature of x)
    const int ew = warp - kProducerWarpsW;
    float* part = p.partials + (int64_t)blockIdx.x * 2 * kFeat * N;
    const int mrow = ew * 32 + lane;
    if (n_chunks > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    for (int cc = 0; cc < kFeat / 32; ++cc) {
      uint32_t v[32];
      if (n_chunks > 0) {
        tmem_ld32(tmem_base + (uint32_t)(cc * 32) + ((uint32_t)(ew * 32) << 16), v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) part[(int64_t)(cc * 32 + i) * N + mrow] = __uint_as_float(v[i]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemColsT) : "memory");
  }
}

// Fixed-order reduction of the per-CTA partials (and of the [grid][32][N] column-sum slabs).  One
// warp per group of 32 consecutive outputs: lane l owns output 32*g + l; the `grid` partials are cut into
// kRedSeg contiguous segments summed by kRedSeg warps of the block concurrently (coalesced 128-byte
// reads, independent loads in flight) and the segment sums are combined in segment order, so the
// summation tree depends on `grid` only -- deterministic run to run.
constexpr int kRedSeg = 8;
__global__ void __launch_bounds__(32 * kRedSeg) wgrad_reduce_kernel(const float* __restrict__ partials,
                                                                     const float* __restrict__ colsum_part, int grid, int n,
                                                                     int feat, float* __restrict__ out1, int64_t ldo1,
                                                                     float* __restrict__ out2, int64_t ldo2,
                                                                     float* __restrict__ colsum) {
  __shared__ float seg_sum[kRedSeg][32];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;          // output index
  const int per = 2 * feat * n;                  // a CTA's partial: [2][feat][n] (stacked: 2*64 rows = one accumulator)
  const int stride = 2 * kFeat * n;              // slab pitch of the workspace
  const int c0 = (int)((int64_t)grid * seg / kRedSeg), c1 = (int)((int64_t)grid * (seg + 1) / kRedSeg);
  float s = 0.f;
  if (i < per) {
    if (i < feat * n || out2 != nullptr)           // the second half is neither produced nor wanted without A2
      for (int c = c0; c < c1; ++c) s += partials[(int64_t)c * stride + i];
  } else if (i < per + n && colsum != nullptr) {
    const int col = i - per;
    for (int c = c0; c < c1; ++c)
      for (int r = 0; r < kRowsPerChunk; ++r) s += colsum_part[((int64_t)c * kRowsPerChunk + r) * n + col];
  }
  seg_sum[seg][lane] = s;
  __syncthreads();
  if (seg != 0) return;
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < kRedSeg; ++k) tot += seg_sum[k][lane];
  if (i < per) {
    const int d = i / (feat * n), r = (i / n) % feat, col = i % n;
    if (d == 0) out1[(int64_t)r * ldo1 + col] = tot;
    else if (out2 != nullptr) out2[(int64_t)r * ldo2 + col] = tot;
  } else if (i < per + n && colsum != nullptr) {
    colsum[i - per] = tot;
  }
}

}  // namespace tcw

int wgrad_tcgen05_supported(int64_t m, int64_t k1, int64_t k2, int64_t n, uint32_t flags) {
  if (!(flags & MPGNN_F_TF32X3)) return 0;
  if (m < 1 || !(n == 64 || n == 128)) return 0;
  if (k1 == tcw::kFeat / 2 && k2 == 0) return 1;          // one 64-wide operand (head layer): stacked tile, second half zero
  if (k1 == tcw::kFeat && k2 == 0) return n == 128;       // one 128-wide operand (compact hop): swapped roles, N operand 128
  return k1 == k2 && (k1 == tcw::kFeat || k1 == tcw::kFeat / 2);
}

static void wgrad_split(int64_t m, int* grid, int64_t* rows_per_cta) {
  int64_t rpc = align_up(ceil_div(m, kNumSMs), tcw::kRowsPerChunk);
  int64_t g = ceil_div(m, rpc);
  *grid = (int)g;
  *rows_per_cta = rpc;
}

int64_t wgrad_tcgen05_workspace_floats(int64_t m, int64_t n) {
  int grid;
  int64_t rpc;
  wgrad_split(m, &grid, &rpc);
  return (int64_t)grid * (2 * tcw::kFeat * n + tcw::kRowsPerChunk * n);
}

int launch_wgrad_tcgen05(const GemmTnArgs& a, float* ws, cudaStream_t s) {
  MPGNN_REQUIRE(wgrad_tcgen05_supported(a.m, a.k1, a.k2, a.n, MPGNN_F_TF32X3), MPGNN_ENOTSUP,
                "wgrad_tcgen05: unsupported shape");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  MPGNN_REQUIRE(al16(a.a1) && al16(a.b) && a.lda1 % 4 == 0 && a.ldb % 4 == 0 &&
                    (a.k2 == 0 || (al16(a.a2) && a.lda2 % 4 == 0)),
                MPGNN_EINVAL, "wgrad_tcgen05: operands must be 16-byte aligned with strides multiple of 4");
  int grid;
  int64_t rpc;
  wgrad_split(a.m, &grid, &rpc);
  tcw::ParamsW p{};
  p.a1 = a.a1; p.lda1 = a.lda1;
  p.a2 = a.k2 > 0 ? a.a2 : nullptr; p.lda2 = a.lda2;
  p.b = a.b; p.ldb = a.ldb; p.n = (int)a.n;
  p.m = a.m; p.rows_per_cta = rpc;
  p.b_actmask = a.b_actmask; p.b_scale = a.b_scale;
  p.partials = ws;
  p.colsum_part = ws + (int64_t)grid * 2 * tcw::kFeat * a.n;
  const int n_stages = tcw::kStagesPair;
  const size_t smem = (size_t)n_stages * (4 * tcw::kATileBytes + 2 * tcw::kRowsPerChunk * a.n * 4) + (2 * n_stages + 1) * 8 + 16;
  auto launch = [&](auto kernel) -> int {
    // the opt-in limit is per function and process wide: always raise it to the device maximum, so that concurrent
    // launches of the same kernel with different tile sizes (candidate trainers on several host threads) cannot
    // lower it under one another
    MPGNN_REQUIRE(smem <= (size_t)kMaxDynSmem, MPGNN_ENOTSUP, "shared memory request %zu exceeds the device limit", smem);
    MPGNN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    kernel<<<grid, tcw::kThreadsW, smem, s>>>(p);
    MPGNN_LAUNCH_CHECK();
    return MPGNN_OK;
  };
  const bool masked = a.b_actmask != nullptr;
  const bool stacked = a.k1 == tcw::kFeat / 2;
  const bool single = a.k1 == tcw::kFeat && a.k2 == 0;
  if (single) {
    MPGNN_REQUIRE(!masked || al16(a.b_actmask), MPGNN_EINVAL, "wgrad_tcgen05: activation mask must be 16-byte aligned");
    CUtensorMap map_x, map_g;
    MPGNN_PROPAGATE(make_tensor_map(&map_x, a.a1, a.m, tcw::kFeat, a.lda1, tcw::kRowsPerChunk, 32));
    MPGNN_PROPAGATE(make_tensor_map(&map_g, a.b, a.m, tcw::kFeat, a.ldb, tcw::kRowsPerChunk, 32));
    auto launch_t = [&](auto kernel) -> int {
      MPGNN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
      kernel<<<grid, tcw::kThreadsT, tcw::kSmemT, s>>>(p, map_x, map_g);
      MPGNN_LAUNCH_CHECK();
      return MPGNN_OK;
    };
    if (masked) MPGNN_PROPAGATE(launch_t(tcw::wgrad_tma_kernel<true>));
    else MPGNN_PROPAGATE(launch_t(tcw::wgrad_tma_kernel<false>));
  } else
  switch ((a.n == 128 ? 4 : 0) + (masked ? 2 : 0) + (stacked ? 1 : 0)) {
    case 0: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<64, false, false>)); break;
    case 1: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<64, false, true>)); break;
    case 2: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<64, true, false>)); break;
    case 3: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<64, true, true>)); break;
    case 4: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<128, false, false>)); break;
    case 5: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<128, false, true>)); break;
    case 6: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<128, true, false>)); break;
    default: MPGNN_PROPAGATE(launch(tcw::wgrad_tc_kernel<128, true, true>)); break;
  }
  const int feat = (int)a.k1;
  const int total = 2 * feat * (int)a.n + (int)a.n;
  tcw::wgrad_reduce_kernel<<<(unsigned)ceil_div(total, 32), 32 * tcw::kRedSeg, 0, s>>>(
      p.partials, p.colsum_part, grid, (int)a.n, feat, a.out1, a.ldo1, a.out2, a.ldo2, a.out_ones);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

}  // namespace mpgnn
