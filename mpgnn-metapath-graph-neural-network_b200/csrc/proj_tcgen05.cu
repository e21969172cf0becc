// K3 -- the weight projection on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   out[M,N] = epi( [A1 | A2][M,K] @ B[K,N] )      fp32 in, fp32 out
//
// fp32 parity mode ("3xTF32"): every fp32 operand is split as v = hi + lo with hi = v rounded to
// the nearest TF32 number and lo = v - hi (exact in fp32), and the product is accumulated in fp32
// in TMEM as  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  (the dropped lo*lo term is < 2^-24 relative),
// which keeps the normalised error of the projection at the 1e-6 level (bar: 1e-5).
//
// Design (one persistent CTA per SM; an opt-in variant pairs CTAs with cta_group::2, see kPair below):
//  * a CTA owns ONE column slice of BN outputs (BN*K*8 bytes of B, hi+lo, resident in shared
//    memory for the whole kernel) and walks 128-row tiles; the CTAs owning the other slices of a
//    row tile run next to it, so the second read of the A tile is an L2 hit.
//  * the A operand is fed to the MMA from TENSOR MEMORY, not shared memory: with three MMAs per
//    K-step the shared-memory operand fetch (A 4 KB + B 2-4 KB per MMA at ~110 B/clk) was the
//    bound, so only the small resident B is read from shared memory now.  One thread streams raw
//    fp32 A chunks (128 rows x 32 K) with TMA (cp.async.bulk.tensor, SWIZZLE_128B, OOB rows zero
//    filled) into a 4-deep shared ring; 16 converter warps in two groups that take alternate chunks
//    (a warp's wait -> LDS -> split -> tcgen05.st -> wait::st -> arrive chain is long, so consecutive
//    chunks must not be serialised behind one warp) read their 32 rows x 16 K-columns conflict-free
//    (the swizzle spreads the 8 rows of a quarter-warp over the 8 bank groups), split hi/lo in
//    registers and tcgen05.st them into a ring of TMEM column stages (lane = row, column = k): as many
//    64-column stages as the two accumulators leave room for (six at BN = 64, four at BN = 128).
//  * one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8; 3 MMAs per K-step, A from
//    TMEM) into one of two TMEM accumulators and tcgen05.commit's the barriers (the stage-free
//    barrier once per pair of chunks: a commit holds the issuing thread for 130-240 cycles).
//  * 8 epilogue warps (two per TMEM lane quarter, each owning alternate 32-column chunks)
//    tcgen05.ld the accumulator and hand it back to the MMA warp at once, apply bias / degree
//    normalisation / relu / dropout in registers (the kernel is specialised on the dropout mode
//    and the degree scaling; the random bits cost 9 integer instructions per 4 elements), put the
//    32x32 block into a swizzled staging tile and let one lane TMA-store it (rows past M are
//    clipped by the tensor map).  It also emits the activation bitmask [out > 0] the backward uses.
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace mpgnn {

namespace tc {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                       // fp32 elements per pipeline stage (4 MMA K-steps)
constexpr int kStages = 6;                        // capacity of the TMEM A-stage ring (64 columns each); p.n_stages are used
constexpr int kRawStages = 4;                     // TMA-filled raw fp32 chunks in flight
constexpr int kProducerWarps = 16;
constexpr int kPairShift = 1;                                     // stage-free barrier shared by 2^kPairShift stages
constexpr int kConvGroups = 2;                                    // converter groups taking alternate chunks
constexpr int kConvGroupWarps = kProducerWarps / kConvGroups;     // 8: 4 TMEM lane quarters x 2 halves of the 32 K-columns
static_assert((kRawStages & (kRawStages - 1)) == 0 && kRawStages % kConvGroups == 0 && kStages % 2 == 0, "ring geometry");
constexpr int kEpiWarps = 8;
// Warp roles by warp id: producers first, epilogue warps next (id % 4 = TMEM lane quarter;
// kProducerWarps is a multiple of 4), the single MMA-issuing warp last.
constexpr int kMmaWarp = kProducerWarps + kEpiWarps;              // 24
constexpr int kTmaWarp = kMmaWarp + 1;                            // 25: one lane issues the TMA loads
constexpr int kThreads = (kTmaWarp + 1) * 32;                     // 832
constexpr int kRawBytes = kTileM * kChunkK * 4;                   // 16 KB per raw chunk (128 rows x 128 B)
constexpr int kACols = 2 * kChunkK;                               // TMEM columns per A stage: hi | lo
constexpr int kEpiCols = 32;                                      // columns per tcgen05.ld
constexpr int kStgBytes = 32 * kEpiCols * 4;                      // per-warp TMA-store staging tile (32 rows x 128 B)
constexpr int kMaxBBytes = 131072;                                // hi + lo of the resident slice

struct Params {
  const float* a1; int64_t lda1; int k1;
  const float* a2; int64_t lda2; int k2;
  const float* b_raw;          // B [K, n] row-major (ld = n): every CTA splits ITS column slice into the hi / lo UMMA images
                               // while it loads it -- no separate image-building launch in front of the GEMM
  int64_t m; int n; int bn; int n_slices;
  const float* bias;
  int relu;
  const int32_t* deg_ptr; int deg_cols;
  int dropout_mode; uint32_t dropout_thr16; float dropout_scale; uint64_t seed; uint64_t offset;
  const uint8_t* mask_bits; const uint64_t* offset_ptr;
  float* out; int64_t ldo;
  int out_split;                               // columns >= out_split are stored through tmap_out2 (0 = off)
  uint32_t* actmask_out;                       // [m][n/32] words, bit j of word c = [out(row, 32c+j) > 0]
  int n_stages;   // TMEM A stages in use (even, <= kStages): 512 columns = acc_bufs * bn + 64 * n_stages
  int acc_bufs;   // TMEM accumulators: 2 (double buffered) or 1 (released right after the epilogue's tcgen05.ld)
  const uint32_t* a_actmask; float a_scale;    // A(r,k) := bit(r,k) ? A(r,k)*a_scale : 0 ([m][K/32] words, k2 == 0)
  // kAdd: out(row, :) += add_src[rank(row), :] before bias/activation, for the rows flagged in add_bits (bit l of word
  // g = row 32g+l has a compact row); rank(row) = add_rank[g] - add_base + popc(add_bits[g] & lanes below l)
  const float* add_src; int64_t ld_add; const uint32_t* add_bits; const uint32_t* add_rank; uint32_t add_base;
#ifdef MPGNN_TC_EXPERIMENT
  int exp;   // profiling build only (scripts/exp_variants.sh): bits switch pipeline stages off; results are WRONG
#endif
};
#ifdef MPGNN_TC_EXPERIMENT
#define TC_EXP(bit) ((p.exp & (bit)) != 0)
#else
#define TC_EXP(bit) false
#endif
#ifdef MPGNN_TC_COUNTERS
// cycle counters of CTA 0 (one lane per role): [role*8 + k]
__device__ unsigned long long g_tc_dbg[64];
#define TC_NOW() clock64()
#define TC_T0() const long long _t0 = clock64()
#define TC_ACC(var) var += clock64() - _t0
#define TC_DUMP(role, k, v) do { if (blockIdx.x == 0 && lane == 0) g_tc_dbg[(role) * 8 + (k)] = (unsigned long long)(v); } while (0)
#else
#define TC_NOW() 0ll
#define TC_T0()
#define TC_ACC(var)
#define TC_DUMP(role, k, v)
#endif

// kDrop: 0 none, 1 seeded, 2 mask bits; kDeg: divide the first deg_cols columns by deg; kMasked: gate A by a_actmask.
// kPair: CTA pairs (clusters of 2, tcgen05 cta_group::2).  The pair owns a 256-row tile and a BN-column slice: each CTA
// streams and converts ITS 128 rows once, keeps HALF of the B slice (BN/2 columns) in shared memory, and the leader's
// MMA thread issues M256 x N(BN) x K8 MMAs that read both halves -- for K = 256 (forward) this is what lets one pass
// over A produce all 128 output columns instead of two CTAs converting the same tile for 64 columns each.
// kAdd: the epilogue adds rows of a compact matrix (the compact hop: y = act(x root + b + scatter(h_c W))).
// kWide (with kAdd): 8 converter warps (one group) and 16 epilogue warps instead of 16 + 8.  The forward's epilogue
// (bias, relu, dropout bits, bitmask, compact addend) executes ~620 instructions per 32x32 block against ~120 for a
// converter's block: with 8 epilogue warps each one ran ~1240 instructions per tile, stalled on latencies 7/8 of the
// time, and the whole pipeline backed up behind them.  Every wide epilogue warp owns 16 columns of two chunks per tile
// (half the latency chain) and a 2 KB staging tile stored with 16-column TMA boxes.
template <int kDrop, bool kDeg, bool kMasked, bool kPair = false, bool kAdd = false, bool kWide = false>
__global__ void __launch_bounds__(kThreads, 1) gemm_rows_tc_kernel(const Params p, const __grid_constant__ CUtensorMap tmap_a1,
                                                                   const __grid_constant__ CUtensorMap tmap_a2,
                                                                   const __grid_constant__ CUtensorMap tmap_out,
                                                                   const __grid_constant__ CUtensorMap tmap_out2) {
  static_assert(!kWide || (kAdd && !kPair && !kDeg && !kMasked), "the wide-epilogue variant exists for the compact forward only");
  constexpr int kProd = kWide ? kConvGroupWarps : kProducerWarps;    // converter warps: 8 or 16
  constexpr int kGroups = kProd / kConvGroupWarps;                   // converter groups taking alternate chunks: 1 or 2
  constexpr int kEpi = kMmaWarp - kProd;                             // epilogue warps: 16 or 8
  constexpr int kStg = kWide ? kStgBytes / 2 : kStgBytes;            // staging tile per epilogue warp
  extern __shared__ __align__(1024) uint8_t smem[];
  const int K = p.k1 + p.k2;
  const int BN = p.bn;                                    // accumulator width
  const int BNB = kPair ? BN / 2 : BN;                    // B columns resident in THIS CTA
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int b_bytes = K * BNB * 4;                        // one of hi / lo
  uint8_t* sm_b_hi = smem;
  uint8_t* sm_b_lo = smem + b_bytes;
  uint8_t* sm_raw = smem + 2 * b_bytes;                   // kRawStages raw chunks, 1024-byte aligned (swizzle)
  uint8_t* sm_stg = sm_raw + kRawStages * kRawBytes;                            // kEpiWarps staging tiles (1024-aligned)
  float* sm_bias = reinterpret_cast<float*>(sm_stg + kEpi * kStg);              // BN floats (slice bias)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_bias + 128);
  // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], raw_full[kRawStages], raw_empty[kRawStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4 + 2 * kRawStages);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + 2);
  const uint32_t bar_rfull = smem_u32(bars + 2 * kStages + 4), bar_rempty = smem_u32(bars + 2 * kStages + 4 + kRawStages);

  // static work split: this CTA owns slice `slice` and row tiles group, group+n_groups, ...
  // (pairs: "unit" = cluster, its tiles are 256 rows of which this CTA takes the half `rank`)
  const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int slice = unit % p.n_slices;
  const int group = unit / p.n_slices;
  const int n_groups = n_units / p.n_slices;
  const int64_t n_tiles = kPair ? (p.m + 2 * kTileM - 1) / (2 * kTileM) : (p.m + kTileM - 1) / kTileM;
  const int my_tiles = (group < n_tiles) ? (int)((n_tiles - group + n_groups - 1) / n_groups) : 0;
  // first row of this CTA's part of tile index t
  auto tile_row0 = [&](int64_t t) -> int64_t { return kPair ? (t * 2 + rank) * kTileM : t * kTileM; };
  const int kch = K / kChunkK;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, kPair ? 2 * kConvGroupWarps : kConvGroupWarps);   // pairs: both CTAs arrive on the leader's
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, kPair ? 2 * kEpi : kEpi);
    }
    for (int r = 0; r < kRawStages; ++r) {
      mbar_init(bar_rfull + 8 * r, 1);                 // one arrive.expect_tx + the TMA's byte count
      mbar_init(bar_rempty + 8 * r, kConvGroupWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {  // TMEM: two fp32 accumulators of BN columns + kStages A stages (hi|lo)
    const uint32_t ncols = 512;
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  {  // resident B slice: read raw (coalesced along n), split hi/lo, store in the UMMA K-major no-swizzle layout --
     // element (n, k) of the slice at float offset (n/8)*SBO + (k/4)*32 + (n%8)*4 + (k%4), SBO = (K/4)*32
    const int img = kPair ? slice * 2 + (int)rank : slice;       // BNB-column slice of B this CTA keeps
    const float* src = p.b_raw + (int64_t)img * BNB;
    float* hi_f = reinterpret_cast<float*>(sm_b_hi);
    float* lo_f = reinterpret_cast<float*>(sm_b_lo);
    const int sbo_f = (K / 4) * 32;
    for (int i = tid; i < K * BNB; i += kThreads) {
      const int kk = i / BNB, nl = i - kk * BNB;
      const float v = __ldg(src + (int64_t)kk * p.n + nl);
      const float hi = tf32_hi(v);
      const int off = (nl >> 3) * sbo_f + (kk >> 2) * 32 + (nl & 7) * 4 + (kk & 3);
      hi_f[off] = hi;
      lo_f[off] = v - hi;
    }
    if (tid < BN) sm_bias[tid] = (p.bias != nullptr) ? __ldg(p.bias + slice * BN + tid) : 0.f;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();        // the peer's barriers and TMEM exist before anything is sent to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + (uint32_t)(p.acc_bufs * BN);   // A stages live after the accumulator(s)
  // barriers the two CTAs of a pair share live in the leader (rank 0)
  const uint32_t bar_full_l = kPair ? mapa_rank(bar_full, 0) : bar_full;
  const uint32_t bar_tempty_l = kPair ? mapa_rank(bar_tempty, 0) : bar_tempty;

  if (warp < kProd) {
    // ================================ A converters (raw smem -> hi/lo -> TMEM) ===========
    // group g = warp/8 takes chunks g, g+2, ...; inside a group warp w owns TMEM lane quarter w&3 (rows
    // 32*(w&3)+lane) and K-columns 16*((w>>2)&1).. of the chunk = four 16-byte pieces of its row in the raw
    // tile.  The tile is laid out by the TMA with the 128-byte swizzle: piece c of row r sits at
    // r*128 + ((c ^ (r&7)) * 16).  Chunk `it` lives in raw stage / TMEM stage it % 4.
    const int grp = warp / kConvGroupWarps;
    const int quarter = warp & 3, colhalf = (warp >> 2) & 1;
    const int row = quarter * 32 + lane;
    int off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) off[j] = row * 128 + (((4 * colhalf + j) ^ (row & 7)) << 4);
    const uint32_t my_taddr = tmem_a + (uint32_t)(colhalf * 16) + ((uint32_t)(quarter * 32) << 16);
    const int total = my_tiles * kch;
    // fused ReLU/dropout backward (dgrad): the operand is g_y gated by the activation bitmask of y;
    // the 32-bit word of (row, chunk) is fetched one of this warp's chunks ahead
    int mc = grp, mtile = 0;
    auto load_mask_word = [&]() -> uint32_t {
      while (mc >= kch) { mc -= kch; ++mtile; }
      const int64_t grow = tile_row0(group + (int64_t)mtile * n_groups) + row;
      const uint32_t w = grow < p.m ? __ldg(p.a_actmask + grow * kch + mc) : 0u;
      mc += kGroups;
      return w;
    };
    uint32_t mw = (kMasked && grp < total) ? load_mask_word() : 0u;
    long long c_rfull = 0, c_empty = 0, c_st = 0, c_total = TC_NOW();
    int s = grp;                    // TMEM stage of chunk `it` (it % n_stages) and the parity of this use of it
    uint32_t ph = 0;
    for (int it = grp; it < total; it += kGroups) {
      const int rs = it & (kRawStages - 1);                    // raw stage and the parity of this use of it
      const uint32_t rph = (uint32_t)(it / kRawStages) & 1u;
      const uint8_t* tile = sm_raw + (size_t)rs * kRawBytes;
      const uint32_t mw_next = (kMasked && it + kGroups < total) ? load_mask_word() : 0u;
      { TC_T0(); mbar_wait(bar_rfull + 8 * rs, rph); TC_ACC(c_rfull); }   // the TMA bytes of this chunk have landed
      if (TC_EXP(64)) __nanosleep(500);
      float vv[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = TC_EXP(2) ? make_float4(1.f, 2.f, 3.f, 4.f) : *reinterpret_cast<const float4*>(tile + off[j]);
        vv[4 * j] = v.x; vv[4 * j + 1] = v.y; vv[4 * j + 2] = v.z; vv[4 * j + 3] = v.w;
      }
      // The refill of this raw stage is an async-proxy write (TMA) and the reads above are generic-proxy
      // loads that may still be in flight when a plain arrive is issued (nothing has consumed their
      // registers yet); the TMA data of the next chunk then lands under them -- seen as a few rows per
      // million picking up 16-byte pieces of the wrong chunk.  So the arrive is made data-dependent on
      // one register of each of the four loads of every lane (cheaper than a proxy fence per chunk).
      uint32_t dep = __float_as_uint(vv[0]) | __float_as_uint(vv[4]) | __float_as_uint(vv[8]) | __float_as_uint(vv[12]);
      dep = __reduce_or_sync(0xFFFFFFFFu, dep);
      if (lane == 0) mbar_arrive_after(bar_rempty + 8 * rs, dep);   // raw stage may be refilled
      if (kMasked) {
        const uint32_t bits = mw >> (colhalf * 16);
#pragma unroll
        for (int e = 0; e < 16; ++e) vv[e] = ((bits >> e) & 1u) ? vv[e] * p.a_scale : 0.f;
        mw = mw_next;
      }
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float h = TC_EXP(2) ? vv[e] : tf32_hi(vv[e]);
        hi[e] = __float_as_uint(h);
        lo[e] = __float_as_uint(TC_EXP(2) ? vv[e] : vv[e] - h);
      }
      // TMEM stage s (hi columns [0,32), lo columns [32,64)) once the MMAs that read it are done
      { TC_T0(); mbar_wait(bar_empty + 8 * (s >> kPairShift), ph ^ 1u); TC_ACC(c_empty); }   // one barrier per pair of stages
      {
        TC_T0();
        tc_fence_after();
        const uint32_t ta = my_taddr + (uint32_t)(s * kACols);
        tmem_st16(ta, hi);
        tmem_st16(ta + kChunkK, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        if (TC_EXP(32)) __nanosleep(500);
          __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster_after(bar_full_l + 8 * s, 0u);
          else mbar_arrive(bar_full + 8 * s);
        }
        TC_ACC(c_st);
      }
      s += kGroups;
      if (s >= p.n_stages) { s -= p.n_stages; ph ^= 1u; }
    }
    if (warp == 0 || warp == kConvGroupWarps) {
      const int role = warp == 0 ? 0 : 1;
      TC_DUMP(role, 0, TC_NOW() - c_total); TC_DUMP(role, 1, c_rfull); TC_DUMP(role, 2, c_empty); TC_DUMP(role, 3, c_st);
      TC_DUMP(role, 4, total);
    }
  } else if (warp < kMmaWarp) {
    // ================================ epilogue =========================================
    const int ew = warp - kProd;                 // 0..kEpi-1
    const int quarter = ew & 3;                  // == warp % 4: the TMEM lanes this warp may read
    const int half = ew >> 2;                    // owns 32-column chunks cc = half, half+2, ...
    // staging tile of this warp in the TMA SWIZZLE_128B layout: 16-byte piece c of row r at r*128 + ((c ^ (r&7))*16)
    const uint32_t stg_tile = smem_u32(sm_stg + ew * kStg);
    const uint32_t stg_row = stg_tile + (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    const int64_t mask_ld = (p.n + 7) / 8;
    const int n_cc = BN / kEpiCols;
    const float relu_floor = p.relu ? 0.f : -INFINITY;
    const uint32_t thr_hi = p.dropout_thr16 << 16;
    uint64_t launch_key = 0;
    if (kDrop == 1) launch_key = dropout_launch_key(p.seed, p.offset + (p.offset_ptr != nullptr ? *p.offset_ptr : 0ull));
    long long e_tfull = 0, e_ld = 0, e_total = TC_NOW();
    for (int ti = 0; ti < my_tiles; ++ti) {
      const int buf = p.acc_bufs == 2 ? (ti & 1) : 0;
      const uint32_t ph = (uint32_t)((p.acc_bufs == 2 ? (ti >> 1) : ti) & 1);
      const int64_t row0 = tile_row0(group + (int64_t)ti * n_groups);
      const int64_t row = row0 + quarter * 32 + lane;       // the TMEM lane this thread reads
      float inv_deg = 1.f;
      if (kDeg && row < p.m) {
        const int d = __ldg(p.deg_ptr + row + 1) - __ldg(p.deg_ptr + row);
        inv_deg = 1.0f / (float)max(d, 1);
      }
      uint64_t row_key = 0;
      if (kDrop == 1) row_key = dropout_row_key(launch_key, (uint64_t)row);
      const float* add_row = nullptr;             // this thread's row of the compact addend, if it has one
      if (kAdd) {
        const int64_t g32 = (row0 + quarter * 32) >> 5;        // row0 and quarter*32 are multiples of 32
        uint32_t bits = 0, rank0 = 0;
        if (row0 + quarter * 32 < p.m) {
          bits = __ldg(p.add_bits + g32);
          rank0 = __ldg(p.add_rank + g32) - p.add_base;
        }
        if ((bits >> lane) & 1u) {
          add_row = p.add_src + (int64_t)(rank0 + __popc(bits & ((1u << lane) - 1u))) * p.ld_add + slice * BN;
          // pull the lines this warp will add into L2 while the tile's MMAs are still running
          if (kWide) {
            for (int u = half; u < 2 * n_cc; u += kEpi / 4)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(add_row + u * 16) : "memory");
          } else {
            for (int cc = half; cc < n_cc; cc += 2)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(add_row + cc * kEpiCols) : "memory");
          }
        }
      }
      if (kWide) {
        // 16 epilogue warps: this one owns the 16-column units u = half, half + 4, ... of its lane quarter (two per tile
        // at BN = 128) and a 32-row x 16-column staging tile in the TMA SWIZZLE_64B layout (piece c of row r at
        // r*64 + ((c ^ ((r >> 1) & 3)) * 16)), stored through the 16-column-box tensor map (tmap_out2)
        const int n_units = 2 * n_cc;
        const uint32_t stg_row16 = stg_tile + (uint32_t)lane * 64u;
        const uint32_t sw16 = (uint32_t)(lane >> 1) & 3u;
        mbar_wait(bar_tfull + 8 * buf, ph);
        tc_fence_after();
        if (half >= n_units) {                           // narrow slices leave some warps without a unit: they still release
          if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
          continue;
        }
        for (int u = half; u < n_units; u += kEpi / 4) {
          const int lcol = u * 16;
          const int col0 = slice * BN + lcol;
          uint32_t v[16];
          tmem_ld16(tmem_base + (uint32_t)(buf * BN + lcol) + ((uint32_t)(quarter * 32) << 16), v);
          if (u + kEpi / 4 >= n_units) {                 // this warp's last read of the accumulator: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_after(bar_tempty + 8 * buf, v[0] | v[15]);
          }
          uint32_t act = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(sm_bias + lcol + 4 * q);
            float x0 = __uint_as_float(v[4 * q + 0]), x1 = __uint_as_float(v[4 * q + 1]);
            float x2 = __uint_as_float(v[4 * q + 2]), x3 = __uint_as_float(v[4 * q + 3]);
            if (add_row != nullptr) {
              const float4 av = __ldg(reinterpret_cast<const float4*>(add_row + lcol) + q);
              x0 += av.x; x1 += av.y; x2 += av.z; x3 += av.w;
            }
            x0 = fmaxf(x0 + bq.x, relu_floor); x1 = fmaxf(x1 + bq.y, relu_floor);
            x2 = fmaxf(x2 + bq.z, relu_floor); x3 = fmaxf(x3 + bq.w, relu_floor);
            if (kDrop == 1) {
              const uint2 w = dropout_block(row_key, (uint32_t)(col0 >> 2) + (uint32_t)q);
              x0 = (w.x << 16) >= thr_hi ? x0 * p.dropout_scale : 0.f;
              x1 = w.x >= thr_hi ? x1 * p.dropout_scale : 0.f;
              x2 = (w.y << 16) >= thr_hi ? x2 * p.dropout_scale : 0.f;
              x3 = w.y >= thr_hi ? x3 * p.dropout_scale : 0.f;
            } else if (kDrop == 2) {
              uint32_t mbits = 0;
              if (row < p.m) {
                const int c = col0 + 4 * q;                     // multiple of 4: the nibble of one byte
                const uint32_t byte = __ldg(p.mask_bits + row * mask_ld + (c >> 3));
                mbits = (c & 4) ? (byte & 0xFu) : (byte >> 4);   // MSB-first: bit 3 = first element
              }
              x0 = (mbits & 8u) ? x0 * p.dropout_scale : 0.f;
              x1 = (mbits & 4u) ? x1 * p.dropout_scale : 0.f;
              x2 = (mbits & 2u) ? x2 * p.dropout_scale : 0.f;
              x3 = (mbits & 1u) ? x3 * p.dropout_scale : 0.f;
            }
            act |= ((x0 > 0.f ? 1u : 0u) | (x1 > 0.f ? 2u : 0u) | (x2 > 0.f ? 4u : 0u) | (x3 > 0.f ? 8u : 0u)) << (4 * q);
            v[4 * q + 0] = __float_as_uint(x0); v[4 * q + 1] = __float_as_uint(x1);
            v[4 * q + 2] = __float_as_uint(x2); v[4 * q + 3] = __float_as_uint(x3);
          }
          if (p.actmask_out != nullptr && row < p.m)     // this unit's 16 bits of the (row, 32-column) word
            reinterpret_cast<uint16_t*>(p.actmask_out)[(row * (int64_t)(p.n >> 5) + (col0 >> 5)) * 2 + ((col0 >> 4) & 1)] =
                (uint16_t)act;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has read the tile
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row16 + (((uint32_t)q ^ sw16) << 4)),
                         "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                         : "memory");
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmap_out2)),
                         "r"(stg_tile), "r"(col0), "r"((int)(row0 + quarter * 32))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        continue;                                        // next tile
      }
      { TC_T0(); mbar_wait(bar_tfull + 8 * buf, ph); TC_ACC(e_tfull); }
      if (TC_EXP(128)) __nanosleep(500);
      tc_fence_after();
      for (int cc = half; cc < n_cc; cc += 2) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (uint32_t)(buf * BN + cc * kEpiCols) + ((uint32_t)(quarter * 32) << 16);
        { TC_T0(); tmem_ld32(taddr, v); TC_ACC(e_ld); }
        const bool last_read = cc + 2 >= n_cc;   // this warp's last read of the accumulator
        if (last_read) {                         // hand the accumulator back before the math
          tc_fence_before();
          __syncwarp();
          // data-dependent on the loaded registers: the MMA warp overwrites the accumulator as soon as the
          // barrier flips, so the LDTM must have delivered (see mbar_arrive_after)
          if (lane == 0) {
            if (kPair) mbar_arrive_cluster_after(bar_tempty_l + 8 * buf, v[0] | v[31]);
            else mbar_arrive_after(bar_tempty + 8 * buf, v[0] | v[31]);
          }
        }
        if (TC_EXP(1)) continue;
        const int lcol0 = cc * kEpiCols;                    // column inside the slice
        const int col0 = slice * BN + lcol0;                // global output column
        const bool scale_deg = kDeg && col0 < p.deg_cols;   // deg_cols is a multiple of 32 (checked on host)
        if (kAdd && add_row != nullptr) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(add_row + lcol0) + q);
            v[4 * q + 0] = __float_as_uint(__uint_as_float(v[4 * q + 0]) + av.x);
            v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + av.y);
            v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + av.z);
            v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + av.w);
          }
        }
        float o[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bq = *reinterpret_cast<const float4*>(sm_bias + lcol0 + 4 * q);
          const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x = __uint_as_float(v[4 * q + j]) + bb[j];
            if (scale_deg) x *= inv_deg;
            o[4 * q + j] = fmaxf(x, relu_floor);
          }
          if (kDrop == 1) {
            const uint2 w = dropout_block(row_key, (uint32_t)(col0 >> 2) + (uint32_t)q);
            o[4 * q + 0] = (w.x << 16) >= thr_hi ? o[4 * q + 0] * p.dropout_scale : 0.f;
            o[4 * q + 1] = w.x >= thr_hi ? o[4 * q + 1] * p.dropout_scale : 0.f;
            o[4 * q + 2] = (w.y << 16) >= thr_hi ? o[4 * q + 2] * p.dropout_scale : 0.f;
            o[4 * q + 3] = w.y >= thr_hi ? o[4 * q + 3] * p.dropout_scale : 0.f;
          } else if (kDrop == 2) {
            uint32_t mbits = 0;
            if (row < p.m) {
              const int c = col0 + 4 * q;                   // multiple of 4: the nibble of one byte
              const uint32_t byte = __ldg(p.mask_bits + row * mask_ld + (c >> 3));
              mbits = (c & 4) ? (byte & 0xFu) : (byte >> 4);   // MSB-first: bit 3 = first element
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) o[4 * q + j] = ((mbits >> (3 - j)) & 1u) ? o[4 * q + j] * p.dropout_scale : 0.f;
          }
        }
        if (p.actmask_out != nullptr) {
          uint32_t act = 0;                                 // [out > 0] of this thread's 32 columns
#pragma unroll
          for (int e = 0; e < 32; ++e) act |= (o[e] > 0.f ? 1u : 0u) << e;
          if (row < p.m) p.actmask_out[row * (int64_t)(p.n >> 5) + (col0 >> 5)] = act;
        }
        // the previous TMA store of this warp must have finished READING the staging tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg_row + (((uint32_t)q ^ sw) << 4)), "f"(o[4 * q]),
                       "f"(o[4 * q + 1]), "f"(o[4 * q + 2]), "f"(o[4 * q + 3])
                       : "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !TC_EXP(16)) {
          const bool second = p.out_split > 0 && col0 >= p.out_split;   // e.g. dgrad: [t | g_z root^T] to two tensors
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(second ? &tmap_out2 : &tmap_out)),
                       "r"(stg_tile), "r"(second ? col0 - p.out_split : col0), "r"((int)(row0 + quarter * 32))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (ew == 0) { TC_DUMP(2, 0, TC_NOW() - e_total); TC_DUMP(2, 1, e_tfull); TC_DUMP(2, 2, e_ld); TC_DUMP(2, 4, my_tiles); }
  } else if (warp == kTmaWarp) {
    // ================================ TMA producer (one lane) ============================
    if (lane == 0) {
      const int total = my_tiles * kch;
      int rs = 0;
      uint32_t rph = 0;
      int tile_i = 0, c = 0;
      long long t_rempty = 0, t_total = TC_NOW();
      for (int it = 0; it < total; ++it) {
        { TC_T0(); mbar_wait(bar_rempty + 8 * rs, rph ^ 1u); TC_ACC(t_rempty); }   // converters are done with this raw stage
        const int kbase = c * kChunkK;
        const bool first = kbase < p.k1;
        const CUtensorMap* map = first ? &tmap_a1 : &tmap_a2;
        const int kcoord = first ? kbase : kbase - p.k1;
        const int64_t row0 = tile_row0(group + (int64_t)tile_i * n_groups);
        const uint32_t dst = smem_u32(sm_raw + (size_t)rs * kRawBytes);
        const uint32_t bar = bar_rfull + 8 * rs;
        if (TC_EXP(8)) {
          mbar_arrive(bar);
          if (++c == kch) { c = 0; ++tile_i; }
          if (++rs == kRawStages) { rs = 0; rph ^= 1u; }
          continue;
        }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kRawBytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(kcoord), "r"((int)row0)
            : "memory");
        if (++c == kch) { c = 0; ++tile_i; }
        if (++rs == kRawStages) { rs = 0; rph ^= 1u; }
      }
      TC_DUMP(3, 0, TC_NOW() - t_total); TC_DUMP(3, 1, t_rempty);
    }
    __syncwarp();
  } else {
    // ================================ MMA issuer ==========================================
    // ONE elected thread runs the whole loop (the per-chunk elect + reconvergence cost ~100 cycles of the
    // serial issue path); it commits the "stage free" barrier once per PAIR of chunks: a tcgen05.commit
    // holds the issuing thread for 130-240 cycles, as long as the MMAs of half a chunk.
    if (rank == 0 && elect_one()) {          // pairs: the leader issues for both CTAs
      const uint32_t idesc = make_idesc(kPair ? 2 * kTileM : kTileM, BN);
      const uint32_t b_sbo = (uint32_t)(K / 4) * 128;
      const uint32_t bh_lo0 = desc_lo(smem_u32(sm_b_hi)), bl_lo0 = desc_lo(smem_u32(sm_b_lo));
      int s = 0;
      uint32_t sph = 0;
      long long m_tempty = 0, m_full = 0, m_issue = 0, m_total = TC_NOW();
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int buf = p.acc_bufs == 2 ? (ti & 1) : 0;
        const uint32_t ph = (uint32_t)((p.acc_bufs == 2 ? (ti >> 1) : ti) & 1);
        { TC_T0(); if (kPair) mbar_wait_cluster(bar_tempty + 8 * buf, ph ^ 1u); else mbar_wait(bar_tempty + 8 * buf, ph ^ 1u); TC_ACC(m_tempty); }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int c = 0; c < kch; ++c) {
          { TC_T0(); if (kPair) mbar_wait_cluster(bar_full + 8 * s, sph); else mbar_wait(bar_full + 8 * s, sph); TC_ACC(m_full); }
          if (TC_EXP(256)) __nanosleep(500);
              tc_fence_after();
          TC_T0();
          const uint32_t ah = tmem_a + (uint32_t)(s * kACols), al = ah + kChunkK;   // TMEM columns of this stage
          const uint32_t boff = (uint32_t)(c * (kChunkK / 4) * 128) >> 4;
#pragma unroll
          for (int j = 0; j < kChunkK / 8; ++j) {
            const uint64_t dbh = desc_make(bh_lo0 + boff + j * 16, b_sbo);
            const uint64_t dbl = desc_make(bl_lo0 + boff + j * 16, b_sbo);
            if (kPair) {
              umma_tf32_ts_pair(d_tmem, ah + j * 8, dbh, idesc, (c | j) != 0 ? 1u : 0u);
              umma_tf32_ts_pair(d_tmem, al + j * 8, dbh, idesc, 1u);
              umma_tf32_ts_pair(d_tmem, ah + j * 8, dbl, idesc, 1u);
              continue;
            }
            umma_tf32_ts(d_tmem, ah + j * 8, dbh, idesc, (c | j) != 0 ? 1u : 0u);
            if (TC_EXP(4)) continue;
            umma_tf32_ts(d_tmem, al + j * 8, dbh, idesc, 1u);
            umma_tf32_ts(d_tmem, ah + j * 8, dbl, idesc, 1u);
          }
          // stages s-1, s reusable once these MMAs have read them; accumulator complete after the tile's last chunk
          if (kPair) {
            if (kPairShift == 0 || (s & 1)) umma_commit_pair(bar_empty + 8 * (s >> kPairShift));
            if (c == kch - 1) umma_commit_pair(bar_tfull + 8 * buf);
          } else {
            if (kPairShift == 0 || (s & 1)) umma_commit(bar_empty + 8 * (s >> kPairShift));
            if (c == kch - 1) umma_commit(bar_tfull + 8 * buf);
          }
          TC_ACC(m_issue);
          if (++s == p.n_stages) { s = 0; sph ^= 1u; }
        }
      }
#ifdef MPGNN_TC_COUNTERS
      if (blockIdx.x == 0) {
        g_tc_dbg[32] = clock64() - m_total; g_tc_dbg[33] = m_tempty; g_tc_dbg[34] = m_full; g_tc_dbg[35] = m_issue;
      }
#endif
    }
    __syncwarp();     // the idle lanes must not reach the teardown barrier (and the TMEM dealloc) ahead of the issuer
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();      // the peer may still be arriving on our barriers / the MMAs reading our smem
  if (warp == kMmaWarp) {
    tc_fence_after();
    const uint32_t ncols = 512;
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

}  // namespace tc

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tensor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int box_cols = tc::kChunkK) {
  static EncodeTiledFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MPGNN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MPGNN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, MPGNN_ECUDA,
                  "proj_tcgen05: cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  // fp32 [rows, cols] row-major with row pitch ld; box = 32 columns (128 bytes, the swizzle span) x box_rows rows
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};      // 32 columns = the 128-byte swizzle span; 16 = 64 bytes
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MPGNN_REQUIRE(r == CUDA_SUCCESS, MPGNN_ECUDA, "proj_tcgen05: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MPGNN_OK;
}

static thread_local int g_tc_cta_cap = 0;
int tc_cta_cap() { return g_tc_cta_cap; }
void set_tc_cta_cap(int cap) { g_tc_cta_cap = cap > 0 ? cap : 0; }

static int pick_bn(int64_t k, int64_t n) {
  if (n % 128 == 0 && k * 128 * 8 <= tc::kMaxBBytes) return 128;
  if (n % 64 == 0 && k * 64 * 8 <= tc::kMaxBBytes) return 64;
  return 0;
}

#ifdef MPGNN_TC_COUNTERS
extern "C" int mpgnn_tc_debug_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, tc::g_tc_dbg, sizeof(unsigned long long) * 64);
}
#endif

int proj_tcgen05_supported(int64_t m, int64_t k1, int64_t k2, int64_t n, uint32_t flags) {
  if (!(flags & MPGNN_F_TF32X3)) return 0;   // the bf16 mode is not built yet
  const int64_t k = k1 + k2;
  if (m < 1 || m >= (1LL << 31) - tc::kTileM || k < tc::kChunkK || k > 256) return 0;
  if (k1 % tc::kChunkK != 0 || k2 % tc::kChunkK != 0) return 0;
  if (n % 64 != 0 || n > 65536) return 0;
  return pick_bn(k, n) != 0;
}

int64_t proj_tcgen05_workspace_floats(int64_t k, int64_t n) { return 2 * k * n; }

int launch_proj_tcgen05_ws(const GemmRowsArgs& a, uint32_t flags, float* b_img, cudaStream_t s) {
  const int64_t k = a.k1 + a.k2;
  MPGNN_REQUIRE(proj_tcgen05_supported(a.m, a.k1, a.k2, a.n, flags), MPGNN_ENOTSUP, "proj_tcgen05: unsupported shape");
  MPGNN_REQUIRE(a.gate == nullptr, MPGNN_ENOTSUP, "proj_tcgen05: gate epilogue not supported");
  MPGNN_REQUIRE(a.a1_actmask == nullptr || a.k2 == 0, MPGNN_ENOTSUP, "proj_tcgen05: operand mask needs k2 == 0");
  MPGNN_REQUIRE(a.actmask_out == nullptr || a.n % 32 == 0, MPGNN_ENOTSUP, "proj_tcgen05: actmask needs n % 32 == 0");
  MPGNN_REQUIRE(a.deg_ptr == nullptr || a.deg_cols % tc::kEpiCols == 0, MPGNN_ENOTSUP,
                "proj_tcgen05: deg_cols must be a multiple of 32");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  MPGNN_REQUIRE(a.out_split == 0 || (a.out_split % tc::kEpiCols == 0 && a.out_split < a.n && a.out2 != nullptr &&
                                     al16(a.out2) && a.ldo2 % 4 == 0),
                MPGNN_EINVAL, "proj_tcgen05: out_split must be a multiple of 32 inside [0, n) with an aligned out2");
  MPGNN_REQUIRE(al16(a.a1) && (a.k2 == 0 || al16(a.a2)) && al16(a.out) && a.lda1 % 4 == 0 &&
                    (a.k2 == 0 || a.lda2 % 4 == 0) && a.ldo % 4 == 0,
                MPGNN_EINVAL, "proj_tcgen05: operands must be 16-byte aligned with strides multiple of 4");
  int bn = pick_bn(k, a.n);
  // CTA pairs when a 128-column B slice does not fit one CTA (K > 128, i.e. the forward's [h | x] operand) but its
  // halves do: the pair converts each A tile once for all 128 columns.  Opt-in (MPGNN_PROJ_PAIR=1): correct and
  // tested, but at C4 it runs at the single-CTA kernel's 3.6 ms -- with the conversions halved the stage loop
  // (converter -> remote arrive -> MMA -> multicast commit -> converter) is what bounds it, see DESIGN.md 4.2.
#ifdef MPGNN_TC_EXPERIMENT      // profiling builds only (scripts/exp_variants.sh): the product library carries no pair kernels
  static const bool pair_ok = getenv("MPGNN_PROJ_PAIR") != nullptr && atoi(getenv("MPGNN_PROJ_PAIR")) != 0;
#else
  const bool pair_ok = false;
#endif
  const bool pair = pair_ok && bn == 64 && a.n % 128 == 0 && k * 64 * 8 <= tc::kMaxBBytes && a.a1_actmask == nullptr &&
                    a.deg_ptr == nullptr;
  const int bnb = pair ? 64 : bn;              // B columns resident per CTA
  if (pair) bn = 128;
  const int n_slices = (int)(a.n / bn);
  (void)b_img; (void)bnb;      // the images are built inside the kernel now; the scratch argument stays for the callers
  tc::Params p{};
  p.a1 = a.a1; p.lda1 = a.lda1; p.k1 = (int)a.k1;
  p.a2 = a.a2; p.lda2 = a.lda2; p.k2 = (int)a.k2;
  p.b_raw = a.b;
  p.m = a.m; p.n = (int)a.n; p.bn = bn; p.n_slices = n_slices;
  p.bias = a.bias; p.relu = a.relu;
  p.deg_ptr = a.deg_ptr; p.deg_cols = a.deg_ptr ? (int)a.deg_cols : 0;
  p.dropout_mode = a.dropout_mode; p.dropout_thr16 = a.dropout_thr16; p.dropout_scale = a.dropout_scale;
  p.seed = a.seed; p.offset = a.offset; p.mask_bits = a.mask_bits; p.offset_ptr = a.offset_ptr;
  p.out = a.out; p.ldo = a.ldo; p.out_split = (int)a.out_split;
  p.actmask_out = a.actmask_out; p.a_actmask = a.a1_actmask; p.a_scale = a.a1_scale;
  p.add_src = a.add_src; p.ld_add = a.ld_add; p.add_bits = a.add_bits; p.add_rank = a.add_rank; p.add_base = a.add_base;
  // TMEM budget (512 columns): two accumulators + as many 64-column A stages as fit: six at BN = 64 (measured:
  // forward 3.81 -> 3.59 ms against four), four at BN = 128 (one accumulator + six stages was slower: 3.39 -> 3.85 ms)
  p.acc_bufs = 2;
  p.n_stages = (512 - p.acc_bufs * bn) / tc::kACols;
  if (p.n_stages > tc::kStages) p.n_stages = tc::kStages;
  p.n_stages &= ~1;
#ifdef MPGNN_TC_EXPERIMENT
  p.exp = getenv("MPGNN_TC_EXP") ? atoi(getenv("MPGNN_TC_EXP")) : 0;
#endif
  const int64_t n_tiles = ceil_div(a.m, pair ? 2 * tc::kTileM : tc::kTileM);
  int64_t grid = n_tiles * n_slices;           // work units: CTAs, or clusters of two
  int64_t max_units = pair ? kNumSMs / 2 : kNumSMs;
  // concurrent callers (candidate trainers of one wave, each on its own stream) may ask for a share of the SMs: a
  // persistent CTA takes a whole SM, so full-width launches of different streams only ever run one after the other.
  // The result does not depend on the grid (every tile is computed the same way whoever owns it).
  const int cap = tc_cta_cap();
  if (cap > 0 && cap < max_units) max_units = cap < n_slices ? n_slices : cap;
  if (grid > max_units) grid = (max_units / n_slices) * n_slices;
  if (pair) grid *= 2;
  const size_t smem = (size_t)2 * k * bnb * 4 + (size_t)tc::kRawStages * tc::kRawBytes +
                      (size_t)tc::kEpiWarps * tc::kStgBytes + 128 * 4 +
                      (2 * tc::kStages + 4 + 2 * tc::kRawStages) * 8 + 16;
  CUtensorMap map1, map2, map_out, map_out2;
  MPGNN_PROPAGATE(make_tensor_map(&map1, a.a1, a.m, a.k1, a.lda1, tc::kTileM));
  if (a.k2 > 0) MPGNN_PROPAGATE(make_tensor_map(&map2, a.a2, a.m, a.k2, a.lda2, tc::kTileM));
  else map2 = map1;
  if (a.out_split > 0) {
    MPGNN_PROPAGATE(make_tensor_map(&map_out, a.out, a.m, a.out_split, a.ldo, 32));
    MPGNN_PROPAGATE(make_tensor_map(&map_out2, a.out2, a.m, a.n - a.out_split, a.ldo2, 32));
  } else {
    MPGNN_PROPAGATE(make_tensor_map(&map_out, a.out, a.m, a.n, a.ldo, 32));
    map_out2 = map_out;
    if (a.add_src != nullptr)       // the wide-epilogue forward stores 32-row x 16-column boxes
      MPGNN_PROPAGATE(make_tensor_map(&map_out2, a.out, a.m, a.n, a.ldo, 32, 16));
  }
  auto launch = [&](auto kernel) -> int {
    // the opt-in limit is per function and process wide: always raise it to the device maximum, so that concurrent
    // launches of the same kernel with different tile sizes (candidate trainers on several host threads) cannot
    // lower it under one another
    MPGNN_REQUIRE(smem <= (size_t)kMaxDynSmem, MPGNN_ENOTSUP, "shared memory request %zu exceeds the device limit", smem);
    MPGNN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    if (pair) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)grid);
      cfg.blockDim = dim3(tc::kThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = s;
      cudaLaunchAttribute attr{};
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      MPGNN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, p, map1, map2, map_out, map_out2));
      count_launch();
      return MPGNN_OK;
    }
    kernel<<<(unsigned)grid, tc::kThreads, smem, s>>>(p, map1, map2, map_out, map_out2);
    MPGNN_LAUNCH_CHECK();
    return MPGNN_OK;
  };
  const bool deg = p.deg_ptr != nullptr && p.deg_cols > 0;
  if (p.add_src != nullptr) {
    MPGNN_REQUIRE(!pair && !deg && p.a_actmask == nullptr && p.add_bits != nullptr && p.add_rank != nullptr &&
                      p.ld_add % 4 == 0 && al16(p.add_src),
                  MPGNN_ENOTSUP, "proj_tcgen05: the compact addend goes with the plain forward epilogue only");
    switch (p.dropout_mode) {
      case 0: return launch(tc::gemm_rows_tc_kernel<0, false, false, false, true, true>);
      case 1: return launch(tc::gemm_rows_tc_kernel<1, false, false, false, true, true>);
      default: return launch(tc::gemm_rows_tc_kernel<2, false, false, false, true, true>);
    }
  }
  if (p.a_actmask != nullptr) {
    MPGNN_REQUIRE(p.dropout_mode == 0, MPGNN_ENOTSUP, "proj_tcgen05: operand mask with a dropout epilogue");
    return deg ? launch(tc::gemm_rows_tc_kernel<0, true, true>) : launch(tc::gemm_rows_tc_kernel<0, false, true>);
  }
#ifdef MPGNN_TC_EXPERIMENT
  if (pair) {
    switch (p.dropout_mode) {
      case 0: return launch(tc::gemm_rows_tc_kernel<0, false, false, true>);
      case 1: return launch(tc::gemm_rows_tc_kernel<1, false, false, true>);
      default: return launch(tc::gemm_rows_tc_kernel<2, false, false, true>);
    }
  }
#endif
  switch (p.dropout_mode * 2 + (deg ? 1 : 0)) {
    case 0: return launch(tc::gemm_rows_tc_kernel<0, false, false>);
    case 1: return launch(tc::gemm_rows_tc_kernel<0, true, false>);
    case 2: return launch(tc::gemm_rows_tc_kernel<1, false, false>);
    case 3: return launch(tc::gemm_rows_tc_kernel<1, true, false>);
    case 4: return launch(tc::gemm_rows_tc_kernel<2, false, false>);
    default: return launch(tc::gemm_rows_tc_kernel<2, true, false>);
  }
}

}  // namespace mpgnn
