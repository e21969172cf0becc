// K4, TMA-fed variant -- weight gradients for the 128-wide layers (A1, A2, B all [M, 128] fp32).
// EXPERIMENTAL, off by default (MPGNN_WGRAD_TMA=1 selects it): correct and parity-tested, but at C4 it runs at
// 4.36 ms against 3.85 ms for the LDG-fed kernel (3.98 ms when the raw tiles are used as the hi operands directly,
// which the MMA's truncation of tf32 operands allows).  Kept because it establishes two facts the next version
// needs: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B delivers exactly the UMMA MN-major tf32 layout, and the bound of this
// kernel family is shared-memory traffic (operand reads + hi/lo writes), not the global loads -- see DESIGN.md 4.2.
//
//   D^T[128 x 256] = g_z^T [h | x]      rows of D^T = output columns n, columns = rows of [g_W ; g_root]
//   colsum[128]    = sum_rows g_z        (g_bias)
//
// Same math and the same per-CTA partial layout as wgrad_tcgen05.cu (3xTF32, both operands MN-major in the
// SWIZZLE_128B_BASE32B layout, fixed-order split-K), but the producers of that kernel -- LDG into registers two
// chunks ahead, then STS -- topped out at 4.8 TB/s on their own and did not overlap well with the MMAs.  Here
//  * one thread streams 16-row chunks of h, x and g_y with TMA (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, which is the
//    UMMA MN-major layout: a box of 32 features x 16 rows = one MN atom column of 4 K-groups) into a 4-deep ring:
//    per chunk 12 boxes of 2 KB, laid out [operand][atom][row][32 floats];
//  * 16 converter warps (warp = node row of the chunk, lane = 16-byte piece of its 512-byte row) read each raw piece
//    once, overwrite it IN PLACE with its tf32-rounded hi part (g_y is gated by the activation bitmask on the way)
//    and write the lo part to the twin tile; they also keep the column sums of g_z;
//  * one thread issues three M128 x N256 x K8 MMAs per K-step (g_z^T as the M operand, [h | x] as one N operand);
//  * 4 warps write the transposed accumulator into the [2][128][N] partial of this CTA at the end.
// Chunk c of CTA b covers rows (c * grid + b) * 16: the grid reads one contiguous window of each operand at a time.
#include <cuda.h>

#include "tc_common.cuh"

namespace mpgnn {

namespace tcw2 {

using namespace tc;

constexpr int kRows = 16;                          // node rows (MMA K) per chunk
constexpr int kFeat = 128;
constexpr int kStages = 4;
constexpr int kConvWarps = 16;                     // == kRows: one node row per warp
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kConvWarps + kEpiWarps;   // 20
constexpr int kTmaWarp = kMmaWarp + 1;             // 21
constexpr int kThreads = (kTmaWarp + 1) * 32;      // 704
constexpr int kBoxBytes = kRows * 128;             // one TMA box: 16 rows x 32 floats = 2 KB
constexpr int kOpBytes = 4 * kBoxBytes;            // one operand (4 MN atoms) = 8 KB
constexpr int kHiBytes = 3 * kOpBytes;             // h | x | g raw/hi tiles = 24 KB
constexpr int kStageBytes = 2 * kHiBytes;          // + the lo twins = 48 KB

struct Params {
  int64_t m;
  const uint32_t* actmask; float scale;            // g_z = bit ? g_y * scale : 0 ([m][4] words) or null
  float* partials;                                 // [grid][2][128][128]
  float* colsum_part;                              // [grid][kRows][128]
  int64_t partial_stride;                          // floats between the partials of consecutive CTAs
  int64_t colsum_stride;
};

__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
// MN-major SWIZZLE_128B_BASE32B descriptor: LBO = byte distance between MN atoms (32 elements), SBO = between groups
// of 4 K-rows.
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 1ull << 61;
  return d;
}

template <bool kMasked>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tma_kernel(const Params p, const __grid_constant__ CUtensorMap tmap_h,
                                                                const __grid_constant__ CUtensorMap tmap_x,
                                                                const __grid_constant__ CUtensorMap tmap_g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  // bars: full[kStages] (TMA bytes landed), conv[kStages] (hi/lo tiles ready), free[kStages] (MMAs done), done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_conv = smem_u32(bars + kStages), bar_free = smem_u32(bars + 2 * kStages);
  const uint32_t bar_done = smem_u32(bars + 3 * kStages);

  const int64_t total_chunks = (p.m + kRows - 1) / kRows;
  const int n_chunks = (int)(total_chunks > blockIdx.x ? (total_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, kConvWarps);
      mbar_init(bar_free + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    const uint32_t ncols = 256;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kConvWarps) {
    // ================================ converters ========================================
    // warp = node row k of the chunk; lane = (atom a = lane/8, 16-byte piece q = lane%8) of that row in each operand
    // tile.  Inside an atom, row k sits at k*128 and its four 32-byte chunks are XOR-swizzled with k%4.
    const int k = warp;
    const int a = lane >> 3, q = lane & 7;
    const uint32_t piece = (uint32_t)(a * kBoxBytes + k * 128 + ((((q >> 1) ^ (k & 3)) << 5) | ((q & 1) << 4)));
    const int col = lane * 4;                                   // first of this thread's 4 feature columns
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    auto split = [&](uint8_t* hi_ptr, float4 v) {
      const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
      *reinterpret_cast<float4*>(hi_ptr) = h;
      *reinterpret_cast<float4*>(hi_ptr + kHiBytes) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    };
    uint32_t mw = 0, mw_next = 0;
    auto mask_word = [&](int c) -> uint32_t {
      const int64_t row = ((int64_t)c * gridDim.x + blockIdx.x) * kRows + k;
      return row < p.m ? __ldg(p.actmask + row * 4 + (col >> 5)) : 0u;
    };
    if (kMasked && n_chunks > 0) mw = mask_word(0);
    for (int it = 0; it < n_chunks; ++it) {
      const int s = it & (kStages - 1);
      const uint32_t ph = (uint32_t)(it / kStages) & 1u;
      uint8_t* st = smem + (size_t)s * kStageBytes;
      if (kMasked && it + 1 < n_chunks) mw_next = mask_word(it + 1);
      mbar_wait(bar_full + 8 * s, ph);
      const float4 vh = *reinterpret_cast<const float4*>(st + piece);
      const float4 vx = *reinterpret_cast<const float4*>(st + kOpBytes + piece);
      float4 vg = *reinterpret_cast<const float4*>(st + 2 * kOpBytes + piece);
      if (kMasked) {
        const uint32_t nib = mw >> (col & 31);
        vg.x = (nib & 1u) ? vg.x * p.scale : 0.f;
        vg.y = (nib & 2u) ? vg.y * p.scale : 0.f;
        vg.z = (nib & 4u) ? vg.z * p.scale : 0.f;
        vg.w = (nib & 8u) ? vg.w * p.scale : 0.f;
        mw = mw_next;
      }
      csum.x += vg.x; csum.y += vg.y; csum.z += vg.z; csum.w += vg.w;
      split(st + piece, vh);
      split(st + kOpBytes + piece, vx);
      split(st + 2 * kOpBytes + piece, vg);
      fence_proxy_async();                       // generic-proxy writes -> the MMA's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_conv + 8 * s);
    }
    *reinterpret_cast<float4*>(p.colsum_part + (int64_t)blockIdx.x * p.colsum_stride + (int64_t)k * kFeat + col) = csum;
    // the shared reduce kernel sums 32 slab rows per CTA: rows 16..31 of this variant are zero
    *reinterpret_cast<float4*>(p.colsum_part + (int64_t)blockIdx.x * p.colsum_stride + (int64_t)(k + kRows) * kFeat + col) =
        make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (warp == kTmaWarp) {
    // ================================ TMA producer ======================================
    if (lane == 0) {
      for (int it = 0; it < n_chunks; ++it) {
        const int s = it & (kStages - 1);
        const uint32_t ph = (uint32_t)(it / kStages) & 1u;
        mbar_wait(bar_free + 8 * s, ph ^ 1u);
        const uint32_t bar = bar_full + 8 * s;
        const uint32_t dst0 = smem_u32(smem + (size_t)s * kStageBytes);
        const int row0 = (int)(((int64_t)it * gridDim.x + blockIdx.x) * kRows);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)kHiBytes) : "memory");
#pragma unroll
        for (int op = 0; op < 3; ++op) {
          const CUtensorMap* map = op == 0 ? &tmap_h : (op == 1 ? &tmap_x : &tmap_g);
#pragma unroll
          for (int at = 0; at < 4; ++at)
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                    dst0 + (uint32_t)(op * kOpBytes + at * kBoxBytes)),
                "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(at * 32), "r"(row0)
                : "memory");
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer ========================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_mn(kFeat, 2 * kFeat);
      const uint32_t s0 = smem_u32(smem);
      for (int it = 0; it < n_chunks; ++it) {
        const int s = it & (kStages - 1);
        const uint32_t ph = (uint32_t)(it / kStages) & 1u;
        mbar_wait(bar_conv + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = s0 + (uint32_t)(s * kStageBytes);
#pragma unroll
        for (int kg = 0; kg < kRows / 8; ++kg) {
          const uint32_t ko = (uint32_t)kg * 1024u;          // 8 K-rows = two 512-byte K-groups
          // [h | x]: 8 MN atoms 2 KB apart (h atoms 0-3, x atoms 4-7); g: 4 atoms
          const uint64_t bh = desc_mn(st + ko, kBoxBytes, 512), bl = desc_mn(st + kHiBytes + ko, kBoxBytes, 512);
          const uint64_t ah = desc_mn(st + 2 * kOpBytes + ko, kBoxBytes, 512);
          const uint64_t al = desc_mn(st + kHiBytes + 2 * kOpBytes + ko, kBoxBytes, 512);
          umma_tf32(tmem_base, ah, bh, idesc, (it | kg) != 0 ? 1u : 0u);
          umma_tf32(tmem_base, al, bh, idesc, 1u);
          umma_tf32(tmem_base, ah, bl, idesc, 1u);
        }
        umma_commit(bar_free + 8 * s);
        if (it == n_chunks - 1) umma_commit(bar_done);
      }
    }
    __syncwarp();
  } else {
    // ================================ final epilogue: TMEM -> per-CTA partial =============
    const int ew = warp - kConvWarps;               // == warp % 4: TMEM lane quarter
    float* part = p.partials + (int64_t)blockIdx.x * p.partial_stride;
    const int n = ew * 32 + lane;                   // output column (feature of g_z)
    if (n_chunks > 0) {
      mbar_wait(bar_done, 0);
      tc_fence_after();
    }
    for (int cc = 0; cc < 2 * kFeat / 32; ++cc) {
      uint32_t v[32];
      if (n_chunks > 0) {
        tmem_ld32(tmem_base + (uint32_t)(cc * 32) + ((uint32_t)(ew * 32) << 16), v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) part[(int64_t)(cc * 32 + j) * kFeat + n] = __uint_as_float(v[j]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    const uint32_t ncols = 256;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

}  // namespace tcw2

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_mn_tensor_map(CUtensorMap* map, const float* base, int64_t rows, int64_t ld) {
  static EncodeTiledFn2 encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MPGNN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MPGNN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, MPGNN_ECUDA,
                  "wgrad_tma: cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn2>(fn);
  }
  // fp32 [rows, 128] with row pitch ld; box = 32 features x 16 rows, 32-byte-atom 128-byte swizzle = the layout the
  // MMA reads an MN-major tf32 operand in
  const cuuint64_t gdim[2] = {(cuuint64_t)tcw2::kFeat, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)tcw2::kRows};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MPGNN_REQUIRE(r == CUDA_SUCCESS, MPGNN_ECUDA, "wgrad_tma: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MPGNN_OK;
}

int wgrad_tma_supported(int64_t m, int64_t k1, int64_t k2, int64_t n) {
  return m >= 1 && m < (1LL << 31) - 64 && k1 == tcw2::kFeat && k2 == tcw2::kFeat && n == tcw2::kFeat;
}

// partials: [grid][2][128][128] at `partials` (CTA pitch partial_stride floats), column sums [grid][16][128]
int launch_wgrad_tma(const GemmTnArgs& a, int grid, float* partials, int64_t partial_stride, float* colsum_part,
                     int64_t colsum_stride, cudaStream_t s) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  MPGNN_REQUIRE(al16(a.a1) && al16(a.a2) && al16(a.b) && a.lda1 % 4 == 0 && a.lda2 % 4 == 0 && a.ldb % 4 == 0,
                MPGNN_EINVAL, "wgrad_tma: operands must be 16-byte aligned with strides multiple of 4");
  tcw2::Params p{};
  p.m = a.m;
  p.actmask = a.b_actmask; p.scale = a.b_scale;
  p.partials = partials; p.partial_stride = partial_stride;
  p.colsum_part = colsum_part; p.colsum_stride = colsum_stride;
  CUtensorMap mh, mx, mg;
  MPGNN_PROPAGATE(make_mn_tensor_map(&mh, a.a1, a.m, a.lda1));
  MPGNN_PROPAGATE(make_mn_tensor_map(&mx, a.a2, a.m, a.lda2));
  MPGNN_PROPAGATE(make_mn_tensor_map(&mg, a.b, a.m, a.ldb));
  const size_t smem = (size_t)tcw2::kStages * tcw2::kStageBytes + (3 * tcw2::kStages + 1) * 8 + 16;
  auto launch = [&](auto kernel) -> int {
    // the opt-in limit is per function and process wide: always raise it to the device maximum, so that concurrent
    // launches of the same kernel with different tile sizes (candidate trainers on several host threads) cannot
    // lower it under one another
    MPGNN_REQUIRE(smem <= (size_t)kMaxDynSmem, MPGNN_ENOTSUP, "shared memory request %zu exceeds the device limit", smem);
    MPGNN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    kernel<<<grid, tcw2::kThreads, smem, s>>>(p, mh, mx, mg);
    MPGNN_LAUNCH_CHECK();
    return MPGNN_OK;
  };
  return a.b_actmask != nullptr ? launch(tcw2::wgrad_tma_kernel<true>) : launch(tcw2::wgrad_tma_kernel<false>);
}

}  // namespace mpgnn
