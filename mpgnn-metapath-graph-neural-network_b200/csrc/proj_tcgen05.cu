// K3 -- the weight projection on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   out[M,N] = epi( [A1 | A2][M,K] @ B[K,N] )      fp32 in, fp32 out
//
// fp32 parity mode ("3xTF32"): every fp32 operand is split as v = hi + lo with
// hi = v with the low 13 mantissa bits cleared (exactly a TF32 number) and lo = v - hi
// (exact in fp32), and the product is accumulated in fp32 in TMEM as
//   A_hi*B_hi + A_lo*B_hi + A_hi*B_lo            (the dropped lo*lo term is < 2^-22 relative)
// which keeps the normalised error of the projection at the 1e-6 level (1e-5 bar).
//
// Design (one persistent CTA per SM, no cluster):
//  * a CTA owns ONE column slice of BN outputs (BN*K*8 bytes of B, hi+lo, stay resident in
//    shared memory for the whole kernel) and walks 128-row tiles; the CTAs that own the other
//    slices of the same row tile run next to it, so the second read of the A tile hits L2.
//  * 16 producer warps stream A: 128-bit global loads issued 4 K-chunks ahead (register
//    prefetch hides HBM latency), hi/lo split, conflict-free stores into a 2-stage ring of
//    K-major no-swizzle UMMA tiles; generic->async proxy fence; mbarrier arrive.
//  * one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) -- 3 MMAs per
//    K-step -- into one of two TMEM accumulators and tcgen05.commit's the stage/accumulator
//    barriers.
//  * 4 epilogue warps tcgen05.ld their 32 TMEM lanes, apply bias / degree normalisation /
//    relu / dropout in registers, transpose through a padded shared staging tile and write
//    128-byte row segments.
#include "common.cuh"

namespace mpgnn {

namespace tc {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                       // fp32 elements per pipeline stage (4 MMA K-steps)
constexpr int kStages = 2;
constexpr int kPrefetch = 4;                      // chunks in flight in producer registers
constexpr int kEpiWarps = 4;
constexpr int kProducerWarps = 16;
constexpr int kThreads = (kEpiWarps + 1 + kProducerWarps) * 32;   // 672
constexpr int kProducerThreads = kProducerWarps * 32;             // 512
constexpr int kStageBytes = kTileM * kChunkK * 4;                 // 16 KB per hi or lo
constexpr int kEpiCols = 32;
constexpr int kEpiLd = kEpiCols + 4;                              // padded staging row (floats)
constexpr int kMaxBBytes = 131072;                                // hi + lo of the resident slice

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// core matrix = 8 rows x 16 bytes stored as 128 contiguous bytes; `lbo` = byte distance between
// the two K-adjacent core matrices of one MMA, `sbo` = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, tf32 x tf32, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = v rounded to nearest TF32 (low 13 mantissa bits zero afterwards), so |v - hi| <= 2^-12 |v| with a
// random sign: the dropped lo*lo term stays below 2^-24 relative and does not accumulate a bias.
__device__ __forceinline__ float tf32_hi(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

struct Params {
  const float* a1; int64_t lda1; int k1;
  const float* a2; int64_t lda2; int k2;
  const float* b_img;          // [n_slices][2][K*BN] floats: hi image then lo image, UMMA K-major layout
  int64_t m; int n; int bn; int n_slices;
  const float* bias;
  int relu;
  const int32_t* deg_ptr; int deg_cols;
  int dropout_mode; float dropout_p; float dropout_scale; uint64_t seed; uint64_t offset; const uint8_t* mask_bits;
  float* out; int64_t ldo;
};

// Re-lays B[K,N] (row-major) into per-slice UMMA images: element (n,k) of a slice lives at byte
// (n/8)*SBO + (k/4)*128 + (n%8)*16 + (k%4)*4 with SBO = (K/4)*128, split into hi and lo.
__global__ void prep_b_images_kernel(const float* __restrict__ b, int k, int n, int bn, float* __restrict__ img) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)k * n) return;
  const int kk = (int)(i / n), nn = (int)(i % n);
  const int slice = nn / bn, nl = nn % bn;
  const float v = b[i];
  const float hi = tf32_hi(v);
  const int64_t sbo_f = (int64_t)(k / 4) * 32;   // floats
  const int64_t off = (int64_t)(nl / 8) * sbo_f + (int64_t)(kk / 4) * 32 + (nl % 8) * 4 + (kk % 4);
  float* base = img + (int64_t)slice * 2 * k * bn;
  base[off] = hi;
  base[(int64_t)k * bn + off] = v - hi;
}

__global__ void __launch_bounds__(kThreads, 1) gemm_rows_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int K = p.k1 + p.k2;
  const int BN = p.bn;
  const int b_bytes = K * BN * 4;                         // one of hi / lo
  uint8_t* sm_b_hi = smem;
  uint8_t* sm_b_lo = smem + b_bytes;
  uint8_t* sm_a = smem + 2 * b_bytes;                     // kStages x {hi, lo}
  float* sm_epi = reinterpret_cast<float*>(sm_a + kStages * 2 * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm_epi) + kEpiWarps * 32 * kEpiLd * 4);
  // bars: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + 2);

  // static work split: this CTA owns slice `slice` and row tiles group, group+n_groups, ...
  const int slice = blockIdx.x % p.n_slices;
  const int group = blockIdx.x / p.n_slices;
  const int n_groups = gridDim.x / p.n_slices;
  const int64_t n_tiles = (p.m + kTileM - 1) / kTileM;
  const int64_t my_tiles = (group < n_tiles) ? (n_tiles - group + n_groups - 1) / n_groups : 0;
  const int kch = K / kChunkK;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, kProducerWarps);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tfull + 8 * b, 1);
      mbar_init(bar_tempty + 8 * b, kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps) {  // TMEM: two fp32 accumulators of BN columns
    const uint32_t ncols = 2 * BN;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // resident B slice (already in UMMA layout in global memory): straight 16-byte copies
    const uint4* src = reinterpret_cast<const uint4*>(p.b_img + (int64_t)slice * 2 * K * BN);
    uint4* dst = reinterpret_cast<uint4*>(sm_b_hi);
    const int n16 = 2 * b_bytes / 16;
    for (int i = tid; i < n16; i += kThreads) dst[i] = __ldg(src + i);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kEpiWarps) {
    // ================================ epilogue =========================================
    float* stg = sm_epi + warp * 32 * kEpiLd;
    const int64_t mask_ld = (p.n + 7) / 8;
    for (int64_t ti = 0; ti < my_tiles; ++ti) {
      const int buf = (int)(ti & 1);
      const uint32_t ph = (uint32_t)((ti >> 1) & 1);
      const int64_t row0 = (group + ti * n_groups) * (int64_t)kTileM;
      const int64_t row = row0 + warp * 32 + lane;          // the TMEM lane this thread reads
      mbar_wait(bar_tfull + 8 * buf, ph);
      tc_fence_after();
      float inv_den = 1.f;
      if (p.deg_ptr != nullptr && row < p.m) {
        const int d = __ldg(p.deg_ptr + row + 1) - __ldg(p.deg_ptr + row);
        inv_den = (float)max(d, 1);
      }
      for (int cc = 0; cc < BN / kEpiCols; ++cc) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (uint32_t)(buf * BN + cc * kEpiCols) + ((uint32_t)(warp * 32) << 16);
        tmem_ld32(taddr, v);
        const int col0 = slice * BN + cc * kEpiCols;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float o[4];
          uint4 rnd = make_uint4(0, 0, 0, 0);
          if (p.dropout_mode == 1) {
            const uint64_t blk = ((uint64_t)row * (uint64_t)p.n + (uint64_t)(col0 + 4 * q)) >> 2;
            rnd = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)p.offset,
                                           (uint32_t)(p.offset >> 32)),
                                make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
          }
          const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = col0 + 4 * q + j;
            float x = __uint_as_float(v[4 * q + j]);
            if (p.bias != nullptr) x += __ldg(p.bias + col);
            if (col < p.deg_cols) x = x / inv_den;
            if (p.relu) x = fmaxf(x, 0.f);
            if (p.dropout_mode == 1) {
              x = ((float)(rr[j] >> 8) * (1.0f / 16777216.0f) >= p.dropout_p) ? x * p.dropout_scale : 0.f;
            } else if (p.dropout_mode == 2 && row < p.m) {
              const uint8_t byte = __ldg(p.mask_bits + row * mask_ld + (col >> 3));
              x = ((byte >> (7 - (col & 7))) & 1) ? x * p.dropout_scale : 0.f;
            }
            o[j] = x;
          }
          *reinterpret_cast<float4*>(stg + lane * kEpiLd + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        // transposed read: 8 lanes cover one 128-byte row segment, 4 rows per instruction
#pragma unroll
        for (int r4 = 0; r4 < 32; r4 += 4) {
          const int rl = r4 + (lane >> 3);
          const int64_t grow = row0 + warp * 32 + rl;
          const float4 val = *reinterpret_cast<const float4*>(stg + rl * kEpiLd + 4 * (lane & 7));
          if (grow < p.m) *reinterpret_cast<float4*>(p.out + grow * p.ldo + col0 + 4 * (lane & 7)) = val;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    }
  } else if (warp == kEpiWarps) {
    // ================================ MMA issuer ========================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kTileM, BN);
      const uint32_t a_sbo = (kChunkK / 4) * 128, b_sbo = (uint32_t)(K / 4) * 128;
      const uint32_t sa = smem_u32(sm_a), sbh = smem_u32(sm_b_hi), sbl = smem_u32(sm_b_lo);
      int64_t it = 0;
      for (int64_t ti = 0; ti < my_tiles; ++ti) {
        const int buf = (int)(ti & 1);
        const uint32_t ph = (uint32_t)((ti >> 1) & 1);
        mbar_wait(bar_tempty + 8 * buf, ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int c = 0; c < kch; ++c, ++it) {
          const int s = (int)(it % kStages);
          const uint32_t sph = (uint32_t)((it / kStages) & 1);
          mbar_wait(bar_full + 8 * s, sph);
          tc_fence_after();
          const uint32_t a_hi = sa + (uint32_t)s * 2 * kStageBytes, a_lo = a_hi + kStageBytes;
#pragma unroll
          for (int j = 0; j < kChunkK / 8; ++j) {
            const uint64_t dah = make_desc(a_hi + j * 256, 128, a_sbo);
            const uint64_t dal = make_desc(a_lo + j * 256, 128, a_sbo);
            const uint32_t boff = (uint32_t)(c * (kChunkK / 4) + j * 2) * 128;
            const uint64_t dbh = make_desc(sbh + boff, 128, b_sbo);
            const uint64_t dbl = make_desc(sbl + boff, 128, b_sbo);
            umma_tf32(d_tmem, dah, dbh, idesc, (c | j) != 0 ? 1u : 0u);
            umma_tf32(d_tmem, dal, dbh, idesc, 1u);
            umma_tf32(d_tmem, dah, dbl, idesc, 1u);
          }
          umma_commit(bar_empty + 8 * s);       // stage reusable once these MMAs have read it
        }
        umma_commit(bar_tfull + 8 * buf);       // accumulator complete
      }
    }
  } else {
    // ================================ A producers =======================================
    const int pt = tid - (kEpiWarps + 1) * 32;     // 0..511
    // two 16-byte units per thread and chunk; lane -> (row-in-core r, k-core c4) keeps the
    // shared-memory stores conflict free and the global loads sector aligned
    int u_row[2], u_core[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = pt + kProducerThreads * u;   // 0..1023
      const int r = idx & 7, c4 = (idx >> 3) & 3, q = idx >> 5;
      u_row[u] = (q & 15) * 8 + r;
      u_core[u] = (q >> 4) * 4 + c4;
    }
    const int64_t total = my_tiles * kch;
    float4 buf[kPrefetch][2];
    auto issue = [&](int64_t it, float4 (&dst)[2]) {
      const int64_t ti = it / kch;
      const int c = (int)(it % kch);
      const int64_t row0 = (group + ti * n_groups) * (int64_t)kTileM;
      const int kbase = c * kChunkK;
      const float* src;
      int64_t ld;
      int koff;
      if (kbase < p.k1) { src = p.a1; ld = p.lda1; koff = kbase; }
      else { src = p.a2; ld = p.lda2; koff = kbase - p.k1; }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t row = row0 + u_row[u];
        dst[u] = (row < p.m) ? __ldg(reinterpret_cast<const float4*>(src + row * ld + koff + u_core[u] * 4))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store = [&](int s, const float4 (&src)[2]) {
      uint8_t* hi_base = sm_a + (size_t)s * 2 * kStageBytes;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 v = src[u];
        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        const int off = (u_row[u] >> 3) * ((kChunkK / 4) * 128) + u_core[u] * 128 + (u_row[u] & 7) * 16;
        *reinterpret_cast<float4*>(hi_base + off) = h;
        *reinterpret_cast<float4*>(hi_base + kStageBytes + off) = l;
      }
    };
#pragma unroll
    for (int j = 0; j < kPrefetch; ++j)
      if (j < total) issue(j, buf[j]);
    for (int64_t it0 = 0; it0 < total; it0 += kPrefetch) {
#pragma unroll
      for (int j = 0; j < kPrefetch; ++j) {
        const int64_t it = it0 + j;
        if (it < total) {
          const int s = (int)(it % kStages);
          const uint32_t sph = (uint32_t)((it / kStages) & 1);
          mbar_wait(bar_empty + 8 * s, sph ^ 1u);
          store(s, buf[j]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full + 8 * s);
          if (it + kPrefetch < total) issue(it + kPrefetch, buf[j]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) {
    tc_fence_after();
    const uint32_t ncols = 2 * BN;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

}  // namespace tc

static int pick_bn(int64_t k, int64_t n) {
  if (n % 128 == 0 && k * 128 * 8 <= tc::kMaxBBytes) return 128;
  if (n % 64 == 0 && k * 64 * 8 <= tc::kMaxBBytes) return 64;
  return 0;
}

int proj_tcgen05_supported(int64_t m, int64_t k1, int64_t k2, int64_t n, uint32_t flags) {
  if (!(flags & MPGNN_F_TF32X3)) return 0;   // the bf16 mode is not built yet
  const int64_t k = k1 + k2;
  if (m < 1 || k < tc::kChunkK || k > 256) return 0;
  if (k1 % tc::kChunkK != 0 || k2 % tc::kChunkK != 0) return 0;
  if (n % 64 != 0 || n > 65536) return 0;
  return pick_bn(k, n) != 0;
}

int64_t proj_tcgen05_workspace_floats(int64_t k, int64_t n) { return 2 * k * n; }

int launch_proj_tcgen05_ws(const GemmRowsArgs& a, uint32_t flags, float* b_img, cudaStream_t s) {
  const int64_t k = a.k1 + a.k2;
  MPGNN_REQUIRE(proj_tcgen05_supported(a.m, a.k1, a.k2, a.n, flags), MPGNN_ENOTSUP, "proj_tcgen05: unsupported shape");
  MPGNN_REQUIRE(a.gate == nullptr, MPGNN_ENOTSUP, "proj_tcgen05: gate epilogue not supported");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  MPGNN_REQUIRE(al16(a.a1) && (a.k2 == 0 || al16(a.a2)) && al16(a.out) && a.lda1 % 4 == 0 &&
                    (a.k2 == 0 || a.lda2 % 4 == 0) && a.ldo % 4 == 0,
                MPGNN_EINVAL, "proj_tcgen05: operands must be 16-byte aligned with strides multiple of 4");
  const int bn = pick_bn(k, a.n);
  const int n_slices = (int)(a.n / bn);
  tc::prep_b_images_kernel<<<(unsigned)ceil_div(k * a.n, 256), 256, 0, s>>>(a.b, (int)k, (int)a.n, bn, b_img);
  MPGNN_LAUNCH_CHECK();
  tc::Params p{};
  p.a1 = a.a1; p.lda1 = a.lda1; p.k1 = (int)a.k1;
  p.a2 = a.a2; p.lda2 = a.lda2; p.k2 = (int)a.k2;
  p.b_img = b_img;
  p.m = a.m; p.n = (int)a.n; p.bn = bn; p.n_slices = n_slices;
  p.bias = a.bias; p.relu = a.relu;
  p.deg_ptr = a.deg_ptr; p.deg_cols = (int)a.deg_cols;
  p.dropout_mode = a.dropout_mode; p.dropout_p = a.dropout_p; p.dropout_scale = a.dropout_scale;
  p.seed = a.seed; p.offset = a.offset; p.mask_bits = a.mask_bits;
  p.out = a.out; p.ldo = a.ldo;
  const int64_t n_tiles = ceil_div(a.m, tc::kTileM);
  int64_t grid = n_tiles * n_slices;
  if (grid > kNumSMs) grid = (kNumSMs / n_slices) * n_slices;
  const size_t smem = (size_t)2 * k * bn * 4 + (size_t)tc::kStages * 2 * tc::kStageBytes +
                      (size_t)tc::kEpiWarps * 32 * tc::kEpiLd * 4 + (2 * tc::kStages + 4) * 8 + 16;
  MPGNN_CUDA_CHECK(cudaFuncSetAttribute(tc::gemm_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc::gemm_rows_tc_kernel<<<(unsigned)grid, tc::kThreads, smem, s>>>(p);
  MPGNN_LAUNCH_CHECK();
  return MPGNN_OK;
}

}  // namespace mpgnn
