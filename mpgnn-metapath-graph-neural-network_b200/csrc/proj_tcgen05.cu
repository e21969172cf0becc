// K3 on the 5th-generation tensor cores (tcgen05 + TMEM).  Placeholder until the kernel
// lands: reports "unsupported" so hop_fwd keeps using the exact-fp32 SIMT projection.
#include "common.cuh"

namespace mpgnn {

int proj_tcgen05_supported(int64_t, int64_t, int64_t, int64_t, uint32_t) { return 0; }

int launch_proj_tcgen05(const GemmRowsArgs&, uint32_t, cudaStream_t) {
  set_error("tcgen05 projection not built");
  return MPGNN_ENOTSUP;
}

}  // namespace mpgnn
