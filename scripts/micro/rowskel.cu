// micro-benchmark (experiment helper): the "row skeleton" of K2 at C4 -- read ptr[row], ptr[row+1], write a 512-byte row --
// under different row-to-warp mappings.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a rowskel.cu -o rowskel
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int F4 = 32;  // float4 per row (128 floats)
// A: one row per warp iteration, grid stride (the current kernel's mapping)
__global__ void __launch_bounds__(256) kA(const int* __restrict__ ptr, int64_t n, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w; r < n; r += nw) {
    const int d = __ldg(ptr + r + 1) - __ldg(ptr + r);
    const float v = d > 1000000 ? 1.f : 0.f;
    out[r * F4 + lane] = make_float4(v, v, v, v);
  }
}
// B: B rows per warp iteration, one coalesced ptr load
template <int B>
__global__ void __launch_bounds__(256) kB(const int* __restrict__ ptr, int64_t n, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = w * B; r0 < n; r0 += nw * B) {
    const int pv = (lane <= B && r0 + lane <= n) ? __ldg(ptr + r0 + lane) : 0;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const int d = __shfl_sync(0xffffffffu, pv, b + 1) - __shfl_sync(0xffffffffu, pv, b);
      const float v = d > 1000000 ? 1.f : 0.f;
      if (r0 + b < n) out[(r0 + b) * F4 + lane] = make_float4(v, v, v, v);
    }
  }
}
// C: each warp owns a contiguous range of rows
__global__ void __launch_bounds__(256) kC(const int* __restrict__ ptr, int64_t n, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t per = (n + nw - 1) / nw;
  const int64_t lo = w * per, hi = lo + per < n ? lo + per : n;
  for (int64_t r0 = lo; r0 < hi; r0 += 31) {
    const int pv = (r0 + lane <= hi) ? __ldg(ptr + r0 + lane) : 0;
    const int cnt = (int)(hi - r0 < 31 ? hi - r0 : 31);
    for (int b = 0; b < cnt; ++b) {
      const int d = __shfl_sync(0xffffffffu, pv, b + 1) - __shfl_sync(0xffffffffu, pv, b);
      const float v = d > 1000000 ? 1.f : 0.f;
      out[(r0 + b) * F4 + lane] = make_float4(v, v, v, v);
    }
  }
}
// D: thread per 16 bytes, plain grid stride over the output (row = i / 32), ptr via L1
__global__ void __launch_bounds__(256) kD(const int* __restrict__ ptr, int64_t n, float4* __restrict__ out) {
  const int64_t total = n * F4, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i >> 5;
    const int d = __ldg(ptr + r + 1) - __ldg(ptr + r);
    const float v = d > 1000000 ? 1.f : 0.f;
    out[i] = make_float4(v, v, v, v);
  }
}
template <typename K>
float timeit(K k) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) k();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}
int main() {
  const int64_t n = 10000000;
  int* ptr; float4* out;
  cudaMalloc(&ptr, (n + 1) * 4); cudaMemset(ptr, 0, (n + 1) * 4);
  cudaMalloc(&out, n * F4 * 16);
  for (int blocks : {148 * 8, 148 * 16, 148 * 64, 148 * 128}) {
    printf("blocks %6d: A %.3f  B2 %.3f  B4 %.3f  B8 %.3f  C %.3f  D %.3f ms\n", blocks,
           timeit([&] { kA<<<blocks, 256>>>(ptr, n, out); }), timeit([&] { kB<2><<<blocks, 256>>>(ptr, n, out); }),
           timeit([&] { kB<4><<<blocks, 256>>>(ptr, n, out); }), timeit([&] { kB<8><<<blocks, 256>>>(ptr, n, out); }),
           timeit([&] { kC<<<blocks, 256>>>(ptr, n, out); }), timeit([&] { kD<<<blocks, 256>>>(ptr, n, out); }));
  }
  printf("memset: %.3f ms\n", timeit([&] { cudaMemsetAsync(out, 0, n * F4 * 16); }));
  return 0;
}
