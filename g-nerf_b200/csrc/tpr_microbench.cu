// Roofline denominator for the plane gather: how fast can this GPU serve random
// 128-byte lines with the access shape the render kernels use (8 lanes x LDG.128 per line, 12 independent
// lines in flight per thread)?  Run over a 25 MB working set it measures the L2 -> SM gather bandwidth (one
// image's planes are L2 resident), over a set much larger than L2 the DRAM random-line bandwidth.
#include <cuda_runtime.h>
#include <stdint.h>
#include "triplane_b200.h"

namespace tpr {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {      // lowbias32
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int kInFlight>
__global__ void __launch_bounds__(1024) gather_bench_kernel(const float4* __restrict__ buf, uint32_t n_lines, int iters,
                                                           float* __restrict__ sink) {
  const int sub = threadIdx.x & 7;
  const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;      // one 8-lane group per line
  const uint32_t n_grp = (gridDim.x * blockDim.x) >> 3;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t ctr = grp;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    float4 v[kInFlight];
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) {
      const uint32_t line = (uint32_t)(((uint64_t)mix32(ctr) * n_lines) >> 32);
      ctr += n_grp;
      v[k] = __ldg(buf + (size_t)line * 8 + sub);
    }
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
  }
  sink[(blockIdx.x * blockDim.x + threadIdx.x) & 0xffff] = acc.x + acc.y + acc.z + acc.w;
}

}  // namespace tpr

extern "C" int64_t tpr_gather_microbench(const float* buf, int64_t n_lines, int32_t ctas, int32_t iters, float* sink,
                                         void* stream) {
  return tpr_gather_microbench_ex(buf, n_lines, ctas, 512, 12, iters, sink, stream);
}

extern "C" int64_t tpr_gather_microbench_ex(const float* buf, int64_t n_lines, int32_t ctas, int32_t threads,
                                            int32_t in_flight, int32_t iters, float* sink, void* stream) {
  if (!buf || !sink || n_lines <= 0 || n_lines > 0x7fffffff || ctas <= 0 || iters <= 0 || threads < 32 || threads > 1024 ||
      (threads & 31))
    return TPR_E_SHAPE;
  const float4* b = reinterpret_cast<const float4*>(buf);
  cudaStream_t st = (cudaStream_t)stream;
  switch (in_flight) {
    case 4: tpr::gather_bench_kernel<4><<<ctas, threads, 0, st>>>(b, (uint32_t)n_lines, iters, sink); break;
    case 6: tpr::gather_bench_kernel<6><<<ctas, threads, 0, st>>>(b, (uint32_t)n_lines, iters, sink); break;
    case 12: tpr::gather_bench_kernel<12><<<ctas, threads, 0, st>>>(b, (uint32_t)n_lines, iters, sink); break;
    case 24: tpr::gather_bench_kernel<24><<<ctas, threads, 0, st>>>(b, (uint32_t)n_lines, iters, sink); break;
    default: return TPR_E_SHAPE;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int64_t)e - 1000;
  return (int64_t)ctas * (threads / 8) * in_flight * iters;          // lines fetched
}
