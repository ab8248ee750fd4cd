// Roofline denominator for the plane gather: how fast can this GPU serve random
// 128-byte lines with the access shape the render kernels use (8 lanes x LDG.128 per line, 12 independent
// lines in flight per thread)?  Run over a 25 MB working set it measures the L2 -> SM gather bandwidth (one
// image's planes are L2 resident), over a set much larger than L2 the DRAM random-line bandwidth.
#include <cuda_runtime.h>
#include <stdint.h>
#include "triplane_b200.h"
#include "triplane_b200_bench.h"

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

// ---------------------------------------------------------------------------------------------------------
// Variants of the gather shape (what should the render kernel's gather warps look like?): 128-bit loads with eight lanes
// per line or 256-bit loads (LDG.E.ENL2.256, sm_100) with four lanes per line; a burst of kInFlight loads consumed together,
// or software-pipelined (the next burst is issued before the previous one is consumed: between kInFlight and 2*kInFlight
// loads in flight per thread at any time).
// ---------------------------------------------------------------------------------------------------------
namespace tpr {
struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ldg256(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
  return r;
}
template <int kVec, int kInFlight, bool kPipe>
__global__ void __launch_bounds__(1024) gather_bench2_kernel(const float* __restrict__ buf, uint32_t n_lines, int iters,
                                                            float* __restrict__ sink) {
  constexpr int kLanes = 32 / kVec;                         // lanes per 128-byte line
  const int sub = threadIdx.x % kLanes;
  const uint32_t grp = (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
  const uint32_t n_grp = (gridDim.x * blockDim.x) / kLanes;
  float acc = 0.f;
  uint32_t ctr = grp;
  float cur[kInFlight][kVec], nxt[kInFlight][kVec];
  auto issue = [&](float (&dst)[kInFlight][kVec]) {
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) {
      const uint32_t line = (uint32_t)(((uint64_t)mix32(ctr) * n_lines) >> 32);
      ctr += n_grp;
      const float* p = buf + (size_t)line * 32 + sub * kVec;
      if (kVec == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        dst[k][0] = v.x; dst[k][1] = v.y; dst[k][2] = v.z; dst[k][3] = v.w;
      } else {
        const F8 v = ldg256(p);
#pragma unroll
        for (int c = 0; c < kVec; ++c) dst[k][c] = v.v[c];
      }
    }
  };
  auto consume = [&](float (&src)[kInFlight][kVec]) {
#pragma unroll
    for (int k = 0; k < kInFlight; ++k)
#pragma unroll
      for (int c = 0; c < kVec; ++c) acc += src[k][c];
  };
  if (!kPipe) {
#pragma unroll 1
    for (int it = 0; it < iters; ++it) { issue(cur); consume(cur); }
  } else {
    issue(cur);
#pragma unroll 1
    for (int it = 0; it + 2 <= iters; it += 2) { issue(nxt); consume(cur); issue(cur); consume(nxt); }
    consume(cur);
  }
  sink[(blockIdx.x * blockDim.x + threadIdx.x) & 0xffff] = acc;
}
}  // namespace tpr

extern "C" int64_t tpr_gather_microbench_v2(const float* buf, int64_t n_lines, int32_t ctas, int32_t threads, int32_t vec_floats,
                                            int32_t in_flight, int32_t pipelined, int32_t iters, float* sink, void* stream) {
  if (!buf || !sink || n_lines <= 0 || n_lines > 0x7fffffff || ctas <= 0 || iters <= 0 || (iters & 1) || threads < 32 ||
      threads > 1024 || (threads & 31))
    return TPR_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t nl = (uint32_t)n_lines;
#define TPR_GB2(V, F, P) tpr::gather_bench2_kernel<V, F, P><<<ctas, threads, 0, st>>>(buf, nl, iters, sink)
  const int key = vec_floats * 1000 + in_flight * 10 + (pipelined ? 1 : 0);
  switch (key) {
    case 4020: TPR_GB2(4, 2, false); break;  case 4021: TPR_GB2(4, 2, true); break;
    case 4040: TPR_GB2(4, 4, false); break;  case 4041: TPR_GB2(4, 4, true); break;
    case 4060: TPR_GB2(4, 6, false); break;  case 4080: TPR_GB2(4, 8, false); break;
    case 8010: TPR_GB2(8, 1, false); break;  case 8011: TPR_GB2(8, 1, true); break;
    case 8020: TPR_GB2(8, 2, false); break;  case 8021: TPR_GB2(8, 2, true); break;
    case 8040: TPR_GB2(8, 4, false); break;  case 8041: TPR_GB2(8, 4, true); break;
    default: return TPR_E_SHAPE;
  }
#undef TPR_GB2
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int64_t)e - 1000;
  return (int64_t)ctas * (threads / (32 / vec_floats)) * in_flight * iters;          // lines fetched
}

// ---------------------------------------------------------------------------------------------------------
// The scatter shape of the backward pass: eight lanes add 16 bytes each (red.global.add.v4.f32) to one random
// 128-byte line, `per_iter` lines per thread group and round.  lines * 128 B / time is the ceiling of the plane-gradient
// scatter (12 such lines per sample).
// ---------------------------------------------------------------------------------------------------------
namespace tpr {
__global__ void scatter_bench_kernel(float* __restrict__ buf, uint32_t n_lines, int per_iter, int iters) {
  const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
  uint32_t state = group * 2654435761u + 12345u;
  const float4 v = make_float4(1.0f, 0.5f, 0.25f, 0.125f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int k = 0; k < per_iter; ++k) {
      state = state * 1664525u + 1013904223u;
      const uint32_t line = (uint32_t)(((uint64_t)(state ^ (state >> 15)) * n_lines) >> 32);
      atomicAdd(reinterpret_cast<float4*>(buf + (size_t)line * 32) + sub, v);
    }
  }
}
}  // namespace tpr

extern "C" int64_t tpr_scatter_microbench(float* buf, int64_t n_lines, int32_t ctas, int32_t threads, int32_t per_iter, int32_t iters,
                                          void* stream) {
  if (!buf || n_lines <= 0 || n_lines > 0x7fffffff || ctas <= 0 || iters <= 0 || per_iter <= 0 || threads < 32 || threads > 1024 ||
      (threads & 31))
    return TPR_E_SHAPE;
  tpr::scatter_bench_kernel<<<ctas, threads, 0, (cudaStream_t)stream>>>(buf, (uint32_t)n_lines, per_iter, iters);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -(int64_t)e - 1000;
  return (int64_t)ctas * (threads / 8) * per_iter * iters;                             // lines added to
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05.mma cost for the small shapes of the decoder: `count` back-to-back MMAs of M = 128, N = n, one K step
// each (tf32: K = 8, bf16: K = 16), A from shared memory (SS) or TMEM (TS), issued by one thread.  Reports
// cycles from the first issue to the commit's arrival, and the issue-only cycles.
// ---------------------------------------------------------------------------------------------------------
#include "tpr_tc.cuh"
namespace tpr {
using namespace tc;
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int n, int bf16, int ts, int count, int layout, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* a_tile = reinterpret_cast<float*>(base);                 // 128 rows x 128 B
  float* b_tile = reinterpret_cast<float*>(base + 16384);         // up to 256 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_sm;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(base)[i] = 0.0f;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_sm, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_sm;
  if (warp == 0) {
    const uint32_t idesc = instr_desc(bf16 ? kFmtBF16 : kFmtTF32, 128, n);
    const uint32_t d = tmem + 256;
    const uint64_t a0 = smem_desc_sw128(smem_u32(a_tile), 0), b0 = smem_desc_sw128(smem_u32(b_tile), 0);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (elect_one_sync()) {
      t0 = clock64();
      if (layout == 0) {
        // as the render kernels issue them: descriptors rebuilt per MMA
        for (int i = 0; i < count; ++i) {
          const uint32_t koff = (i & 3) * 32;
          if (ts) {
            if (bf16) mma_f16_ts(d, tmem + (i & 3) * 8, smem_desc_sw128(smem_u32(b_tile), koff), idesc, true);
            else mma_tf32_ts(d, tmem + (i & 3) * 8, smem_desc_sw128(smem_u32(b_tile), koff), idesc, true);
          } else {
            if (bf16) mma_f16_ss(d, smem_desc_sw128(smem_u32(a_tile), koff), smem_desc_sw128(smem_u32(b_tile), koff), idesc, true);
            else mma_tf32_ss(d, smem_desc_sw128(smem_u32(a_tile), koff), smem_desc_sw128(smem_u32(b_tile), koff), idesc, true);
          }
        }
      } else {
        // tight: precomputed descriptors, four K steps unrolled with constant increments (+2 = 32 bytes)
        if (ts) {
          for (int i = 0; i < count; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (bf16) mma_f16_ts(d, tmem + k * 8, b0 + 2 * k, idesc, true);
              else mma_tf32_ts(d, tmem + k * 8, b0 + 2 * k, idesc, true);
            }
          }
        } else {
          for (int i = 0; i < count; i += 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (bf16) mma_f16_ss(d, a0 + 2 * k, b0 + 2 * k, idesc, true);
              else mma_tf32_ss(d, a0 + 2 * k, b0 + 2 * k, idesc, true);
            }
          }
        }
      }
      t1 = clock64();
      mma_commit(&bar);
      mbar_wait(&bar, 0);
      t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}
}  // namespace tpr

extern "C" int tpr_mma_microbench(int32_t n, int32_t bf16, int32_t ts, int32_t count, int32_t tight, long long* out_dev, void* stream) {
  if (!out_dev || n < 16 || n > 256 || (n & 15) || count <= 0) return TPR_E_SHAPE;
  const int smem = 1024 + 16384 + 32768;
  cudaError_t e = cudaFuncSetAttribute(tpr::mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  tpr::mma_bench_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(n, bf16, ts, count, tight, out_dev);
  return (int)cudaGetLastError();
}
