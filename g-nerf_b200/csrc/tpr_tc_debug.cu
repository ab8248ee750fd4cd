// Diagnostic kernel for the tcgen05 decoder plumbing: one 128-row tile at a time, fully synchronous.
// Layer 1 is an SS MMA (features in SWIZZLE_128B shared memory), layer 2 a TS MMA (activations
// written back to TMEM with tcgen05.st and used as the A operand).  Exposed as tpr_debug_tc_decode so
// tests can check the raw layer outputs against a float64 matmul; not part of the render path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cuda_bf16.h>
#include "triplane_b200.h"
#include "triplane_b200_bench.h"
#include "tpr_device.cuh"
#include "tpr_tc.cuh"

namespace tpr {
using namespace tc;

constexpr int kTileRows = 128;
constexpr int kN1 = 64, kN2 = 48;

// mode 0: tf32 single pass; 1: 3xTF32; 2: bf16
struct TcSmem {
  // all tiles: rows of 128 B, SWIZZLE_128B
  float a1_hi[kTileRows * 32];
  float a1_lo[kTileRows * 32];
  float b1_hi[kN1 * 32];
  float b1_lo[kN1 * 32];
  float b2_hi[2][kN2 * 32];      // two K blocks of 32 (tf32) -- bf16 uses only block 0 (64 bf16 per row)
  float b2_lo[2][kN2 * 32];
  uint64_t bar;
  uint32_t tmem_base;
};

__device__ __forceinline__ void store_swz(float* tile, int row, int k, float v) {
  tile[row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3))] = v;
}
// bf16 element k of a 128-byte row (64 bf16 per row)
__device__ __forceinline__ void store_swz_bf16(float* tile, int row, int k, float v) {
  __nv_bfloat16* t = reinterpret_cast<__nv_bfloat16*>(tile);
  t[row * 64 + ((((k >> 3) ^ (row & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(128, 1) tc_debug_kernel(const float* __restrict__ x, long long P,
                                                          const float* __restrict__ dec, int mode,
                                                          float* __restrict__ hidden, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  TcSmem& s = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- weights -> swizzled operand tiles
  for (int i = tid; i < kN1 * 32; i += 128) {
    const int n = i >> 5, k = i & 31;
    const float w = dec[kW1tOff + k * kHid + n];
    if (mode == 2) store_swz_bf16(s.b1_hi, n, k, w);
    else { float hi, lo; split_tf32(w, hi, lo); store_swz(s.b1_hi, n, k, hi); store_swz(s.b1_lo, n, k, lo); }
  }
  for (int i = tid; i < kN2 * 64; i += 128) {
    const int n = i >> 6, k = i & 63;
    const float w = n < kOutPad ? dec[kW2tOff + k * kOutPad + n] : 0.0f;
    if (mode == 2) store_swz_bf16(s.b2_hi[0], n, k, w);
    else { float hi, lo; split_tf32(w, hi, lo); store_swz(s.b2_hi[k >> 5], n, k & 31, hi); store_swz(s.b2_lo[k >> 5], n, k & 31, lo); }
  }
  if (tid == 0) { mbar_init(&s.bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&s.tmem_base, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s.tmem_base;
  const uint32_t t_d1 = tmem + 0, t_a2hi = tmem + 64, t_a2lo = tmem + 128, t_d2 = tmem + 192;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t parity = 0;
  const uint32_t fmt = mode == 2 ? kFmtBF16 : kFmtTF32;
  const uint32_t idesc1 = instr_desc(fmt, 128, kN1), idesc2 = instr_desc(fmt, 128, kN2);

  for (long long t0 = (long long)blockIdx.x * kTileRows; t0 < P; t0 += (long long)gridDim.x * kTileRows) {
    const long long g = t0 + tid;
    // ---- A1: this thread's row
    for (int k = 0; k < 32; ++k) {
      const float v = g < P ? x[g * 32 + k] : 0.0f;
      if (mode == 2) store_swz_bf16(s.a1_hi, tid, k, v);
      else { float hi, lo; split_tf32(v, hi, lo); store_swz(s.a1_hi, tid, k, hi); store_swz(s.a1_lo, tid, k, lo); }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      const uint32_t a_hi = smem_u32(s.a1_hi), a_lo = smem_u32(s.a1_lo), b_hi = smem_u32(s.b1_hi), b_lo = smem_u32(s.b1_lo);
      if (mode == 2) {
        for (int ks = 0; ks < 2; ++ks)     // K = 16 bf16 = 32 B per MMA
          mma_f16_ss(t_d1, smem_desc_sw128(a_hi, ks * 32), smem_desc_sw128(b_hi, ks * 32), idesc1, ks > 0);
      } else {
        for (int ks = 0; ks < 4; ++ks) {   // K = 8 tf32 = 32 B per MMA
          mma_tf32_ss(t_d1, smem_desc_sw128(a_hi, ks * 32), smem_desc_sw128(b_hi, ks * 32), idesc1, ks > 0);
          if (mode == 1) {
            mma_tf32_ss(t_d1, smem_desc_sw128(a_lo, ks * 32), smem_desc_sw128(b_hi, ks * 32), idesc1, true);
            mma_tf32_ss(t_d1, smem_desc_sw128(a_hi, ks * 32), smem_desc_sw128(b_lo, ks * 32), idesc1, true);
          }
        }
      }
      mma_commit(&s.bar);
    }
    mbar_wait(&s.bar, parity); parity ^= 1;
    tcgen05_fence_after();
    // ---- epilogue 1: + bias, dump, softplus, back to TMEM as the layer-2 A operand
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(t_d1 + lane_base + c0, r);
      tmem_wait_ld();
      float h[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pre = __uint_as_float(r[j]) + dec[kB1Off + c0 + j];
        if (g < P) hidden[g * 64 + c0 + j] = pre;
        h[j] = softplus_f(pre);
      }
      if (mode == 2) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(h[2 * j], h[2 * j + 1]);
        tmem_st8(t_a2hi + lane_base + (c0 >> 1), pk);
      } else {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { float a, b; split_tf32(h[j], a, b); hi[j] = __float_as_uint(a); lo[j] = __float_as_uint(b); }
        tmem_st16(t_a2hi + lane_base + c0, hi);
        tmem_st16(t_a2lo + lane_base + c0, lo);
      }
    }
    tmem_wait_st();
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) {
      tcgen05_fence_after();
      if (mode == 2) {
        const uint32_t b = smem_u32(s.b2_hi[0]);
        for (int ks = 0; ks < 4; ++ks)     // K = 16 bf16 per MMA: 8 TMEM columns, 32 B of the B row
          mma_f16_ts(t_d2, t_a2hi + ks * 8, smem_desc_sw128(b, ks * 32), idesc2, ks > 0);
      } else {
        for (int ks = 0; ks < 8; ++ks) {   // K = 8 tf32 per MMA: 8 TMEM columns
          const uint32_t bh = smem_u32(s.b2_hi[ks >> 2]), bl = smem_u32(s.b2_lo[ks >> 2]);
          const uint32_t off = (ks & 3) * 32;
          mma_tf32_ts(t_d2, t_a2hi + ks * 8, smem_desc_sw128(bh, off), idesc2, ks > 0);
          if (mode == 1) {
            mma_tf32_ts(t_d2, t_a2lo + ks * 8, smem_desc_sw128(bh, off), idesc2, true);
            mma_tf32_ts(t_d2, t_a2hi + ks * 8, smem_desc_sw128(bl, off), idesc2, true);
          }
        }
      }
      mma_commit(&s.bar);
    }
    mbar_wait(&s.bar, parity); parity ^= 1;
    tcgen05_fence_after();
    for (int c0 = 0; c0 < kN2; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(t_d2 + lane_base + c0, r);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int o = c0 + j;
        if (g < P) out[g * kN2 + o] = __uint_as_float(r[j]) + (o < kOutPad ? dec[kB2Off + o] : 0.0f);
      }
    }
    tcgen05_fence_before();
    __syncthreads();       // TMEM and the A1 tile are reused by the next tile
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace tpr

extern "C" int tpr_debug_tc_decode(const float* x, int64_t n_rows, const float* decoder_packed, int32_t mode,
                                   float* hidden, float* out, void* stream) {
  if (!x || !decoder_packed || !hidden || !out || n_rows <= 0 || mode < 0 || mode > 2) return TPR_E_NULL;
  const size_t smem = sizeof(tpr::TcSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(tpr::tc_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return (int)e; }
  long long tiles = (n_rows + 127) / 128;
  int grid = (int)(tiles < 148 ? tiles : 148);
  tpr::tc_debug_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(x, n_rows, decoder_packed, mode, hidden, out);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}
