// Building blocks shared by the warp-specialised kernels (tpr_render_ws.cu: the fused forward; tpr_run_model_ws.cu:
// run_model for arbitrary points): operand tiles, decoder weight staging, the tap table entry, the bilinear blend of
// one sample, tcgen05.mma issue for the two decoder layers and the softplus epilogue.
// Citations relative to /root/reference/g_nerf/ (VR/ = training/volumetric_rendering/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <cuda_bf16.h>
#include "tpr_device.cuh"
#include "tpr_tc.cuh"

namespace tpr {
namespace ws {
using namespace tc;

constexpr int kGatherWarps = 16, kDecodeWarps = 8, kRayWarps = 8;
constexpr int kThreads = 32 * (kGatherWarps + kDecodeWarps + kRayWarps);      // 1024
constexpr int kFirstDecodeWarp = kGatherWarps, kFirstRayWarp = kGatherWarps + kDecodeWarps;
constexpr int kRows = 128;                  // samples per tile = TMEM lanes
constexpr int kN1 = 64, kNc = 32;           // layer 1 width, colour outputs
constexpr int kBufs = 3;                    // A1 operand tiles in flight
constexpr int kCtx = 4;                     // ray-group contexts in flight
constexpr int kSlotCols = 32;

// TMEM column maps.  3xTF32: the hi half of the layer-2 A operand overwrites D1 in place (layer 1 of the next
// tile is issued after layer 2 of this one, and the tensor pipe executes in issue order).
template <int MODE> struct Cols;
template <> struct Cols<0> { static constexpr uint32_t d1 = 0, a2hi = 0, a2lo = 64, slots = 128; static constexpr int ns = 12; };
template <> struct Cols<1> { static constexpr uint32_t d1 = 0, a2hi = 64, a2lo = 64, slots = 96; static constexpr int ns = 13; };

// Only the two dense contractions run on the tensor cores: hidden = A1.W1^T (N = 64) and colours = A2.W2c^T
// (N = 32).  Biases and the single sigma row of layer 2 are applied by the epilogue in fp32 FFMA: a tcgen05.mma
// costs ~75 cycles of issue time whatever its N (measured, profiles/), so 1-row and bias MMAs are poor value.
template <int MODE> struct Tiles;           // every MMA operand member is a multiple of 1024 B: tiles stay swizzle-aligned
template <> struct Tiles<0> {               // 3xTF32: [hi, lo] copies
  float a1[kBufs][2][kRows * 32];
  float b1[2][kN1 * 32];
  float b2c[2][2][kNc * 32];                // [hi, lo][k block]
  float bias1[kN1];                         // b1 * log2e
  float w2s[kN1];                           // sigma row of W2, * ln2 (hidden activations are softplus/ln2)
  float bias2[kNc + 4];                     // [0..31] = -log2e * colour bias, [32] = sigma bias
  float psig[kRows];                        // sigma partial sums of hidden units 32..63
};
template <> struct Tiles<1> {               // bf16 (rows are still 128 B; layer 1 uses the first 64 B)
  float a1[kBufs][1][kRows * 32];
  float b1[1][kN1 * 32];
  float b2c[1][1][kNc * 32];
  float bias1[kN1];
  float w2s[kN1];
  float bias2[kNc + 4];
  float psig[kRows];
};

__device__ __forceinline__ void st_swz_f32(float* tile, int row, int k, float v) {
  tile[row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3))] = v;
}
__device__ __forceinline__ void st_swz_bf16(float* tile, int row, int k, float v) {     // 64 bf16 per 128-byte row
  reinterpret_cast<__nv_bfloat16*>(tile)[row * 64 + ((((k >> 3) ^ (row & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
}

// Packed decoder (tpr_device.cuh) -> MMA B operands.  Layer-1 outputs are produced in the log2 domain (log2e
// folded into W1/b1), the activation returns softplus/ln2, so sigma weights carry ln2 and colour weights a minus
// sign: the colour slot holds -logit*log2e, which colour_act_neglog2 turns into the sigmoid with one EX2.
template <int MODE>
__device__ void stage_weights(const float* __restrict__ dec, Tiles<MODE>& tl) {
  const int tid = threadIdx.x, nthreads = blockDim.x;
  constexpr int L = MODE == 0 ? 1 : 0;      // index of the lo copy (aliases hi in bf16 mode, never written there)
  for (int i = tid; i < kN1 * 32; i += nthreads) {
    const int n = i >> 5, k = i & 31;
    const float w = dec[kW1tOff + k * kHid + n] * kLog2e;
    if (MODE == 1) st_swz_bf16(tl.b1[0], n, k, w);
    else { float hi, lo; split_tf32(w, hi, lo); st_swz_f32(tl.b1[0], n, k, hi); st_swz_f32(tl.b1[L], n, k, lo); }
  }
  for (int i = tid; i < kNc * 64; i += nthreads) {
    const int n = i >> 6, k = i & 63;                       // colour n = decoder output n + 1
    const float w = -dec[kW2tOff + k * kOutPad + n + 1];
    if (MODE == 1) st_swz_bf16(tl.b2c[0][0], n, k, w);
    else {
      float hi, lo; split_tf32(w, hi, lo);
      st_swz_f32(tl.b2c[0][k >> 5], n, k & 31, hi);
      st_swz_f32(tl.b2c[L][k >> 5], n, k & 31, lo);
    }
  }
  for (int i = tid; i < kN1; i += nthreads) {
    tl.bias1[i] = dec[kB1Off + i] * kLog2e;
    tl.w2s[i] = dec[kW2tOff + i * kOutPad] * kLn2;
  }
  for (int i = tid; i < kNc + 1; i += nthreads) tl.bias2[i] = i < kNc ? -kLog2e * dec[kB2Off + 1 + i] : dec[kB2Off];
}

// Tap table entry for one (sample, plane): the four texels as offsets in 16-byte units from the image's plane
// block (plane offset included) and the four bilinear weights, each stored twice so that a weight is directly the
// {w, w} operand of a packed FFMA2.
struct __align__(16) Tap2 { uint32_t off[4]; float w2[8]; };

// Bilinear blend of one sample (VR/renderer.py:64) by the eight lanes that own it: lane `sub` fetches its four
// channels of the twelve texels listed in the sample's three tap-table entries `te`, one plane (four texels) at a
// time -- tpr_gather_microbench (profiles/) shows that on B200 a shallow queue per thread and many warps sustains
// more random-line bandwidth than twelve loads in flight per thread -- sums the planes and writes its 16 bytes of
// row `row` of the A1 operand tile (hi / lo tf32 copies, or packed bf16).
template <int MODE>
// `keep` (training only): where this lane's four summed features also go in HBM.
__device__ __forceinline__ void blend_sample(float* a1_hi, float* a1_lo, const ulonglong2* base, const Tap2* te, int row, int sub,
                                             float4* keep = nullptr) {
  uint64_t f01 = 0ull, f23 = 0ull;           // channels (0,1) and (2,3) of this lane, summed over the planes
#pragma unroll 1
  for (int p = 0; p < 3; ++p) {
    const uint4 o = *reinterpret_cast<const uint4*>(te[p].off);
    const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(te[p].w2), wb = *reinterpret_cast<const ulonglong2*>(te[p].w2 + 4);
    const ulonglong2 v0 = __ldg(base + o.x), v1 = __ldg(base + o.y), v2 = __ldg(base + o.z), v3 = __ldg(base + o.w);
    uint64_t a01 = fma2(wa.x, v0.x, 0ull), a23 = fma2(wa.x, v0.y, 0ull);
    a01 = fma2(wa.y, v1.x, a01); a23 = fma2(wa.y, v1.y, a23);
    a01 = fma2(wb.x, v2.x, a01); a23 = fma2(wb.x, v2.y, a23);
    a01 = fma2(wb.y, v3.x, a01); a23 = fma2(wb.y, v3.y, a23);
    f01 = add2(f01, a01); f23 = add2(f23, a23);
  }
  float4 f;
  unpack2(f01, f.x, f.y); unpack2(f23, f.z, f.w);
  if (keep != nullptr) *keep = f;
  if (MODE == 1) {
    uint2 pk = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
    uint8_t* dst = reinterpret_cast<uint8_t*>(a1_hi) + row * 128 + ((((sub >> 1) ^ (row & 7)) << 4) | ((sub & 1) << 3));
    *reinterpret_cast<uint2*>(dst) = pk;
  } else {
    float4 hi, lo;
    split_tf32(f.x, hi.x, lo.x); split_tf32(f.y, hi.y, lo.y); split_tf32(f.z, hi.z, lo.z); split_tf32(f.w, hi.w, lo.w);
    *reinterpret_cast<float4*>(a1_hi + row_chunk_off(row, sub)) = hi;
    *reinterpret_cast<float4*>(a1_lo + row_chunk_off(row, sub)) = lo;
  }
}

// ---------------------------------------------------------------------------------------------------------
// DECODE: MMA issue (one thread)
// ---------------------------------------------------------------------------------------------------------
// The issuing thread's own instruction stream bounds the MMA rate for these small shapes (tpr_mma_microbench,
// profiles/: ~70-100 cycles per tcgen05.mma when the descriptors are rebuilt each time, 28 (TS, N = 32) to 76
// (SS, N = 64) when they are not), so every descriptor is the tile block's base descriptor plus a compile-time
// constant: the start-address field counts 16-byte units and all operand tiles live in one Tiles<> block.
#define TPR_OFF16(member) ((uint32_t)(offsetof(Tiles<MODE>, member) >> 4))
#define TPR_D(lo) desc_sw128_from_lo(lo)
template <int MODE>
__device__ __forceinline__ void issue_layer1(uint32_t dlo, int buf, uint32_t tmem) {
  const uint32_t d1 = tmem + Cols<MODE>::d1;
  constexpr uint32_t kTile16 = (kRows * 32 * 4) >> 4;                       // one 16 KB A1 tile
  const uint32_t ah = dlo + TPR_OFF16(a1) + (uint32_t)buf * ((MODE == 0 ? 2 : 1) * kTile16);
  const uint32_t bh = dlo + TPR_OFF16(b1);
  if (MODE == 1) {
    const uint32_t idesc = instr_desc(kFmtBF16, 128, kN1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) mma_f16_ss(d1, TPR_D(ah + 2 * ks), TPR_D(bh + 2 * ks), idesc, ks > 0);
  } else {
    const uint32_t idesc = instr_desc(kFmtTF32, 128, kN1);
    const uint32_t al = ah + kTile16, bl = bh + ((kN1 * 32 * 4) >> 4);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      mma_tf32_ss(d1, TPR_D(ah + 2 * ks), TPR_D(bh + 2 * ks), idesc, ks > 0);
      mma_tf32_ss(d1, TPR_D(al + 2 * ks), TPR_D(bh + 2 * ks), idesc, true);
      mma_tf32_ss(d1, TPR_D(ah + 2 * ks), TPR_D(bl + 2 * ks), idesc, true);
    }
  }
}

// layer 2, colour rows only (N = 32): A from the activation block at `tmem` (+ Cols::a2hi / a2lo), result into the
// 32 TMEM columns at `dc`
template <int MODE>
__device__ __forceinline__ void issue_layer2(uint32_t dlo, uint32_t tmem, uint32_t dc) {
  const uint32_t bh = dlo + TPR_OFF16(b2c);
  constexpr uint32_t kB16 = (kNc * 32 * 4) >> 4;                            // one 4 KB W2 block
  if (MODE == 1) {
    const uint32_t ic = instr_desc(kFmtBF16, 128, kNc);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_f16_ts(dc, tmem + Cols<MODE>::a2hi + ks * 8, TPR_D(bh + 2 * ks), ic, ks > 0);
  } else {
    const uint32_t ic = instr_desc(kFmtTF32, 128, kNc);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t bhk = bh + (ks >> 2) * kB16 + 2 * (ks & 3), blk = bhk + 2 * kB16;   // [hi, lo][k block]
      mma_tf32_ts(dc, tmem + Cols<MODE>::a2hi + ks * 8, TPR_D(bhk), ic, ks > 0);
      mma_tf32_ts(dc, tmem + Cols<MODE>::a2lo + ks * 8, TPR_D(bhk), ic, true);
      mma_tf32_ts(dc, tmem + Cols<MODE>::a2hi + ks * 8, TPR_D(blk), ic, true);
    }
  }
}

// E1: D1 + b1 -> softplus -> A2 (TMEM).  Decode warp (q, h) owns lane quarter q and hidden columns [32h, 32h+32).
// Returns this thread's part of sigma = w2s . hidden over those columns (fp32 FFMA, training/triplane.py:135).
// kStoreA2 = false: sigma only (density grids), the activations are not written back.
template <int MODE, bool kStoreA2 = true>
__device__ __forceinline__ float epilogue1(const Tiles<MODE>& tl, uint32_t tmem, uint32_t lane_base, int h) {
  uint64_t sg2 = 0ull;
  const uint64_t kOne2 = pack2(1.0f, 1.0f);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int col = 32 * h + 16 * c;
    uint32_t r[16];
    tmem_ld16(tmem + Cols<MODE>::d1 + lane_base + col, r);
    tmem_wait_ld();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint64_t act2[4];                        // four packed pairs of activations
#pragma unroll
      for (int i4 = 0; i4 < 2; ++i4) {
        const int o = 8 * half + 4 * i4;
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(tl.bias1 + col + o);
        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(tl.w2s + col + o);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          // softplus in the log2 domain on a pair: x' = D1 + b1', y = x' > 20 log2e ? x' : lg2(1 + 2^x')
          const uint64_t xs2 = add2(pack2(__uint_as_float(r[o + 2 * h2]), __uint_as_float(r[o + 2 * h2 + 1])), h2 == 0 ? b.x : b.y);
          float x0, x1, y0, y1;
          unpack2(xs2, x0, x1);
          unpack2(add2(pack2(ex2_fast(x0), ex2_fast(x1)), kOne2), y0, y1);
          y0 = x0 > 20.0f * kLog2e ? x0 : lg2_fast(y0);
          y1 = x1 > 20.0f * kLog2e ? x1 : lg2_fast(y1);
          act2[2 * i4 + h2] = pack2(y0, y1);
          sg2 = fma2(act2[2 * i4 + h2], h2 == 0 ? w.x : w.y, sg2);
        }
      }
      if (!kStoreA2) {
      } else if (MODE == 1) {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { float y0, y1; unpack2(act2[i], y0, y1); pk[i] = pack_bf16(y0, y1); }
        tmem_st4(tmem + Cols<MODE>::a2hi + lane_base + ((col + 8 * half) >> 1), pk);
      } else {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float y0, y1, l0, l1;
          unpack2(act2[i], y0, y1);
          hi[2 * i] = (__float_as_uint(y0) + 0x1000u) & 0xffffe000u;          // split_tf32, the subtraction packed
          hi[2 * i + 1] = (__float_as_uint(y1) + 0x1000u) & 0xffffe000u;
          unpack2(sub2(act2[i], pack2(__uint_as_float(hi[2 * i]), __uint_as_float(hi[2 * i + 1]))), l0, l1);
          lo[2 * i] = __float_as_uint(l0); lo[2 * i + 1] = __float_as_uint(l1);
        }
        tmem_st8(tmem + Cols<MODE>::a2hi + lane_base + col + 8 * half, hi);
        tmem_st8(tmem + Cols<MODE>::a2lo + lane_base + col + 8 * half, lo);
      }
    }
  }
  if (kStoreA2) tmem_wait_st();
  float s0, s1;
  unpack2(sg2, s0, s1);
  return s0 + s1;
}

}  // namespace ws
}  // namespace tpr
