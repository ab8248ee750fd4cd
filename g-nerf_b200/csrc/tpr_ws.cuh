// Building blocks shared by the warp-specialised kernels (tpr_render_ws.cu: the fused forward; tpr_run_model_ws.cu:
// run_model for arbitrary points): operand tiles, decoder weight staging, the tap table entry, the bilinear blend of
// one sample, tcgen05.mma issue for the two decoder layers and the softplus epilogue.
// Citations relative to /root/reference/g_nerf/ (VR/ = training/volumetric_rendering/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
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
#ifndef TPR_A1_BUFS
#define TPR_A1_BUFS 3
#endif
constexpr int kBufs = TPR_A1_BUFS;          // A1 operand tiles in flight
constexpr int kCtx = 4;                     // ray-group contexts in flight
constexpr int kSlotCols = 32;

// Decoder operand schemes (the template parameter MODE of the kernels):
//   MODE 1  bf16       operands rounded to bf16, one product per layer: the >= 50 dB PSNR mode
//   MODE 2  2xFP16     x = hi + lo with both halves fp16 (hi = rn(x), lo = rn(x - hi): ~22 significant bits), three products
//                      hi.hi + lo.hi + hi.lo per layer: the fp32-grade mode (1e-4 max-abs gate).  It replaced 3xTF32
//                      (kind::tf32, K = 8 per instruction): kind::f16 takes K = 16 per instruction, so the same precision
//                      costs HALF the MMAs (18 instead of 36 per tile), half the operand bytes in shared memory and half the
//                      TMEM columns for the activations -- measured 2.17 vs 2.42 ms at config 2 (profiles/r02_ab4_*.json).
//
// TMEM: layer 1 works in a 64-column STAGE, in which the activations overwrite D1 IN PLACE, 16 hidden units per 16 columns:
// 2xFP16 [hi x 8 columns | lo x 8 columns], bf16 [8 columns | unused]; a decode warp only ever writes columns it has already
// read itself.  The fused forward has ONE stage and fourteen 32-column colour slots behind it: the slot pool -- how far
// gather and decode may run ahead of the ray warps' composite -- is worth more than overlapping layer 1 of the next tile
// with the epilogue (two stages + 12 slots measured 2.215 ms against 2.165 ms at config 2, and the look-ahead itself only
// bought 0.6 %: profiles/r02_ab6_*.json).  run_model (no composite, colours leave TMEM at once) uses two stages.
// bf16: the packed activations go to their own 32 columns behind D1 instead (in place measured 3 % slower there, 2.120 vs
// 2.053 ms, profiles/r02_ab8_*.json), which leaves thirteen slots.
constexpr uint32_t kStageCols = 64;
template <int MODE> struct Cols;
template <> struct Cols<1> { static constexpr uint32_t a2hi = 64, a2lo = 64, a2step = 8, slots = 96; static constexpr int ns = 13; };
template <> struct Cols<2> { static constexpr uint32_t a2hi = 0, a2lo = 8, a2step = 16, slots = 64; static constexpr int ns = 14; };

// Only the two dense contractions run on the tensor cores: hidden = A1.W1^T (N = 64) and colours = A2.W2c^T
// (N = 32).  Biases and the single sigma row of layer 2 are applied by the epilogue in fp32 FFMA: a tcgen05.mma
// costs ~75 cycles of issue time whatever its N (measured, profiles/), so 1-row and bias MMAs are poor value.
template <int MODE> struct Tiles;           // every MMA operand member is a multiple of 1024 B: tiles stay swizzle-aligned
template <> struct Tiles<1> {               // bf16 (rows are still 128 B; layer 1 uses the first 64 B)
  float a1[kBufs][1][kRows * 32];
  float b1[1][kN1 * 32];
  float b2c[1][1][kNc * 32];
  float bias1[kN1];                         // b1 * log2e
  float w2s[kN1];                           // sigma row of W2, * ln2 (hidden activations are softplus/ln2)
  float bias2[kNc + 4];                     // [0..31] = -log2e * colour bias, [32] = sigma bias
  float psig[2 * kRows];                    // sigma partial sums of hidden units 32..63, double buffered by tile parity
};
template <> struct Tiles<2> {               // 2xFP16: a 128-byte row = [hi: 32 x fp16 | lo: 32 x fp16] (layer 1), 64 x fp16 (W2)
  float a1[kBufs][1][kRows * 32];
  float b1[1][kN1 * 32];                    // row n = [W1 hi | W1 lo]
  float b2c[2][1][kNc * 32];                // [hi, lo]: row n = 64 x fp16 of colour row n
  float bias1[kN1];
  float w2s[kN1];
  float bias2[kNc + 4];
  float psig[2 * kRows];
};

__device__ __forceinline__ void st_swz_f32(float* tile, int row, int k, float v) {
  tile[row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3))] = v;
}
__device__ __forceinline__ void st_swz_bf16(float* tile, int row, int k, float v) {     // 64 bf16 per 128-byte row
  reinterpret_cast<__nv_bfloat16*>(tile)[row * 64 + ((((k >> 3) ^ (row & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
}

__device__ __forceinline__ void st_swz_f16(float* tile, int row, int k, __half v) {      // 64 fp16 per 128-byte row
  reinterpret_cast<__half*>(tile)[row * 64 + ((((k >> 3) ^ (row & 7)) << 3) | (k & 7))] = v;
}
// x = hi + lo, both fp16: hi = rn(x) (11 significant bits), lo = rn(x - hi) (x - hi is exact in fp32): ~22 bits, the
// precision class of the tf32 split.  |x| must stay below the fp16 range (65504): plane features and decoder weights are
// O(1..10); saturating conversions keep an out-of-range value finite instead of poisoning the tile.
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f));
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {          // low 16 bits = lower k
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ void unpack_f16x2(uint32_t v, float& lo_elem, float& hi_elem) {
  const __half2 h = *reinterpret_cast<const __half2*>(&v);
  const float2 f = __half22float2(h);
  lo_elem = f.x; hi_elem = f.y;
}

// Packed decoder (tpr_device.cuh) -> MMA B operands.  Layer-1 outputs are produced in the log2 domain (log2e
// folded into W1/b1), the activation returns softplus/ln2, so sigma weights carry ln2 and colour weights a minus
// sign: the colour slot holds -logit*log2e, which colour_act_neglog2 turns into the sigmoid with one EX2.
template <int MODE>
__device__ void stage_weights(const float* __restrict__ dec, Tiles<MODE>& tl) {
  const int tid = threadIdx.x, nthreads = blockDim.x;
  for (int i = tid; i < kN1 * 32; i += nthreads) {
    const int n = i >> 5, k = i & 31;
    const float w = dec[kW1tOff + k * kHid + n] * kLog2e;
    if (MODE == 1) st_swz_bf16(tl.b1[0], n, k, w);
    else { __half hi, lo; split_f16(w, hi, lo); st_swz_f16(tl.b1[0], n, k, hi); st_swz_f16(tl.b1[0], n, 32 + k, lo); }
  }
  for (int i = tid; i < kNc * 64; i += nthreads) {
    const int n = i >> 6, k = i & 63;                       // colour n = decoder output n + 1
    const float w = -dec[kW2tOff + k * kOutPad + n + 1];
    if (MODE == 1) st_swz_bf16(tl.b2c[0][0], n, k, w);
    else { __half hi, lo; split_f16(w, hi, lo); st_swz_f16(tl.b2c[0][0], n, k, hi); st_swz_f16(tl.b2c[MODE == 2 ? 1 : 0][0], n, k, lo); }
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
// row `row` of the A1 operand tile (packed bf16, or fp16 hi and lo halves of the row).
template <int MODE>
// `keep` (training only): where this lane's four summed features also go in HBM.
__device__ __forceinline__ void blend_sample(float* a1_hi, const ulonglong2* base, const Tap2* te, int row, int sub,
                                             float4* keep = nullptr) {
  uint64_t f01 = 0ull, f23 = 0ull;           // channels (0,1) and (2,3) of this lane, summed over the planes
#pragma unroll 1
  for (int p = 0; p < 3; ++p) {
    const uint4 o = *reinterpret_cast<const uint4*>(te[p].off);
    const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(te[p].w2), wb = *reinterpret_cast<const ulonglong2*>(te[p].w2 + 4);
#ifdef TPR_SKIP_PADDING_PLANES   // (skipping the loads of a plane whose four taps are all padding -- 9 % of the coarse samples -- measured
    // SLOWER: the divergent branch costs more than the loads it saves, bf16 2.115 vs 2.070 ms: profiles/r02_ab7_*.json)
    if ((wa.x | wa.y | wb.x | wb.y) == 0ull) continue;
#endif
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
    // four channels -> 8 bytes of the row's hi half (bytes 0..63) and 8 bytes of its lo half (bytes 64..127)
    const uint2 hi = make_uint2(pack_f16x2(f.x, f.y), pack_f16x2(f.z, f.w));
    float h0, h1, h2, h3;
    unpack_f16x2(hi.x, h0, h1); unpack_f16x2(hi.y, h2, h3);
    const uint2 lo = make_uint2(pack_f16x2(f.x - h0, f.y - h1), pack_f16x2(f.z - h2, f.w - h3));
    uint8_t* rowp = reinterpret_cast<uint8_t*>(a1_hi) + row * 128 + ((sub & 1) << 3);
    *reinterpret_cast<uint2*>(rowp + (((sub >> 1) ^ (row & 7)) << 4)) = hi;
    *reinterpret_cast<uint2*>(rowp + ((((sub >> 1) + 4) ^ (row & 7)) << 4)) = lo;
  }
}

// The same blend with 256-bit loads (LDG.E.ENL2.256, sm_100): FOUR lanes own a sample, lane `sub` fetches eight channels
// (32 bytes) of each texel, so one warp-level load instruction covers eight texels and the warp's eight samples of a tile
// are blended in ONE round of three planes instead of two -- half as many dependent L2 round trips per tile, twice the
// bytes in flight per thread (tpr_gather_microbench_v2, profiles/r02_gather_shapes.json: the gather is latency bound, and
// deepening the queue with wider loads is what pays -- 16 warps x 4 x LDG.128: 14.6 TB/s, 16 x 4 x LDG.256: 15.6, whereas
// 16 x 8 x LDG.128 drops to 13.2).  `base` = the image's plane block + sub * 32 bytes, as 16-byte units.
struct U256 { uint64_t a, b, c, d; };        // channels (0,1) (2,3) (4,5) (6,7) of this lane
__device__ __forceinline__ U256 ldg256(const ulonglong2* p) {
  U256 r;
  asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
  return r;
}
template <int MODE>
__device__ __forceinline__ void blend_sample8(float* a1_hi, const ulonglong2* base, const Tap2* te, int row, int sub,
                                              float4* keep = nullptr) {
  uint64_t f0 = 0ull, f1 = 0ull, f2 = 0ull, f3 = 0ull;      // eight channels of this lane, summed over the planes
#pragma unroll 1
  for (int p = 0; p < 3; ++p) {
    const uint4 o = *reinterpret_cast<const uint4*>(te[p].off);
    const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(te[p].w2), wb = *reinterpret_cast<const ulonglong2*>(te[p].w2 + 4);
    const U256 v0 = ldg256(base + o.x), v1 = ldg256(base + o.y), v2 = ldg256(base + o.z), v3 = ldg256(base + o.w);
    uint64_t a0 = fma2(wa.x, v0.a, 0ull), a1 = fma2(wa.x, v0.b, 0ull), a2 = fma2(wa.x, v0.c, 0ull), a3 = fma2(wa.x, v0.d, 0ull);
    a0 = fma2(wa.y, v1.a, a0); a1 = fma2(wa.y, v1.b, a1); a2 = fma2(wa.y, v1.c, a2); a3 = fma2(wa.y, v1.d, a3);
    a0 = fma2(wb.x, v2.a, a0); a1 = fma2(wb.x, v2.b, a1); a2 = fma2(wb.x, v2.c, a2); a3 = fma2(wb.x, v2.d, a3);
    a0 = fma2(wb.y, v3.a, a0); a1 = fma2(wb.y, v3.b, a1); a2 = fma2(wb.y, v3.c, a2); a3 = fma2(wb.y, v3.d, a3);
    f0 = add2(f0, a0); f1 = add2(f1, a1); f2 = add2(f2, a2); f3 = add2(f3, a3);
  }
  float4 fa, fb;
  unpack2(f0, fa.x, fa.y); unpack2(f1, fa.z, fa.w); unpack2(f2, fb.x, fb.y); unpack2(f3, fb.z, fb.w);
  if (keep != nullptr) { keep[0] = fa; keep[1] = fb; }
  if (MODE == 1) {
    // eight bf16 = 16 bytes = chunk `sub` of the row's first 64 bytes
    uint4 pk = make_uint4(pack_bf16(fa.x, fa.y), pack_bf16(fa.z, fa.w), pack_bf16(fb.x, fb.y), pack_bf16(fb.z, fb.w));
    uint8_t* dst = reinterpret_cast<uint8_t*>(a1_hi) + row * 128 + ((sub ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = pk;
  } else {
    const uint4 hi = make_uint4(pack_f16x2(fa.x, fa.y), pack_f16x2(fa.z, fa.w), pack_f16x2(fb.x, fb.y), pack_f16x2(fb.z, fb.w));
    float h[8];
    unpack_f16x2(hi.x, h[0], h[1]); unpack_f16x2(hi.y, h[2], h[3]); unpack_f16x2(hi.z, h[4], h[5]); unpack_f16x2(hi.w, h[6], h[7]);
    const uint4 lo = make_uint4(pack_f16x2(fa.x - h[0], fa.y - h[1]), pack_f16x2(fa.z - h[2], fa.w - h[3]),
                                pack_f16x2(fb.x - h[4], fb.y - h[5]), pack_f16x2(fb.z - h[6], fb.w - h[7]));
    uint8_t* rowp = reinterpret_cast<uint8_t*>(a1_hi) + row * 128;
    *reinterpret_cast<uint4*>(rowp + ((sub ^ (row & 7)) << 4)) = hi;
    *reinterpret_cast<uint4*>(rowp + (((sub + 4) ^ (row & 7)) << 4)) = lo;
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
// layer 1 of A1 buffer `buf` into the 64 columns at `d1` (a stage)
template <int MODE>
__device__ __forceinline__ void issue_layer1(uint32_t dlo, int buf, uint32_t d1) {
  constexpr uint32_t kTile16 = (kRows * 32 * 4) >> 4;                       // one 16 KB A1 tile
  const uint32_t ah = dlo + TPR_OFF16(a1) + (uint32_t)buf * kTile16;
  const uint32_t bh = dlo + TPR_OFF16(b1);
  if (MODE == 1) {
    const uint32_t idesc = instr_desc(kFmtBF16, 128, kN1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) mma_f16_ss(d1, TPR_D(ah + 2 * ks), TPR_D(bh + 2 * ks), idesc, ks > 0);
  } else {
    // rows are [hi | lo]: the lo half starts 64 bytes (= 4 sixteen-byte units) into the row
    const uint32_t idesc = instr_desc(kFmtF16, 128, kN1);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      mma_f16_ss(d1, TPR_D(ah + 2 * ks), TPR_D(bh + 2 * ks), idesc, ks > 0);              // hi . hi
      mma_f16_ss(d1, TPR_D(ah + 4 + 2 * ks), TPR_D(bh + 2 * ks), idesc, true);            // lo . hi
      mma_f16_ss(d1, TPR_D(ah + 2 * ks), TPR_D(bh + 4 + 2 * ks), idesc, true);            // hi . lo
    }
  }
}

// layer 2, colour rows only (N = 32): A = the activations of the stage at `st` (hidden units [16 ks, 16 ks + 16) in its
// columns 16 ks + [0, 8), their lo halves in 16 ks + [8, 16); bf16: columns 64 + 8 ks + [0, 8)), result into the 32 TMEM columns at `dc`
template <int MODE>
__device__ __forceinline__ void issue_layer2(uint32_t dlo, uint32_t st, uint32_t dc) {
  const uint32_t bh = dlo + TPR_OFF16(b2c);
  constexpr uint32_t kB16 = (kNc * 32 * 4) >> 4;                            // one 4 KB W2 block
  if (MODE == 1) {
    const uint32_t ic = instr_desc(kFmtBF16, 128, kNc);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_f16_ts(dc, st + Cols<MODE>::a2hi + ks * Cols<MODE>::a2step, TPR_D(bh + 2 * ks), ic, ks > 0);
  } else {
    const uint32_t ic = instr_desc(kFmtF16, 128, kNc);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      mma_f16_ts(dc, st + Cols<MODE>::a2hi + ks * 16, TPR_D(bh + 2 * ks), ic, ks > 0);
      mma_f16_ts(dc, st + Cols<MODE>::a2lo + ks * 16, TPR_D(bh + 2 * ks), ic, true);
      mma_f16_ts(dc, st + Cols<MODE>::a2hi + ks * 16, TPR_D(bh + kB16 + 2 * ks), ic, true);
    }
  }
}

// E1: D1 + b1 -> softplus -> A2, in place in the stage at `st`.  Decode warp (q, h) owns lane quarter q and hidden columns [32h, 32h+32).
// Returns this thread's part of sigma = w2s . hidden over those columns (fp32 FFMA, training/triplane.py:135).
// kStoreA2 = false: sigma only (density grids), the activations are not written back.
template <int MODE, bool kStoreA2 = true>
__device__ __forceinline__ float epilogue1(const Tiles<MODE>& tl, uint32_t st, uint32_t lane_base, int h) {
  uint64_t sg2 = 0ull;
  const uint64_t kOne2 = pack2(1.0f, 1.0f);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int col = 32 * h + 16 * c;
    uint32_t r[16];
    tmem_ld16(st + lane_base + col, r);
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
        tmem_st4(st + Cols<MODE>::a2hi + lane_base + ((col + 8 * half) >> 1), pk);
      } else {
        // in place: these 16 hidden units' own 16 columns (already read) become [hi x 8 | lo x 8]
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float y0, y1, h0, h1;
          unpack2(act2[i], y0, y1);
          hi[i] = pack_f16x2(y0, y1);
          unpack_f16x2(hi[i], h0, h1);
          lo[i] = pack_f16x2(y0 - h0, y1 - h1);
        }
        tmem_st4(st + Cols<MODE>::a2hi + lane_base + col + 4 * half, hi);
        tmem_st4(st + Cols<MODE>::a2lo + lane_base + col + 4 * half, lo);
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
