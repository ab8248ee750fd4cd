// Fused ImportanceRenderer.forward with the OSGDecoder on the 5th-generation tensor cores.
//
// One persistent CTA per SM (16 warps, all 512 TMEM columns).  A "group" is R rays of one image
// (R = 8 or 4); a pass over the group is cut into tiles of 128 samples = R rays x DPT consecutive depths
// (DPT = 128/R), so sample (ray r, depth d) is row r*DPT + d%DPT of tile d/DPT -- every TMEM lane
// belongs to one ray for the whole group, which makes the final colour sum a per-lane accumulation.
//
//   per tile   G   gather: 8 lanes x float4 per sample -> A1 operand tile in shared memory
//                  (SWIZZLE_128B, K-major; 3xTF32 keeps a hi and a lo copy, bf16 one packed copy)
//              M1  tcgen05.mma  D1[128x64]  = A1 . W1^T (+ bias through a ones-column MMA)    smem x smem
//              E1  tcgen05.ld D1 -> softplus -> tcgen05.st A2 (hi/lo or bf16) back into TMEM
//              M2  tcgen05.mma  slot[128x48] = A2 . W2^T (+ bias)                              TMEM x smem
//   The raw layer-2 outputs stay in their TMEM slot (6 slots x 48 columns) until the group's final
//   composite; only sigma (column 0) is read back after each pass.  Nothing per-sample touches HBM.
//
// 3xTF32 (x = hi + lo, D = hi.hi + lo.hi + hi.lo) gives fp32-grade accuracy (measured 3e-6 max-abs on
// the decoder outputs) so this path serves the 1e-4 parity mode; the bf16 variant is the >= 50 dB mode.
// Citations relative to /root/reference/g_nerf/ (VR/ = training/volumetric_rendering/).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <cuda_bf16.h>
#include "triplane_b200.h"
#include "tpr_render.cuh"
#include "tpr_tc.cuh"

namespace tpr {
using namespace tc;

constexpr int kTcWorkers = 512;                  // 16 worker warps: gather, activations, per-ray phases
constexpr int kTcWarps = kTcWorkers / 32;
constexpr int kTcThreads = kTcWorkers + 32;      // + one warp that only issues tcgen05.mma
constexpr int kRows = 128;
constexpr int kN1 = 64, kN2 = 48;
// TMEM column map (512 columns allocated)
constexpr uint32_t kColD1 = 0, kColA2Hi = 64, kColA2Lo = 128, kColOnes = 192, kColSlots = 200, kSlotCols = 48;
constexpr int kMaxSlots = 6;

template <int MODE> struct TcTiles;          // every member is a multiple of 1024 B: tiles stay swizzle-aligned
template <> struct TcTiles<0> {              // 3xTF32
  float a1[2][2][kRows * 32];                // [buffer][hi, lo]
  float b1[2][kN1 * 32];                     // [hi, lo]
  float b2[2][2][kN2 * 32];                  // [hi, lo][k block]
  float bias1[kN1 * 32];                     // k0 = hi(b), k1 = lo(b), rest 0
  float bias2[kN2 * 32];
};
template <> struct TcTiles<1> {              // bf16 (rows are still 128 B; layer 1 uses the first 64 B)
  float a1[2][1][kRows * 32];
  float b1[1][kN1 * 32];
  float b2[1][1][kN2 * 32];
  float bias1[kN1 * 32];
  float bias2[kN2 * 32];
};

__device__ __forceinline__ void st_swz_f32(float* tile, int row, int k, float v) {
  tile[row * 32 + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3))] = v;
}
__device__ __forceinline__ void st_swz_bf16(float* tile, int row, int k, float v) {     // 64 bf16 per 128-byte row
  reinterpret_cast<__nv_bfloat16*>(tile)[row * 64 + ((((k >> 3) ^ (row & 7)) << 3) | (k & 7))] = __float2bfloat16_rn(v);
}

template <int MODE>
__device__ void tc_stage_weights(const float* __restrict__ dec, TcTiles<MODE>& tl) {
  const int tid = threadIdx.x;
  for (int i = tid; i < kN1 * 32; i += kTcThreads) { tl.bias1[i] = 0.0f; }
  for (int i = tid; i < kN2 * 32; i += kTcThreads) { tl.bias2[i] = 0.0f; }
  __syncthreads();
  for (int i = tid; i < kN1 * 32; i += kTcThreads) {
    const int n = i >> 5, k = i & 31;
    const float w = dec[kW1tOff + k * kHid + n] * kLog2e;          // layer-1 output in the log2 domain
    if (MODE == 1) st_swz_bf16(tl.b1[0], n, k, w);
    else { float hi, lo; split_tf32(w, hi, lo); st_swz_f32(tl.b1[0], n, k, hi); st_swz_f32(tl.b1[MODE == 0 ? 1 : 0], n, k, lo); }
  }
  for (int i = tid; i < kN2 * 64; i += kTcThreads) {
    const int n = i >> 6, k = i & 63;
    // hidden activations arrive as softplus/ln2; colour rows (n >= 1) produce logits times -log2(e)
    const float w0 = n < kOutPad ? dec[kW2tOff + k * kOutPad + n] : 0.0f;
    const float w = n == 0 ? w0 * kLn2 : -w0;
    if (MODE == 1) st_swz_bf16(tl.b2[0][0], n, k, w);
    else {
      float hi, lo; split_tf32(w, hi, lo);
      st_swz_f32(tl.b2[0][MODE == 0 ? (k >> 5) : 0], n, k & 31, hi);
      st_swz_f32(tl.b2[MODE == 0 ? 1 : 0][MODE == 0 ? (k >> 5) : 0], n, k & 31, lo);
    }
  }
  for (int n = tid; n < kN1 + kN2; n += kTcThreads) {
    const bool l1 = n < kN1;
    const int r = l1 ? n : n - kN1;
    const float b = l1 ? dec[kB1Off + r] * kLog2e
                       : (r < kOutPad ? (r == 0 ? dec[kB2Off] : -kLog2e * dec[kB2Off + r]) : 0.0f);
    float* tile = l1 ? tl.bias1 : tl.bias2;
    if (MODE == 1) {
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      st_swz_bf16(tile, r, 0, hi); st_swz_bf16(tile, r, 1, b - hi);
    } else {
      float hi, lo; split_tf32(b, hi, lo);
      st_swz_f32(tile, r, 0, hi); st_swz_f32(tile, r, 1, lo);
    }
  }
}

struct TcRaySmem {
  float* dep; float* sig; float* wa; float* wb; float* wc; float* ray; float* rayw;
};

// G: gather one tile (rows = R rays x DPT depths) into A1[buf].  Warp w owns rows [8w, 8w+8).
// Step 1: lane 3s+p computes the bilinear taps of (sample s, plane p) once into the warp's tap table.
// Step 2: eight lanes per sample fetch whole 128-byte texels (four channels per lane) and blend.
template <int MODE>
__device__ __forceinline__ void tc_gather_tile(const RenderArgs& a, TcTiles<MODE>& tl, int buf, const float* __restrict__ img,
                                               const TcRaySmem& rs, TapEntry* taps_all, int nr, int Dx, int off, int S,
                                               int t, int dpt_shift) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
  const int dpt = 1 << dpt_shift;
  TapEntry* tw = taps_all + warp * 24;
  {
    const int s = lane / 3, p = lane - s * 3;
    const int row = warp * 8 + s;
    const int r = row >> dpt_shift, di = t * dpt + (row & (dpt - 1));
    if (lane < 24 && r < nr && di < Dx) {
      const float d = rs.dep[r * S + off + di];
      const float* ry = rs.ray + r * 8;
      // origin + depth * direction (VR/renderer.py:105,123), then * 2/box_warp (:61)
      const float px = __fmul_rn(__fadd_rn(ry[0], __fmul_rn(d, ry[3])), a.box_scale);
      const float py = __fmul_rn(__fadd_rn(ry[1], __fmul_rn(d, ry[4])), a.box_scale);
      const float pz = __fmul_rn(__fadd_rn(ry[2], __fmul_rn(d, ry[5])), a.box_scale);
      Taps tp;
      plane_taps(p == 2 ? pz : px, p == 0 ? py : (p == 1 ? pz : px), a.H, a.W, tp);   // (x,y) (x,z) (z,x)
      const int po = p * a.H * a.W * kC;
      *reinterpret_cast<int4*>(tw[lane].off) = make_int4(tp.off[0] + po, tp.off[1] + po, tp.off[2] + po, tp.off[3] + po);
      *reinterpret_cast<float4*>(tw[lane].w) = make_float4(tp.w[0], tp.w[1], tp.w[2], tp.w[3]);
    }
  }
  __syncwarp();
  const float* img_sub = img + sub * 4;
#pragma unroll
  for (int rd = 0; rd < 2; ++rd) {
    const int s = rd * 4 + grp;
    const int row = warp * 8 + s;
    const int r = row >> dpt_shift, di = t * dpt + (row & (dpt - 1));
    if (r < nr && di < Dx) {
      const float4 f = gather_point_taps(img_sub, tw + s * 3);
      if (MODE == 1) {
        uint2 pk = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
        uint8_t* dst = reinterpret_cast<uint8_t*>(tl.a1[buf][0]) + row * 128 + ((((sub >> 1) ^ (row & 7)) << 4) | ((sub & 1) << 3));
        *reinterpret_cast<uint2*>(dst) = pk;
      } else {
        float4 hi, lo;
        split_tf32(f.x, hi.x, lo.x); split_tf32(f.y, hi.y, lo.y); split_tf32(f.z, hi.z, lo.z); split_tf32(f.w, hi.w, lo.w);
        *reinterpret_cast<float4*>(tl.a1[buf][0] + row_chunk_off(row, sub)) = hi;
        *reinterpret_cast<float4*>(tl.a1[buf][MODE == 0 ? 1 : 0] + row_chunk_off(row, sub)) = lo;
      }
    }
  }
  __syncwarp();           // the tap table is rewritten by the next tile
}

// M1: D1 = A1 . W1^T + b1      (one thread)
template <int MODE>
__device__ __forceinline__ void tc_issue_layer1(TcTiles<MODE>& tl, int buf, uint32_t tmem) {
  const uint32_t d1 = tmem + kColD1, ones = tmem + kColOnes;
  if (MODE == 1) {
    const uint32_t idesc = instr_desc(kFmtBF16, 128, kN1);
    const uint32_t a = smem_u32(tl.a1[buf][0]), b = smem_u32(tl.b1[0]);
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) mma_f16_ss(d1, smem_desc_sw128(a, ks * 32), smem_desc_sw128(b, ks * 32), idesc, ks > 0);
    mma_f16_ts(d1, ones, smem_desc_sw128(smem_u32(tl.bias1), 0), idesc, true);
  } else {
    const uint32_t idesc = instr_desc(kFmtTF32, 128, kN1);
    const uint32_t ah = smem_u32(tl.a1[buf][0]), al = smem_u32(tl.a1[buf][MODE == 0 ? 1 : 0]);
    const uint32_t bh = smem_u32(tl.b1[0]), bl = smem_u32(tl.b1[MODE == 0 ? 1 : 0]);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      mma_tf32_ss(d1, smem_desc_sw128(ah, ks * 32), smem_desc_sw128(bh, ks * 32), idesc, ks > 0);
      mma_tf32_ss(d1, smem_desc_sw128(al, ks * 32), smem_desc_sw128(bh, ks * 32), idesc, true);
      mma_tf32_ss(d1, smem_desc_sw128(ah, ks * 32), smem_desc_sw128(bl, ks * 32), idesc, true);
    }
    mma_tf32_ts(d1, ones, smem_desc_sw128(smem_u32(tl.bias1), 0), idesc, true);
  }
}

// M2: slot = A2 . W2^T + b2    (one thread)
template <int MODE>
__device__ __forceinline__ void tc_issue_layer2(TcTiles<MODE>& tl, uint32_t tmem, int slot) {
  const uint32_t d2 = tmem + kColSlots + slot * kSlotCols, ones = tmem + kColOnes;
  if (MODE == 1) {
    const uint32_t idesc = instr_desc(kFmtBF16, 128, kN2);
    const uint32_t b = smem_u32(tl.b2[0][0]);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) mma_f16_ts(d2, tmem + kColA2Hi + ks * 8, smem_desc_sw128(b, ks * 32), idesc, ks > 0);
    mma_f16_ts(d2, ones, smem_desc_sw128(smem_u32(tl.bias2), 0), idesc, true);
  } else {
    const uint32_t idesc = instr_desc(kFmtTF32, 128, kN2);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint32_t bh = smem_u32(tl.b2[0][MODE == 0 ? (ks >> 2) : 0]), bl = smem_u32(tl.b2[MODE == 0 ? 1 : 0][MODE == 0 ? (ks >> 2) : 0]);
      const uint32_t off = (ks & 3) * 32;
      mma_tf32_ts(d2, tmem + kColA2Hi + ks * 8, smem_desc_sw128(bh, off), idesc, ks > 0);
      mma_tf32_ts(d2, tmem + kColA2Lo + ks * 8, smem_desc_sw128(bh, off), idesc, true);
      mma_tf32_ts(d2, tmem + kColA2Hi + ks * 8, smem_desc_sw128(bl, off), idesc, true);
    }
    mma_tf32_ts(d2, ones, smem_desc_sw128(smem_u32(tl.bias2), 0), idesc, true);
  }
}

// E1: D1 -> softplus -> A2 (TMEM).  Warp (q, j) owns lane quarter q and hidden columns [16j, 16j+16).
template <int MODE>
__device__ __forceinline__ void tc_epilogue1(uint32_t tmem) {
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int j = warp >> 2;
  uint32_t r[16];
  tmem_ld16(tmem + kColD1 + lane_base + 16 * j, r);
  tmem_wait_ld();
  if (MODE == 1) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(softplus_log2(__uint_as_float(r[2 * i])), softplus_log2(__uint_as_float(r[2 * i + 1])));
    tmem_st8(tmem + kColA2Hi + lane_base + 8 * j, pk);
  } else {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a, b;
      split_tf32(softplus_log2(__uint_as_float(r[i])), a, b);
      hi[i] = __float_as_uint(a); lo[i] = __float_as_uint(b);
    }
    tmem_st16(tmem + kColA2Hi + lane_base + 16 * j, hi);
    tmem_st16(tmem + kColA2Lo + lane_base + 16 * j, lo);
  }
  tmem_wait_st();
}

// mbarriers shared by the worker warps and the MMA warp
struct TcBarriers {
  uint64_t a1_full[2];     // 16 worker arrivals: A1[b] gathered (and published to the async proxy)
  uint64_t d1_full;        // tcgen05.commit: layer 1 of the current tile is in TMEM
  uint64_t a2_full;        // 16 worker arrivals: activations written back to TMEM
  uint64_t m2_done;        // tcgen05.commit: layer 2 of the current tile has consumed A2 / filled its slot
};

template <int MODE, int E>
__global__ void __launch_bounds__(kTcThreads, 1) render_tc_kernel(const RenderArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the swizzled tiles; offsetting the __shared__ array (rather than round-tripping
  // through an integer) keeps every access an LDS/STS instead of a generic LD/ST
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  TcTiles<MODE>& tl = *reinterpret_cast<TcTiles<MODE>*>(base);
  TapEntry* taps = reinterpret_cast<TapEntry*>(base + sizeof(TcTiles<MODE>));
  float* fl = reinterpret_cast<float*>(base + sizeof(TcTiles<MODE>) + sizeof(TapEntry) * kTcWarps * 24);
  const int R = a.R, Dc = a.Dc, Df = a.Df, S = Dc + Df;
  const int dpt_shift = R == 8 ? 4 : 5, dpt = 1 << dpt_shift;
  TcRaySmem rs;
  rs.dep = fl; rs.sig = rs.dep + R * S; rs.wa = rs.sig + R * S; rs.wb = rs.wa + R * S; rs.wc = rs.wb + R * S;
  rs.ray = rs.wc + R * S; rs.rayw = rs.ray + R * 8;
  __shared__ TcBarriers bars;
  __shared__ uint32_t tmem_base_sm;
  __shared__ unsigned range_sm[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, j = (warp >> 2) & 3;
  const uint32_t lane_base = (uint32_t)(q * 32) << 16;

  if (tid == 0) {
    range_sm[0] = 0xffffffffu; range_sm[1] = 0u;
    mbar_init(&bars.a1_full[0], kTcWarps); mbar_init(&bars.a1_full[1], kTcWarps);
    mbar_init(&bars.d1_full, 1); mbar_init(&bars.a2_full, kTcWarps); mbar_init(&bars.m2_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_base_sm, 512); tmem_relinquish(); }
  tc_stage_weights<MODE>(a.dec, tl);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_sm;
  if (warp < 4) {                               // the ones block (A operand of the bias MMAs): k0 = k1 = 1
    uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (MODE == 1) v[0] = 0x3f803f80u; else { v[0] = 0x3f800000u; v[1] = 0x3f800000u; }
    tmem_st8(tmem + kColOnes + lane_base, v);
    tmem_wait_st();
  }
  tcgen05_fence_before();
  __syncthreads();

  const int nsl_c = (Dc + dpt - 1) >> dpt_shift, nsl_f = (Df + dpt - 1) >> dpt_shift;
  const int n_pass = Df > 0 ? 2 : 1;

  if (warp == kTcWarps) {
    // =========================== MMA warp: waits for operands, issues, commits ===========================
    uint32_t pa[2] = {0, 0}, pa2 = 0;
    tcgen05_fence_after();
    for (long long grp = blockIdx.x; grp < a.n_tiles; grp += gridDim.x) {
#pragma unroll 1
      for (int pass = 0; pass < n_pass; ++pass) {
        const int T = pass == 0 ? nsl_c : nsl_f, slot0 = pass == 0 ? 0 : nsl_c;
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
          const int b = t & 1;
          mbar_wait(&bars.a1_full[b], pa[b]); pa[b] ^= 1;
          tcgen05_fence_after();
          if (elect_one_sync()) {
            tc_issue_layer1<MODE>(tl, b, tmem);
            mma_commit(&bars.d1_full);
          }
          __syncwarp();
          mbar_wait(&bars.a2_full, pa2); pa2 ^= 1;
          tcgen05_fence_after();
          if (elect_one_sync()) {
            tc_issue_layer2<MODE>(tl, tmem, slot0 + t);
            mma_commit(&bars.m2_done);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== worker warps ===========================
    uint32_t pd1 = 0, pm2 = 0;                    // mbarrier phase parities
    const bool per_ray = a.rs != nullptr;
    const size_t img_stride = (size_t)3 * a.H * a.W * kC;
    float mn = __int_as_float(0x7f800000), mx = -__int_as_float(0x7f800000);
    float smn = mn, smx = mx;                         // running depth range of the current clamp slot
    int cur_slot = 0;
    long long tprev = clock64();
#define TPR_MARK(i) do { if (a.dbg != nullptr && blockIdx.x == 0 && tid == 0) { long long now_ = clock64(); a.dbg[i] += now_ - tprev; tprev = now_; } } while (0)
#define WORKER_SYNC() named_bar_sync(1, kTcWorkers)

    for (long long grp = blockIdx.x; grp < a.n_tiles; grp += gridDim.x) {
      const long long n = grp / a.tiles_per_img;
      const long long gi = grp - n * a.tiles_per_img;
      // Which rays form the group.  Column mode (rays are a col_w-wide image, x fastest, VR/ray_sampler.py:44):
      // R vertically adjacent pixels of one image column -- they share their (x,z) footprint, i.e. their taps on
      // two of the three planes (plane 1 and 2 are both functions of (x,z), VR/renderer.py:29-37), which cuts the
      // L2->L1 traffic of the gather several-fold.  Otherwise R consecutive rays.
      const bool colm = a.col_w > 0;
      const long long ray0 = n * a.rays_per_img + (colm ? (gi / a.col_w) * R * a.col_w + gi % a.col_w : gi * R);
      const long long rstride = colm ? a.col_w : 1;
      const int nr = colm ? R : (int)min((long long)R, a.rays_per_img - gi * R);
#define RAY_G(r) (ray0 + (long long)(r) * rstride)
      const float* img = a.planes + (size_t)(n % a.plane_sets) * img_stride;
      if (range_slot(a, (int)n) != cur_slot) { range_fold(a, cur_slot, smn, smx, mn, mx, lane); cur_slot = range_slot(a, (int)n); }
      // ---- rays + coarse depths
      if (tid < nr * 6) {
        const int r = tid / 6, c = tid - r * 6;
        const long long g = RAY_G(r);
        rs.ray[r * 8 + c] = c < 3 ? a.origins[g * 3 + c] : a.dirs[g * 3 + c - 3];
        if (c == 0) {
          rs.ray[r * 8 + 6] = per_ray ? a.rs[g] : a.ray_start;
          rs.ray[r * 8 + 7] = per_ray ? a.re[g] : a.ray_end;
        }
      }
      WORKER_SYNC();
      for (int s = tid; s < nr * Dc; s += kTcWorkers) {
        const int r = s / Dc, k = s - r * Dc;
        rs.dep[r * S + k] = coarse_depth(a, k, __ldg(a.jitter + RAY_G(r) * Dc + k), rs.ray[r * 8 + 6], rs.ray[r * 8 + 7], per_ray);
      }
      WORKER_SYNC();
      TPR_MARK(0);

#pragma unroll 1
      for (int pass = 0; pass < n_pass; ++pass) {
        const int Dx = pass == 0 ? Dc : Df, off = pass == 0 ? 0 : Dc;
        const int T = pass == 0 ? nsl_c : nsl_f, slot0 = pass == 0 ? 0 : nsl_c;
        if (pass == 1) {
          // ---- importance resampling, one warp per ray
          for (int r = warp; r < nr; r += kTcWarps)
            warp_resample_ray(a, rs.dep + r * S, rs.sig + r * S, rs.wa + r * S, rs.wb + r * S, rs.wc + r * S,
                              rs.dep + r * S + Dc, RAY_G(r), lane, a.u + RAY_G(r) * Df);
          WORKER_SYNC();
          TPR_MARK(8);
        }
        // ---- tile pipeline
        tc_gather_tile<MODE>(a, tl, 0, img, rs, taps, nr, Dx, off, S, 0, dpt_shift);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.a1_full[0]);
        TPR_MARK(1);
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
          if (t + 1 < T) {
            tc_gather_tile<MODE>(a, tl, (t + 1) & 1, img, rs, taps, nr, Dx, off, S, t + 1, dpt_shift);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.a1_full[(t + 1) & 1]);
          }
          TPR_MARK(3);
          mbar_wait(&bars.d1_full, pd1); pd1 ^= 1;
          if (t > 0) { mbar_wait(&bars.m2_done, pm2); pm2 ^= 1; }   // layer 2 of the previous tile has consumed A2
          tcgen05_fence_after();
          TPR_MARK(4);
          tc_epilogue1<MODE>(tmem);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.a2_full);
          TPR_MARK(5);
        }
        mbar_wait(&bars.m2_done, pm2); pm2 ^= 1;
        tcgen05_fence_after();
        // ---- sigma = column 0 of every slot of this pass
        for (int sl = j; sl < T; sl += 4) {
          const float sg = __uint_as_float(tmem_ld1(tmem + kColSlots + (slot0 + sl) * kSlotCols + lane_base));
          tmem_wait_ld();
          const int row = q * 32 + lane, r = row >> dpt_shift, di = sl * dpt + (row & (dpt - 1));
          if (r < nr && di < Dx) rs.sig[r * S + off + di] = sg;
        }
        tcgen05_fence_before();
        WORKER_SYNC();
        TPR_MARK(7);
      }
      // ---- sort + final march: omega per sample (scattered to original order), depth, weight sum
      for (int r = warp; r < nr; r += kTcWarps) {
        float wsum, dnum;
        warp_sort_and_weights<E, true>(rs.dep + r * S, rs.sig + r * S, rs.wa + r * S, nullptr, S, lane, wsum, dnum, smn, smx);
        if (lane == 0) {
          rs.rayw[r] = wsum;
          a.depth[RAY_G(r)] = dnum / wsum;            // NaN -> inf and the clamp happen in finish_kernel
          a.wsum[RAY_G(r)] = wsum;
        }
      }
      WORKER_SYNC();
      TPR_MARK(9);
      // ---- composite: warp (q, j) sums channels [8j, 8j+8) over the samples held by its lanes
      {
        tcgen05_fence_after();
        const int row = q * 32 + lane, r = row >> dpt_shift, i = row & (dpt - 1);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int nsl = nsl_c + (Df > 0 ? nsl_f : 0);
#pragma unroll 1
        for (int sl = 0; sl < nsl; ++sl) {
          const bool fine = sl >= nsl_c;
          const int di = (fine ? sl - nsl_c : sl) * dpt + i;
          const bool valid = r < nr && di < (fine ? Df : Dc);
          const float om = valid ? rs.wa[r * S + (fine ? Dc : 0) + di] : 0.0f;
          uint32_t v[8];
          tmem_ld8(tmem + kColSlots + sl * kSlotCols + lane_base + 1 + 8 * j, v);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = valid ? fmaf(om, colour_act_neglog2(__uint_as_float(v[c])), acc[c]) : acc[c];
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          if (o < dpt) {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(kFull, acc[c], o);
          }
        }
        if (i == 0 && r < nr) {
          const float ws = rs.rayw[r];
          float o8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float v = acc[c];
            if (a.white_back) v = v + 1.0f - ws;         // VR/ray_marcher.py:52-53
            o8[c] = v * 2.0f - 1.0f;                     // :55
          }
          long long cstride;
          float* dst1 = rgb_ptr(a, RAY_G(r), (int)n, cstride) + 8 * j * cstride;
          if (a.nchw) {
#pragma unroll
            for (int c = 0; c < 8; ++c) dst1[c * cstride] = o8[c];
          } else {
            float4* dst = reinterpret_cast<float4*>(dst1);
            dst[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
            dst[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
          }
        }
        tcgen05_fence_before();
      }
      WORKER_SYNC();                                     // slots, ray arrays free for the next group
      TPR_MARK(10);
    }
    range_fold(a, cur_slot, smn, smx, mn, mx, lane);
    mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0 && mn <= mx) {
      atomicMin(&range_sm[0], float_to_ordered(mn));
      atomicMax(&range_sm[1], float_to_ordered(mx));
    }
  }
  __syncthreads();
  if (tid == 0 && range_sm[0] <= range_sm[1]) {
    atomicMin(a.range_enc + 0, range_sm[0]);
    atomicMax(a.range_enc + 1, range_sm[1]);
  }
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- host-side selection ---------------------------------------------------------------------
// Returns the rays-per-group (8 or 4) the tensor-core kernel would use, or 0 if (Dc, Df) does not fit
// its 6 TMEM slots (the caller then uses the FFMA kernel).
int tc_rays_per_group(int Dc, int Df) {
  for (int R = 8; R >= 4; R >>= 1) {
    const int dpt = kRows / R;
    const int slots = (Dc + dpt - 1) / dpt + (Df > 0 ? (Df + dpt - 1) / dpt : 0);
    if (slots <= kMaxSlots) return R;
  }
  return 0;
}

template <int MODE>
static size_t tc_smem_bytes(int R, int S) {
  return 1024 + sizeof(TcTiles<MODE>) + sizeof(TapEntry) * kTcWarps * 24 + sizeof(float) * ((size_t)5 * R * S + (size_t)R * 8 + R);
}

typedef void (*TcKernel)(const RenderArgs);
template <int MODE>
static TcKernel pick_kernel(int S) {
  return S <= 64 ? render_tc_kernel<MODE, 2> : S <= 128 ? render_tc_kernel<MODE, 4> : render_tc_kernel<MODE, 8>;
}

// Launch; returns cudaError_t (0 = ok).  `a.R`, `a.n_tiles`, `a.tiles_per_img` are filled here.
int launch_render_tc(RenderArgs a, int bf16, int sms, int smem_optin, long long n_img, long long n_rays, cudaStream_t st) {
  const int S = a.Dc + a.Df;
  a.R = tc_rays_per_group(a.Dc, a.Df);
  if (a.col_w > 0 && (a.col_w % a.R != 0 || (long long)a.col_w * a.col_w != n_rays)) a.col_w = 0;
  a.tiles_per_img = (n_rays + a.R - 1) / a.R;
  a.n_tiles = a.tiles_per_img * n_img;
  TcKernel k = bf16 ? pick_kernel<1>(S) : pick_kernel<0>(S);
  const size_t smem = bf16 ? tc_smem_bytes<1>(a.R, S) : tc_smem_bytes<0>(a.R, S);
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, k);
  if (e != cudaSuccess) return (int)e;
  if ((int)smem > smem_optin - (int)fa.sharedSizeBytes) return (int)cudaErrorInvalidValue;
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const long long grid = a.n_tiles < sms ? a.n_tiles : sms;      // one CTA per SM: each owns all 512 TMEM columns
  k<<<(unsigned)grid, kTcThreads, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace tpr
