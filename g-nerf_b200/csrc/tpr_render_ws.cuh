#pragma once
// (template header: included by tpr_render_ws.cu and the four instantiation files)
// Warp-specialised fused ImportanceRenderer.forward (VR/renderer.py:88-140) for sm_100a.
//
// Plane gather, OSGDecoder on tcgen05 (2xFP16 = the fp32-grade mode, or bf16), in-register ray march / CDF / sort; the
// three kinds of work run CONCURRENTLY on one SM instead of taking turns, because each of them alone leaves the SM mostly
// idle (ncu on a turn-taking predecessor: 36 % issue utilisation; the gather is bound by L2 latency, the per-ray phases by
// dependent-instruction latency):
//
//   warps  0-15  GATHER    per tile of 128 samples: bilinear taps -> 12 x 128-byte texel reads per sample -> A1
//                          operand tile in shared memory (SWIZZLE_128B), three tiles deep
//   warps 16-23  DECODE    warp 16 lane 0 issues the tcgen05.mma; all eight run the softplus epilogue
//                          (tcgen05.ld D1 -> EX2/LG2 -> tcgen05.st A2) and read sigma back after layer 2
//   warps 24-31  RAYS      one warp per ray: coarse depths, coarse march + pdf + CDF + inverse-CDF draws,
//                          depth sort + final march, then the colour composite straight out of TMEM
//
// The roles are decoupled by a software pipeline over ray groups (a group = R rays of one image):
//   GATHER/DECODE job order:  C(0) C(1) F(0) C(2) F(1) ...    (C = coarse pass, F = fine pass of a group)
//   RAYS step g:              resample(g)  setup(g+2)  sort+composite(g-1)
// so the importance resampling of group g hides behind the coarse gather of group g+1 and its sort/composite
// behind the next jobs.  Layer 1 accumulates into a 64-column TMEM stage and its activations overwrite it in place; layer-2
// colour outputs stay in TMEM until the group's composite: a pool of fourteen 32-column slots with flow control (a slot is
// reused only after the composite of the group that owned it); sigma is the epilogue's fp32 dot product with the sigma row
// of layer 2 and goes to shared memory per tile.
// mbarriers carry every hand-off; nothing per-sample ever touches HBM.
//
// Citations relative to /root/reference/g_nerf/ (VR/ = training/volumetric_rendering/).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <cuda_bf16.h>
#include "triplane_b200.h"
#include "tpr_render.cuh"
#include "tpr_ws.cuh"

namespace tpr {
using namespace tc;

namespace ws {

struct Barriers {
  uint64_t a1_full[kBufs];      // 16 gather-warp arrivals: tile gathered and published to the async proxy
  uint64_t a1_free[kBufs];      // tcgen05.commit: layer 1 has consumed the tile
  uint64_t d1_full;             // tcgen05.commit: layer 1 of the current tile is in the TMEM stage
  uint64_t a2_full;             // 8 decode-warp arrivals: activations are back in TMEM
  uint64_t coarse_ready[kCtx];  // 8 ray-warp arrivals: coarse depths of the group are in shared memory
  uint64_t fine_ready[kCtx];    // 8 ray-warp arrivals: importance depths are in shared memory
  uint64_t csig_ready[kCtx];    // 4 decode-warp arrivals: every coarse sigma of the group is in shared memory
  uint64_t fsig_ready[kCtx];    // 4 decode-warp arrivals (every sigma of the group's last pass is in shared memory) + 1
                                // tcgen05.commit (the last layer 2, and with it every colour slot of the group, is final)
  uint64_t stg_full[2];         // expect_tx: the bulk copies of a group's jitter / u rows have landed in the staging buffer
};

// per-group shared-memory context
struct Ctx { float* dep; float* sig; float* u; float* ray; int* rk; };

struct Geom { long long ray0; int n, rstride, nr; };
// Which rays form group `grp`.  Column mode (rays are a col_w-wide image, x fastest, VR/ray_sampler.py:44):
// R vertically adjacent pixels of one image column -- they nearly share their (x,z) footprint, i.e. their taps on
// two of the three planes (VR/renderer.py:29-37).  Otherwise R consecutive rays.  32-bit arithmetic: the launcher
// checks that the group count and the rays per image fit an int (64-bit divisions cost ~100 instructions each and
// every role calls this once per job).
__device__ __forceinline__ Geom group_geom(const RenderArgs& a, unsigned grp, int R) {
  Geom g;
  const unsigned tpi = (unsigned)a.tiles_per_img;
  const unsigned n = grp / tpi, gi = grp - n * tpi;
  const bool colm = a.col_w > 0;
  const unsigned cw = colm ? (unsigned)a.col_w : 1u;
  const unsigned gy = gi / cw, gx = gi - gy * cw;
  g.n = (int)n;
  g.ray0 = (long long)n * a.rays_per_img + (colm ? (long long)(gy * R * cw + gx) : (long long)gi * R);
  g.rstride = (int)cw;
  g.nr = colm ? R : (int)min((long long)R, a.rays_per_img - (long long)gi * R);
  return g;
}

// ---------------------------------------------------------------------------------------------------------
// GATHER: one tile (rows = R rays x DPT depths) into an A1 buffer.  Warp w owns rows [8w, 8w+8).
// Step 1: lane 3s+p computes the bilinear taps of (sample s, plane p) into the warp's tap table.
// Step 2: eight lanes per sample fetch whole 128-byte texels (four channels per lane), one plane (four texels)
//         at a time: tpr_gather_microbench (profiles/) shows that on B200 a shallow queue per thread and many
//         warps sustains more random-line bandwidth than twelve loads in flight per thread.
// ---------------------------------------------------------------------------------------------------------
template <int MODE, bool TRAIN = false>
__device__ __forceinline__ void gather_tile(const RenderArgs& a, float* a1_hi, const float* __restrict__ img,
                                            const Ctx& cx, Tap2* tw, int nr, int Dx, int off, int S, int t, int dpt_shift,
                                            int warp, int lane, long long ray0 = 0, int rstride = 0) {
#ifndef TPR_GATHER_256
  const int grp = lane >> 3, sub = lane & 7;
#endif
  const int dpt = 1 << dpt_shift;
  {
    const int s = lane / 3, p = lane - s * 3;
    const int row = warp * 8 + s;
    const int r = row >> dpt_shift, di = t * dpt + (row & (dpt - 1));
    if (lane < 24 && r < nr && di < Dx) {
      const float d = cx.dep[r * S + off + di];
      const float* ry = cx.ray + r * 8;
      // origin + depth * direction (VR/renderer.py:105,123), then * 2/box_warp (:61)
      const float px = __fmul_rn(__fadd_rn(ry[0], __fmul_rn(d, ry[3])), a.box_scale);
      const float py = __fmul_rn(__fadd_rn(ry[1], __fmul_rn(d, ry[4])), a.box_scale);
      const float pz = __fmul_rn(__fadd_rn(ry[2], __fmul_rn(d, ry[5])), a.box_scale);
      Taps tp;
      plane_taps(p == 2 ? pz : px, p == 0 ? py : (p == 1 ? pz : px), a.H, a.W, tp);   // (x,y) (x,z) (z,x)
      const int po = p * a.H * a.W * kC;
      // float offsets are multiples of 32: >> 2 gives 16-byte units
      *reinterpret_cast<uint4*>(tw[lane].off) = make_uint4((unsigned)(tp.off[0] + po) >> 2, (unsigned)(tp.off[1] + po) >> 2,
                                                           (unsigned)(tp.off[2] + po) >> 2, (unsigned)(tp.off[3] + po) >> 2);
      *reinterpret_cast<float4*>(tw[lane].w2) = make_float4(tp.w[0], tp.w[0], tp.w[1], tp.w[1]);
      *reinterpret_cast<float4*>(tw[lane].w2 + 4) = make_float4(tp.w[2], tp.w[2], tp.w[3], tp.w[3]);
    }
  }
  __syncwarp();
#ifdef TPR_GATHER_256          // (measured slower in the kernel: 2.45 vs 2.30 ms at config 2, profiles/r02_ab1_*.json; kept for A/B builds)
  // Step 2, 256-bit loads: four lanes per sample, the warp's eight rows in one round (see blend_sample8)
  {
    const int s = lane >> 2, sub4 = lane & 3;
    const ulonglong2* base = reinterpret_cast<const ulonglong2*>(img) + 2 * sub4;
    asm volatile("" : "+l"(base));           // opaque to the compiler so that an address is one IMAD.WIDE
    const int row = warp * 8 + s;
    const int r = row >> dpt_shift, di = t * dpt + (row & (dpt - 1));
    if (r < nr && di < Dx) {
      // training: the summed features of the sample are kept for the backward (128 contiguous bytes per sample)
      float4* keep = (TRAIN && a.sample_features != nullptr)
          ? reinterpret_cast<float4*>(a.sample_features + ((ray0 + (long long)r * rstride) * S + off + di) * 32) + 2 * sub4 : nullptr;
      blend_sample8<MODE>(a1_hi, base, tw + s * 3, row, sub4, keep);
    }
  }
#else
  // this lane's four channels of every texel; opaque to the compiler so that an address is one IMAD.WIDE
  const ulonglong2* base = reinterpret_cast<const ulonglong2*>(img) + sub;
  asm volatile("" : "+l"(base));
#pragma unroll 1
  for (int rd = 0; rd < 2; ++rd) {
    // One warp instruction stores four rows of the operand tile.  The 128-byte swizzle XORs a row's 16-byte chunk index with
    // (row & 7), so rows that differ only in bits 0-1 put their 64 bytes into the SAME 16 banks: rows {0,1,2,3} per round was a
    // 4-way conflict on every operand store (ncu: 49.8 M conflicts of 232 M shared wavefronts).  Rows {0,1,4,5} / {2,3,6,7}
    // split a round over both bank halves: two wavefronts for 256 bytes, the minimum.  (Same-box A/B, profiles/r02_ab_rows_*:
    // neutral -- fp32 2.222 vs 2.216 ms, bf16 2.100 vs 2.105 ms: the L1 data pipe is not where the gather waits.)
#if !defined(TPR_ROWS_INTERLEAVED) || TPR_ROWS_INTERLEAVED
    const int s = (grp & 1) | ((grp >> 1) << 2) | (rd << 1);
#else
    const int s = rd * 4 + grp;
#endif
    const int row = warp * 8 + s;
    const int r = row >> dpt_shift, di = t * dpt + (row & (dpt - 1));
    if (r < nr && di < Dx) {
      const Tap2* te = tw + s * 3;
      // training: the summed features of the sample are kept for the backward (128 contiguous bytes per sample)
      float4* keep = (TRAIN && a.sample_features != nullptr)
          ? reinterpret_cast<float4*>(a.sample_features + ((ray0 + (long long)r * rstride) * S + off + di) * 32) + sub : nullptr;
      blend_sample<MODE>(a1_hi, base, te, row, sub, keep);
    }
  }
#endif
  __syncwarp();           // the tap table is rewritten by the next tile
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
// TRAIN: the composite also writes every sample's colours (and the sort every sample's sigma) to HBM for the backward
// (tpr_render_train).  A template parameter, not a run-time branch: the extra address arithmetic in the composite loop
// costs the inference kernel 3 % at its 64-register budget when it is merely predicated off.
template <int MODE, int E, int ER, bool PROF, bool TRAIN = false>
__global__ void __launch_bounds__(kThreads, 1) render_ws_kernel(const RenderArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Tiles<MODE>& tl = *reinterpret_cast<Tiles<MODE>*>(base);
  Tap2* taps = reinterpret_cast<Tap2*>(base + sizeof(Tiles<MODE>));
  float* fl = reinterpret_cast<float*>(base + sizeof(Tiles<MODE>) + sizeof(Tap2) * kGatherWarps * 24);
  const int R = a.R, Dc = a.Dc, Df = a.Df, S = Dc + Df;
  const int dpt_shift = R == 8 ? 4 : 5, dpt = 1 << dpt_shift;
  // contexts: dep [R*S], sig [R*S], u [R*Df], ray [R*8]
  const int ctx_floats = 2 * R * S + R * Df + R * 8 + (R == 4 ? R * Df : 0);      // (+ rk [R*Df]: ranks of the draws, R = 4)
  float* scratch = fl + kCtx * ctx_floats;            // ray-warp scratch: wa, wb, wc [R*S] each, rayw [R]
  auto ctx_of = [&](int gi) {
    float* p = fl + (gi & (kCtx - 1)) * ctx_floats;
    Ctx c; c.dep = p; c.sig = p + R * S; c.u = c.sig + R * S; c.ray = c.u + R * Df; c.rk = reinterpret_cast<int*>(c.ray + R * 8);
    return c;
  };
  __shared__ Barriers bars;
  __shared__ uint32_t tmem_base_sm;
  __shared__ unsigned range_sm[2];
  // Colour-slot allocation.  Slot lifetimes are not FIFO (job order C(g+1) F(g): the coarse slots of group g+1 are
  // handed out before the fine slots of group g but released after them), so the MMA issuer keeps a free mask,
  // records the slot of every tile of a group in slot_tab[ctx] (read by the ray warps for the composite) and takes
  // a group's slots back once all eight ray warps have added 1 to freed_warps after their last read of them.
  // A counter rather than an mbarrier: several groups may be released between two looks of the issuer.
  __shared__ unsigned freed_warps;
  __shared__ int slot_tab[kCtx][16];
  // TPR_PHASE_TIMING=1: cycles CTA 0 spends in each wait / work section of each role (one lane per role)
  __shared__ long long prof[24];
  const bool profiling = PROF && a.dbg != nullptr && blockIdx.x == 0;
#define PROF_T0() long long pt0_ = (PROF && profiling) ? clock64() : 0
#define PROF_ADD(i, cond) do { if (PROF && profiling && (cond)) { const long long now_ = clock64(); prof[i] += now_ - pt0_; pt0_ = now_; } } while (0)
  // the warp index goes through a shuffle so that the compiler knows it is warp-uniform: everything the MMA issuer
  // derives from role-dependent control flow (buffer index, slot, descriptors) then lives in uniform registers and a
  // tcgen05.mma costs one or two instructions instead of an elect/broadcast loop
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  // Waits poll (mbarrier.try_wait with a suspend hint).  Measured alternatives, same box (profiles/r02_ab2_*.json, r02_ab3_*.json):
  // sleeping between polls (mbar_wait_sleep) for the waits that have slack is neutral (fp32 2.339 vs 2.344 ms, bf16 2.044 vs
  // 2.050) although polls are 21 % of all executed warp instructions; handing d1_full to the eight decode warps through a
  // named barrier polled by the issuer alone is 3 % SLOWER.  TPR_WS_VARIANT & 2 selects the sleeping waits (A/B).
  const bool sleep_waits = (a.variant & 2) != 0;
#define WAIT_SLACK(bar, par) do { if (sleep_waits) mbar_wait_sleep(bar, par); else mbar_wait_parked(bar, par); } while (0)

  const int nf_host = a.Df;          // (coarse-only renders: see csig_ready below)
  if (tid == 0) {
    range_sm[0] = 0xffffffffu; range_sm[1] = 0u;
    for (int b = 0; b < kBufs; ++b) { mbar_init(&bars.a1_full[b], kGatherWarps); mbar_init(&bars.a1_free[b], 1); }
    mbar_init(&bars.d1_full, 1); mbar_init(&bars.a2_full, kDecodeWarps);
    mbar_init(&bars.stg_full[0], 1); mbar_init(&bars.stg_full[1], 1);
    for (int c = 0; c < kCtx; ++c) {
      mbar_init(&bars.coarse_ready[c], kRayWarps); mbar_init(&bars.fine_ready[c], kRayWarps);
      // coarse-only renders (nf == 0) composite straight after the coarse pass: csig then also carries the MMA completion
      mbar_init(&bars.csig_ready[c], nf_host == 0 ? 5 : 4); mbar_init(&bars.fsig_ready[c], 5);
    }
    freed_warps = 0u;
    for (int i = 0; i < 24; ++i) prof[i] = 0;
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&tmem_base_sm, 512); tmem_relinquish(); }
  for (int i = tid; i < (int)(sizeof(tl.a1) / 16); i += kThreads) reinterpret_cast<float4*>(tl.a1)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  stage_weights<MODE>(a.dec, tl);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_base_sm, 0);
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

  const int nc = (Dc + dpt - 1) >> dpt_shift, nf = Df > 0 ? (Df + dpt - 1) >> dpt_shift : 0;
  // groups of this CTA: blockIdx.x + gi * gridDim.x, gi = 0 .. G-1
  const int G = (int)((a.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const size_t img_stride = (size_t)3 * a.H * a.W * kC;
  const bool per_ray = a.rs != nullptr;

  if (warp < kGatherWarps) {
    // ====================================== GATHER ======================================
    Tap2* tw = taps + warp * 24;
    int b = 0; uint32_t ph = 0;                      // A1 buffer ring position / phase
    for (int step = 0; step <= G; ++step) {
#pragma unroll 1
      for (int jb = 0; jb < 2; ++jb) {
        const int pass = jb;
        int gi;
        if (pass == 0) { if (step >= G) continue; gi = step; }
        else { if (step < 1 || nf == 0) continue; gi = step - 1; }
        const Geom gg = group_geom(a, blockIdx.x + (unsigned)gi * gridDim.x, R);
        const Ctx cx = ctx_of(gi);
        const uint32_t cpar = (uint32_t)(gi >> 2) & 1u;
        PROF_T0();
        WAIT_SLACK(pass == 0 ? &bars.coarse_ready[gi & 3] : &bars.fine_ready[gi & 3], cpar);
        PROF_ADD(pass, tid == 0);
        const float* img = a.planes + (size_t)((unsigned)gg.n % (unsigned)a.plane_sets) * img_stride;
        const int T = pass == 0 ? nc : nf, Dx = pass == 0 ? Dc : Df, off = pass == 0 ? 0 : Dc;
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
          WAIT_SLACK(&bars.a1_free[b], ph ^ 1u);            // passes immediately the first time round
          PROF_ADD(2, tid == 0);
          gather_tile<MODE, TRAIN>(a, tl.a1[b][0], img, cx, tw, gg.nr, Dx, off, S, t, dpt_shift, warp, lane, gg.ray0, gg.rstride);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.a1_full[b]);
          PROF_ADD(3, tid == 0);
          if (++b == kBufs) { b = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp < kFirstRayWarp) {
    // ====================================== DECODE ======================================
    const int dw = warp - kFirstDecodeWarp, q = dw & 3, h = dw >> 2;
    const bool issuer = dw == 0;
    uint32_t pt = 0;                                 // per-tile phase of d1_full / a2_full (and the psig buffer)
    const uint32_t dbase = smem_desc_lo(smem_u32(&tl));
    uint32_t free_mask = (1u << Cols<MODE>::ns) - 1u;  // issuer only (warp-uniform): free colour slots
    int freed_groups = 0;                              // groups whose slots have been taken back
    int b = 0; uint32_t ph = 0;                      // A1 buffer ring
    for (int step = 0; step <= G; ++step) {
#pragma unroll 1
      for (int jb = 0; jb < 2; ++jb) {
        const int pass = jb;
        int gi;
        if (pass == 0) { if (step >= G) continue; gi = step; }
        else { if (step < 1 || nf == 0) continue; gi = step - 1; }
        const Geom gg = group_geom(a, blockIdx.x + (unsigned)gi * gridDim.x, R);
        const Ctx cx = ctx_of(gi);
        const int T = pass == 0 ? nc : nf, Dx = pass == 0 ? Dc : Df, off = pass == 0 ? 0 : Dc;
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
          PROF_T0();
          const bool pl = dw == 0 && lane == 0;
          const uint32_t st = tmem;                        // the layer-1 stage: columns [0, 64)
          if (issuer) {
            mbar_wait_parked(&bars.a1_full[b], ph);
            PROF_ADD(4, pl);
            tcgen05_fence_after();
            const int bu = __shfl_sync(0xffffffffu, b, 0);     // uniform register for the descriptor arithmetic
            if (elect_one_sync()) {
              issue_layer1<MODE>(dbase, bu, st);
              mma_commit(&bars.a1_free[b]);
              mma_commit(&bars.d1_full);
            }
            __syncwarp();
          }
          if (++b == kBufs) { b = 0; ph ^= 1u; }
          int slot = 0;
          if (issuer) {
            // take back the slots of every group composited since the last look (eagerly: slot_tab[ctx] is rewritten
            // four groups later); spin only while every slot holds colours of a group that is not composited yet
            do {
              const int fg = (int)(*reinterpret_cast<volatile unsigned*>(&freed_warps) / kRayWarps);
              for (; freed_groups < fg; ++freed_groups)
                for (int i = 0; i < nc + nf; ++i) free_mask |= 1u << slot_tab[freed_groups & (kCtx - 1)][i];
            } while (free_mask == 0u);
            __threadfence_block();
            slot = __shfl_sync(0xffffffffu, __ffs(free_mask) - 1, 0);        // (warp-uniform by construction)
            free_mask &= ~(1u << slot);
            if (lane == 0) slot_tab[gi & (kCtx - 1)][(pass == 0 ? 0 : nc) + t] = slot;
          }
          PROF_ADD(5, pl);
          mbar_wait_parked(&bars.d1_full, pt);               // also: layer 2 of the previous tile has consumed the activations
          PROF_ADD(6, pl);
          tcgen05_fence_after();
          const float sgp = epilogue1<MODE>(tl, st, lane_base, h);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.a2_full);
          PROF_ADD(7, pl);
          if (issuer) {
            mbar_wait_parked(&bars.a2_full, pt);
            PROF_ADD(8, pl);
            tcgen05_fence_after();
            if (elect_one_sync()) {
              issue_layer2<MODE>(dbase, st, tmem + Cols<MODE>::slots + slot * kSlotCols);
              // Last tile of the group's last pass: the composite needs every colour slot of the group, i.e. this layer 2
              // (and with it all earlier MMAs) complete.  The commit arrives on the group's barrier itself, so no decode
              // warp -- least of all this one, which issues the next job's MMAs -- ever waits for the tensor pipe to drain.
              if (t == T - 1 && (pass == 1 || nf == 0)) mma_commit(pass == 0 ? &bars.csig_ready[gi & 3] : &bars.fsig_ready[gi & 3]);
            }
            __syncwarp();
          }
          PROF_ADD(9, pl);
          // sigma of this tile -> shared memory: the two warps of a lane quarter add their halves.  psig is double buffered
          // by tile parity: the h == 1 warp of tile t+2 can only get here after the epilogue barrier (a2_full) of tile t+1,
          // which the h == 0 warp arrives at after it has read tile t's partial sums -- one named barrier per tile is enough.
          float* ps = tl.psig + (pt ? kRows : 0);
          if (h == 1) ps[q * 32 + lane] = sgp;
          named_bar_sync(3 + q, 64);
          if (h == 0) {
            const int row = q * 32 + lane, r = row >> dpt_shift, di = t * dpt + (row & (dpt - 1));
            if (r < gg.nr && di < Dx) cx.sig[r * S + off + di] = sgp + ps[row] + tl.bias2[kNc];
          }
          PROF_ADD(10, pl);
          if (h == 0 && t == T - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(pass == 0 ? &bars.csig_ready[gi & 3] : &bars.fsig_ready[gi & 3]);
          }
          pt ^= 1u;
          PROF_ADD(11, pl);
        }
      }
    }
  } else {
    // ====================================== RAYS ======================================
    const int rw = warp - kFirstRayWarp, rtid = tid - kFirstRayWarp * 32;
    const int q = rw & 3, hc = rw >> 2;               // composite: lane quarter, channel half
    constexpr int kRayThreads = kRayWarps * 32;
    float* wa = scratch; float* wb = wa + R * S; float* wc = wb + R * S; float* rayw = wc + R * S;
    float mn = __int_as_float(0x7f800000), mx = -__int_as_float(0x7f800000);
    float smn = mn, smx = mx;                         // running depth range of the current clamp slot
    int cur_slot = 0;
#define RAY_SYNC() named_bar_sync(2, kRayThreads)

    const bool pl = rtid == 0;
    // The group's per-ray inputs (6 ray floats, Dc jitter draws, Df uniform draws per ray) are DRAM reads with
    // nothing to overlap them inside setup, so they are fetched with cp.async into a staging area one step ahead;
    // each thread later converts exactly the elements it copied itself (no barrier needed, only wait_group).
    // Jitter / u rows, optional path (a.bulk_inputs, opt-in: measured 2 % slower at config 2): one cp.async.bulk per ray and
    // array (192-byte rows at 48 + 48), issued by one thread, completion counted on an mbarrier, staging double buffered (any
    // thread may then convert any element); needs 16-byte aligned rows.  Default, and always for the 12-byte origin /
    // direction rows: 4-byte cp.async, each thread converting the elements it copied itself.
    const bool bulk = a.bulk_inputs != 0;
    float* stg_jit = rayw + R; float* stg_u = stg_jit + 2 * R * Dc; float* stg_ray = stg_u + 2 * R * Df;
    int* hist = reinterpret_cast<int*>(stg_ray + R * 8 + 8);      // [R][Dc + 4] (R = 4: warp_merge_scatter)
    // R = 4 (more than 64 samples per pass): a ray's samples are merged instead of rank-counted (tpr_render.cuh)
    const bool merge = R == 4 && nf > 0 && (Df & 3) == 0 && a.variant != 3;
    // R = 4, SPLIT steps: a group has four rays but there are eight ray warps, and the ray warps' serial timeline (resample
    // 10 k + merge 8.6 k + march 5.4 k + composite 8.7 k cycles per group at 96+96) was the critical path of the whole kernel
    // (phase counters, profiles/r02_phase96_*.txt: gather and decode waited for it).  So warps 0-3 resample group g (and rank
    // its draws) WHILE warps 4-7 sort and march group g-1, each on its own scratch rows; all eight then composite g-1.
    const bool split = merge && a.noise_c == nullptr && (a.variant & 16) == 0;
    float* so_a = split ? reinterpret_cast<float*>(hist + R * (Dc + 4)) : wa;      // omega rows of the concurrent sort
    float* so_b = split ? so_a + R * S : wb;                                         // its 2*S scratch floats per ray
    auto prefetch = [&](int gi) {
      const Geom gg = group_geom(a, blockIdx.x + (unsigned)gi * gridDim.x, R);
      if (rtid < gg.nr * 6) {
        const int r = rtid / 6, c = rtid - r * 6;
        const long long g = gg.ray0 + (long long)r * gg.rstride;
        cp_async4(stg_ray + rtid, c < 3 ? a.origins + g * 3 + c : a.dirs + g * 3 + c - 3);
      }
      float* sj = stg_jit + (gi & 1) * R * Dc; float* su = stg_u + (gi & 1) * R * Df;
      if (bulk) {
        if (rtid == 0) {
          mbar_arrive_expect_tx(&bars.stg_full[gi & 1], (uint32_t)(gg.nr * (Dc + Df)) * 4u);
          for (int r = 0; r < gg.nr; ++r) {
            const long long g = gg.ray0 + (long long)r * gg.rstride;
            bulk_copy_g2s(sj + r * Dc, a.jitter + g * Dc, (uint32_t)Dc * 4u, &bars.stg_full[gi & 1]);
            if (Df > 0) bulk_copy_g2s(su + r * Df, a.u + g * Df, (uint32_t)Df * 4u, &bars.stg_full[gi & 1]);
          }
        }
      } else {
        for (int s = rtid; s < gg.nr * Dc; s += kRayThreads) {
          const int r = s / Dc, k = s - r * Dc;
          cp_async4(sj + s, a.jitter + (gg.ray0 + (long long)r * gg.rstride) * Dc + k);
        }
        for (int s = rtid; s < gg.nr * Df; s += kRayThreads) {
          const int r = s / Df, k = s - r * Df;
          cp_async4(su + s, a.u + (gg.ray0 + (long long)r * gg.rstride) * Df + k);
        }
      }
      cp_async_commit();
    };
    auto setup = [&](int gi) {
      // rays, coarse depths (VR/renderer.py:169-192) and the group's uniform draws into its context
      PROF_T0();
      const Geom gg = group_geom(a, blockIdx.x + (unsigned)gi * gridDim.x, R);
      const Ctx cx = ctx_of(gi);
      cp_async_wait_all();
      if (bulk) mbar_wait_parked(&bars.stg_full[gi & 1], (uint32_t)(gi >> 1) & 1u);
      const float* sj = stg_jit + (gi & 1) * R * Dc; const float* su = stg_u + (gi & 1) * R * Df;
      if (rtid < gg.nr * 6) { const int r = rtid / 6; cx.ray[r * 8 + (rtid - r * 6)] = stg_ray[rtid]; }
      for (int s = rtid; s < gg.nr * Dc; s += kRayThreads) {
        const int r = s / Dc, k = s - r * Dc;
        const long long g = gg.ray0 + (long long)r * gg.rstride;
        const float lo = per_ray ? __ldg(a.rs + g) : a.ray_start, hi = per_ray ? __ldg(a.re + g) : a.ray_end;
        cx.dep[r * S + k] = coarse_depth(a, k, sj[s], lo, hi, per_ray);
      }
      for (int s = rtid; s < gg.nr * Df; s += kRayThreads) cx.u[s] = su[s];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.coarse_ready[gi & 3]);
      PROF_ADD(12, pl);
    };

    auto resample = [&](int gi) {
      const Geom gg = group_geom(a, blockIdx.x + (unsigned)gi * gridDim.x, R);
      const Ctx cx = ctx_of(gi);
      PROF_T0();
      WAIT_SLACK(&bars.csig_ready[gi & 3], (uint32_t)(gi >> 2) & 1u);
      PROF_ADD(13, pl);
      if (a.noise_c != nullptr) {     // density_noise (VR/renderer.py:146), coarse pass
        add_density_noise(a, a.noise_c, cx.sig, S, 0, Dc, gg.nr * Dc, gg.ray0, gg.rstride, rtid, kRayThreads);
        RAY_SYNC();
      }
      // the uniform draws were staged by other warps in setup(); coarse_ready has completed (the coarse pass ran)
      for (int r = rw; r < gg.nr; r += kRayWarps)
        warp_resample_ray(a, cx.dep + r * S, cx.sig + r * S, wa + r * S, wb + r * S, wc + r * S, cx.dep + r * S + Dc,
                          gg.ray0 + (long long)r * gg.rstride, lane, cx.u + r * Df);
      // R = 4: warps 4-7 have no ray to resample; they rank the group's uniform draws for the merge of the next step
      if (!split && merge && rw >= 4 && rw - 4 < gg.nr) warp_rank_draws(cx.u + (rw - 4) * Df, cx.rk + (rw - 4) * Df, Df, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.fine_ready[gi & 3]);
      // split steps: the resampling warp ranks its own ray's draws, after the importance depths have been handed to the gather
      // (the sorting warps doing it instead -- TPR_WS_VARIANT & 32 -- is +0.6 % in the 2xFP16 mode, -3.5 % in the bf16 mode)
      if (split && (a.variant & 32) == 0 && rw < gg.nr) warp_rank_draws(cx.u + rw * Df, cx.rk + rw * Df, Df, lane);
      PROF_ADD(14, pl);
    };

    // split steps: sort + final march of ray rw - 4 of group gi by warp rw >= 4 alone (merge, else the single-warp rank count)
    auto sort_split = [&](int gi) {
      const Geom gg = group_geom(a, blockIdx.x + (unsigned)gi * gridDim.x, R);
      const Ctx cx = ctx_of(gi);
      PROF_T0();
      mbar_wait_parked(&bars.fsig_ready[gi & 3], (uint32_t)(gi >> 2) & 1u);
      PROF_ADD(15, pl);
      if (range_slot(a, gg.n) != cur_slot) { range_fold(a, cur_slot, smn, smx, mn, mx, lane); cur_slot = range_slot(a, gg.n); }
      const int r = rw - 4;
      if (r < gg.nr) {
        float wsum, dnum;
        bool pre = false;
        if (cx.rk[r * Df] >= 0)
          pre = warp_merge_scatter(cx.dep + r * S, cx.sig + r * S, cx.rk + r * Df, so_a + r * S, so_b + r * 2 * S, hist + r * (Dc + 4),
                                   S, Dc, lane);
        warp_sort_and_weights<E, true, ER>(cx.dep + r * S, cx.sig + r * S, so_a + r * S, nullptr, S, lane, wsum, dnum, smn, smx,
                                           so_b + r * 2 * S, pre);
        if (lane == 0) {
          const long long g = gg.ray0 + (long long)r * gg.rstride;
          rayw[r] = wsum;
          const float dq = dnum / wsum;               // NaN -> inf and the clamp happen in finish_kernel
          a.depth[g] = dq;
          a.wsum[g] = wsum;
          for (int p = 0; p < a.peers.n; ++p) { a.peers.depth[p][g] = dq; a.peers.wsum[p][g] = wsum; }   // NVLink stores
        }
      }
      PROF_ADD(18, pl);
    };

    // `sorted` = true: the group has been sorted already (split steps); the composite reads omega from `om`
    auto sort_composite = [&](int gi, bool sorted, const float* om_rows) {
      const Geom gg = group_geom(a, blockIdx.x + (unsigned)gi * gridDim.x, R);
      const Ctx cx = ctx_of(gi);
      PROF_T0();
      WAIT_SLACK(nf > 0 ? &bars.fsig_ready[gi & 3] : &bars.csig_ready[gi & 3], (uint32_t)(gi >> 2) & 1u);
      PROF_ADD(15, pl);
      if (a.noise_c != nullptr) {     // density_noise (VR/renderer.py:146): the pass whose sigma has just arrived
        if (nf > 0) add_density_noise(a, a.noise_f, cx.sig, S, Dc, Df, gg.nr * Df, gg.ray0, gg.rstride, rtid, kRayThreads);
        else add_density_noise(a, a.noise_c, cx.sig, S, 0, Dc, gg.nr * Dc, gg.ray0, gg.rstride, rtid, kRayThreads);
        RAY_SYNC();
      }
      if (range_slot(a, gg.n) != cur_slot) { range_fold(a, cur_slot, smn, smx, mn, mx, lane); cur_slot = range_slot(a, gg.n); }
      // ---- sort + final march: omega per sample (scattered to original order), depth, weight sum
      // R = 4: two warps per ray share the rank count (pair_rank_scatter); warp rw < 4 then runs the march alone
      const bool pairs = R == 4 && nf > 0 && (Dc & 31) == 0 && (Df & 7) == 0;
      for (int r = (pairs || merge) ? (rw & 3) : rw; !sorted && r < gg.nr; r += kRayWarps) {
        float wsum, dnum;
        bool pre = false;
        const bool mg = merge && cx.rk[r * Df] >= 0;   // (both warps of the ray read the same verdict on the draws)
        if (mg) {
          if (rw >= 4) break;
          pre = warp_merge_scatter(cx.dep + r * S, cx.sig + r * S, cx.rk + r * Df, wa + r * S, wb + r * 2 * S, hist + r * (Dc + 4),
                                   S, Dc, lane);
          PROF_ADD(18, pl);
        } else if (!pairs && merge && rw >= 4) {
          break;
        } else if (pairs) {
          pre = pair_rank_scatter<ER>(cx.dep + r * S, cx.sig + r * S, wa + r * S, wb + r * 2 * S, S, Dc, lane, rw >> 2, 7 + r);
          if (rw >= 4) break;
          PROF_ADD(18, pl);                            // (profiling build: the rank count's share of sort+march)
        }
        warp_sort_and_weights<E, true, ER>(cx.dep + r * S, cx.sig + r * S, wa + r * S, nullptr, S, lane, wsum, dnum, smn, smx,
                                       wb + r * 2 * S, pre);     // wb and wc are contiguous: 2*S floats per ray
        if (TRAIN && a.sample_sigma != nullptr) {    // training: the backward reads sigma of every sample instead of recomputing it
          float* dst = a.sample_sigma + (gg.ray0 + (long long)r * gg.rstride) * S;
          for (int p = lane; p < S; p += 32) dst[p] = cx.sig[r * S + p];
        }
        if (lane == 0) {
          const long long g = gg.ray0 + (long long)r * gg.rstride;
          rayw[r] = wsum;
          const float dq = dnum / wsum;               // NaN -> inf and the clamp happen in finish_kernel
          a.depth[g] = dq;
          a.wsum[g] = wsum;
          for (int p = 0; p < a.peers.n; ++p) { a.peers.depth[p][g] = dq; a.peers.wsum[p][g] = wsum; }   // NVLink stores
        }
      }
      RAY_SYNC();
      PROF_ADD(16, pl);
      // ---- composite: ray warp (q, hc) sums channels [16hc, 16hc+16) over the samples held by its lanes
      tcgen05_fence_after();
      const int row = q * 32 + lane, r = row >> dpt_shift, i = row & (dpt - 1);
      uint64_t acc2[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc2[c] = 0ull;
      const uint64_t kOne2 = pack2(1.0f, 1.0f), kScale2 = pack2(1.002f, 1.002f), kShift2 = pack2(-0.001f, -0.001f);
#pragma unroll 1
      for (int sl = 0; sl < nc + nf; ++sl) {
        const bool fine = sl >= nc;
        const int tl_i = fine ? sl - nc : sl;
        const int slot = slot_tab[gi & (kCtx - 1)][sl];
        const int di = tl_i * dpt + i;
        const bool valid = r < gg.nr && di < (fine ? Df : Dc);
        const float om = valid ? om_rows[r * S + (fine ? Dc : 0) + di] : 0.0f;
        uint32_t v[16];
        tmem_ld16(tmem + Cols<MODE>::slots + slot * kSlotCols + lane_base + 16 * hc, v);
        tmem_wait_ld();
        // rows outside the group (om = 0) still hold finite values: the operand tiles start zeroed and only ever
        // receive finite features, so no select is needed to keep NaNs out of the sum
        const uint64_t om2 = pack2(om, om);
        // training: this lane's sixteen colours of the sample also go to HBM -- what tpr_render_backward reads instead of
        // re-running the decoder.  Layout [ray][8 chunks of 4 channels][S samples][4 floats] (samples in the forward's order,
        // coarse then importance): the lanes of a warp are consecutive samples of a ray, so one store instruction writes
        // 16 B x 16 lanes = 256 contiguous bytes per ray.  (Sample-major rows -- 64 B per lane, 128 B apart -- cost the
        // training forward 0.62 ms at config 2: 32 partial lines per store instruction.)
        ulonglong2* cdst = (TRAIN && a.sample_colours != nullptr && valid)
            ? reinterpret_cast<ulonglong2*>(a.sample_colours + ((gg.ray0 + (long long)r * gg.rstride) * 8 + 4 * hc) * (long long)S * 4) +
                  ((fine ? Dc : 0) + di)
            : nullptr;
        const int cstep = S;                       // 16-byte units between two chunks of a ray
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const ulonglong2 bz = *reinterpret_cast<const ulonglong2*>(tl.bias2 + 16 * hc + 4 * c4);
          ulonglong2 cpair;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int c = 4 * c4 + 2 * h2;
            float z0, z1;
            unpack2(add2(pack2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), h2 == 0 ? bz.x : bz.y), z0, z1);
            // sigmoid(x) * 1.002 - 0.001 with z = -x * log2e (training/triplane.py:134)
            const uint64_t den = add2(pack2(ex2_fast(z0), ex2_fast(z1)), kOne2);
            float d0, d1;
            unpack2(den, d0, d1);
            const uint64_t col = fma2(pack2(rcp_fast(d0), rcp_fast(d1)), kScale2, kShift2);
            acc2[c >> 1] = fma2(om2, col, acc2[c >> 1]);
            if (h2 == 0) cpair.x = col; else cpair.y = col;
          }
          if (TRAIN && cdst != nullptr) cdst[c4 * cstep] = cpair;
        }
      }
      float acc[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) unpack2(acc2[c], acc[2 * c], acc[2 * c + 1]);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { __threadfence_block(); atomicAdd(&freed_warps, 1u); }   // this warp is done with the group's slots
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        if (o < dpt) {
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] += __shfl_xor_sync(kFull, acc[c], o);
        }
      }
      if (i == 0 && r < gg.nr) {
        const float wsr = rayw[r];
        long long cstride;
        float* dst1 = rgb_ptr(a, gg.ray0 + (long long)r * gg.rstride, gg.n, cstride);
        dst1 += 16 * hc * cstride;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float v = acc[c];
          if (a.white_back) v = v + 1.0f - wsr;        // VR/ray_marcher.py:52-53
          acc[c] = v * 2.0f - 1.0f;                    // :55
        }
        // this GPU's buffer first, then the same element of every peer's gather buffer (peer-mapped pointers: the
        // stores travel over NVLink while the other warps of the SM keep rendering)
        const long long eoff = dst1 - a.rgb;
#pragma unroll 1
        for (int p = -1; p < a.peers.n; ++p) {
          float* d1 = p < 0 ? dst1 : a.peers.rgb[p] + eoff;
          if (a.nchw) {
#pragma unroll
            for (int c = 0; c < 16; ++c) d1[c * cstride] = acc[c];
          } else {
            float4* dst = reinterpret_cast<float4*>(d1);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) dst[c4] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
          }
        }
      }
      RAY_SYNC();                                      // wa / rayw are reused by the next resample / sort
      PROF_ADD(17, pl);
    };

    prefetch(0); setup(0);
    if (G > 1) { prefetch(1); setup(1); }
    if (G > 2) prefetch(2);
    for (int g = 0; g < G; ++g) {
      if (split) {
        if (rw < 4) resample(g);
        else {
          if (lane == 0) mbar_arrive(&bars.fine_ready[g & 3]);       // (nothing to contribute: the four resampling warps complete it)
          if (g >= 1) sort_split(g - 1);
          if ((a.variant & 32) != 0) {                                // (A/B) rank the draws of group g for its merge in the next step
            const Geom gg = group_geom(a, blockIdx.x + (unsigned)g * gridDim.x, R);
            const Ctx cx = ctx_of(g);
            if (rw - 4 < gg.nr) warp_rank_draws(cx.u + (rw - 4) * Df, cx.rk + (rw - 4) * Df, Df, lane);
          }
        }
        RAY_SYNC();
      } else if (nf > 0) {
        resample(g);
      }
      if (g + 2 < G) { setup(g + 2); if (g + 3 < G) prefetch(g + 3); }
      if (nf > 0) { if (g >= 1) sort_composite(g - 1, split, so_a); }
      else sort_composite(g, false, wa);
    }
    if (nf > 0) {
      if (split) { if (rw >= 4) sort_split(G - 1); RAY_SYNC(); }
      sort_composite(G - 1, split, so_a);
    }
    range_fold(a, cur_slot, smn, smx, mn, mx, lane);
    mn = warp_min(mn); mx = warp_max(mx);
    if (lane == 0 && mn <= mx) {
      atomicMin(&range_sm[0], float_to_ordered(mn));
      atomicMax(&range_sm[1], float_to_ordered(mx));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0 && range_sm[0] <= range_sm[1]) {
    atomicMin(a.range_enc + 0, range_sm[0]);
    atomicMax(a.range_enc + 1, range_sm[1]);
  }
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
  if (PROF && profiling && tid < 24) a.dbg[tid] = prof[tid];
}

template <int MODE>
static size_t smem_bytes(int R, int S, int Df) {
  return 1024 + sizeof(Tiles<MODE>) + sizeof(Tap2) * kGatherWarps * 24 +
         sizeof(float) * ((size_t)kCtx * (2 * R * S + R * Df + R * 8 + (R == 4 ? R * Df : 0)) + (size_t)3 * R * S + R +
                          (size_t)2 * R * S + R * 8 + 8 + (R == 4 ? R * (S - Df + 4) + (size_t)3 * R * S : 0));
}

typedef void (*Kernel)(const RenderArgs);
// The instantiations live in four translation units (tpr_render_ws_{f16x2,bf16}_{a,b}.cu), one per decoder mode and
// sample-count class, so that they compile in parallel: kernel_small<MODE> covers S <= 96 (and the training / profiling
// variants of 48+48), kernel_large<MODE> the rest.
template <int MODE> Kernel kernel_small(int S, bool prof, bool train);
template <int MODE> Kernel kernel_large(int S, bool prof);

}  // namespace ws
}  // namespace tpr
