// Pieces shared by the two fused render kernels (FFMA decoder: triplane_b200.cu, tensor-core
// decoder: tpr_render_tc.cu).  Citations relative to /root/reference/g_nerf/ (VR/ = training/volumetric_rendering/).
#pragma once
#include "tpr_device.cuh"

namespace tpr {

// Multi-GPU: where this call's outputs are ALSO stored -- the same slice of every peer GPU's gather buffers, through
// peer-mapped (NVLink) device pointers.  The render kernel's epilogue writes the 136 bytes per ray to every sink, so
// the "all-gather" of SURVEY.md section 8(e) is part of the render and overlaps it ray group by ray group.
constexpr int kMaxPeers = 15;
struct PeerSinks { int n; float* rgb[kMaxPeers]; float* depth[kMaxPeers]; float* wsum[kMaxPeers]; };

struct RenderArgs {
  const float* planes; int H, W;
  const float* dec;
  const float* origins; const float* dirs;
  const float* jitter; const float* u;
  const float* rs; const float* re;      // optional per-ray limits
  long long n_rays_total, rays_per_img, n_tiles, tiles_per_img;
  float ray_start, ray_end, box_scale, lin_step, jitter_scale, inv_start, inv_end;
  int Dc, Df, disparity, white_back, R;
  float* rgb; float* depth; float* wsum; float* fine_depths; int* fine_inds;
  unsigned* range_enc;                   // [2]: ordered-uint encoded (min, max) of all depths
  int variant;                           // debug: A/B switches (TPR_WS_VARIANT)
  int bulk_inputs;                       // jitter / u rows are 16-byte aligned multiples of 16 bytes: fetch them with cp.async.bulk
  int col_w;                             // > 0: rays form an image col_w pixels wide and a group is R rays of one image COLUMN
  long long* dbg;                        // optional [16] per-phase cycle counters of CTA 0 (TPR_PHASE_TIMING=1)
  int plane_sets;                        // >= 1: image (camera) n samples plane set n % plane_sets
  int nchw;                              // != 0: rgb is written channels-first, [N,32,M] (training/triplane.py:81)
  int clamp_group;                       // k > 0: images [j*k, (j+1)*k) share a depth-clamp range (slot j); 0: one range
  PeerSinks peers;                       // n > 0: every output store is repeated into these peer buffers
  float* sample_colours; float* sample_sigma;   // training: per-sample colours [rays,S,32] / sigma [rays,S] kept for the backward
  float* sample_features;                       // training: per-sample summed plane features [rays,S,32] (optional)
  // density_noise > 0 (VR/renderer.py:146): standard-normal draws [rays,Dc] / [rays,Df], added to sigma times density_noise
  const float* noise_c; const float* noise_f; float density_noise;
};

// sigma += randn * density_noise (VR/renderer.py:146) for the `count` = nr * Dx samples of a ray group's pass, by `nthreads`
// cooperating threads; sig rows are S-strided, the noise is laid out like the pass's sigma tensor [N, M*Dx, 1].
__device__ __forceinline__ void add_density_noise(const RenderArgs& a, const float* __restrict__ noise, float* sig, int S, int off,
                                                  int Dx, int count, long long ray0, int rstride, int tid, int nthreads) {
  for (int s = tid; s < count; s += nthreads) {
    const int r = s / Dx, k = s - r * Dx;
    const long long g = ray0 + (long long)r * rstride;
    sig[r * S + off + k] = __fadd_rn(sig[r * S + off + k], __fmul_rn(__ldg(noise + g * Dx + k), a.density_noise));
  }
}

// Depth ranges in the scratch block (unsigned words, ordered-uint encoded): [0..1] whole call, then from
// kRangeSlotOff one (min, max) pair per clamp slot.  Bytes 64..191 hold the optional phase counters.
constexpr int kRangeSlotOff = 64;
__device__ __forceinline__ int range_slot(const RenderArgs& a, int n) { return a.clamp_group > 0 ? n / a.clamp_group : 0; }
// Fold the running range of slot `slot` (smn, smx; per-lane partials) into the call-wide running range (mn, mx) and,
// with per-slot clamping on, into the slot's words; then restart it.  Warp-collective.
__device__ __forceinline__ void range_fold(const RenderArgs& a, int slot, float& smn, float& smx, float& mn, float& mx, int lane) {
  mn = fminf(mn, smn); mx = fmaxf(mx, smx);
  if (a.clamp_group > 0) {
    const float lo = warp_min(smn), hi = warp_max(smx);
    if (lane == 0 && lo <= hi) {
      atomicMin(a.range_enc + kRangeSlotOff + 2 * slot, float_to_ordered(lo));
      atomicMax(a.range_enc + kRangeSlotOff + 2 * slot + 1, float_to_ordered(hi));
    }
  }
  smn = __int_as_float(0x7f800000); smx = -__int_as_float(0x7f800000);
}

// this ray's first rgb element and the stride between its channels, for either output layout
__device__ __forceinline__ float* rgb_ptr(const RenderArgs& a, long long g, int n, long long& cstride) {
  if (a.nchw) { cstride = a.rays_per_img; return a.rgb + (long long)n * (kC - 1) * a.rays_per_img + g; }   // n*32*M + (g - n*M)
  cstride = 1;
  return a.rgb + g * kC;
}

constexpr int kRenderMaxThreads = 512;

// coarse depth k of a ray (VR/renderer.py:169-192)
__device__ __forceinline__ float coarse_depth(const RenderArgs& a, int k, float jit, float rs, float re, bool per_ray) {
  const int D = a.Dc;
  if (a.disparity) {                      // :174-181
    const float step = 1.0f / (float)(D - 1);
    float t = (k < D / 2) ? __fmul_rn(step, (float)k) : __fsub_rn(1.0f, __fmul_rn(step, (float)(D - 1 - k)));
    t = __fadd_rn(t, __fmul_rn(jit, step));
    float lo = __fmul_rn(a.inv_start, __fsub_rn(1.0f, t));
    float hi = __fmul_rn(a.inv_end, t);
    return __fdiv_rn(1.0f, __fadd_rn(lo, hi));
  }
  if (per_ray) {                          // :183-186 with math_utils.linspace (math_utils.py:101-118)
    float steps = __fdiv_rn((float)k, (float)(D - 1));
    float base = __fadd_rn(rs, __fmul_rn(steps, __fsub_rn(re, rs)));
    float delta = __fdiv_rn(__fsub_rn(re, rs), (float)(D - 1));
    return __fadd_rn(base, __fmul_rn(jit, delta));
  }
  // :188-190 with torch.linspace's two-sided formula
  float base = (k < D / 2) ? __fadd_rn(a.ray_start, __fmul_rn(a.lin_step, (float)k))
                           : __fsub_rn(a.ray_end, __fmul_rn(a.lin_step, (float)(D - 1 - k)));
  return __fadd_rn(base, __fmul_rn(jit, a.jitter_scale));
}




// a12 + a9 for the final pass: sort a ray's S samples by depth (VR/renderer.py:157-167), run the
// march (VR/ray_marcher.py:26-46) and emit, per sample, the weight its colour carries in the composite:
//   rgb = sum_i w_i (c_i + c_{i+1})/2 = sum_p c_p * omega_p,  omega_p = (w_{p-1} + w_p)/2.
// kScatter = false: om[p], oi[p] by sorted position p (oi = original index);
// kScatter = true : om[original index] = omega (oi unused).
// Also returns sum(w), sum(w * mid depth) and folds the ray's depth range into (mn, mx).
// Sorting: with `tmp` (2*S floats of scratch; `om` serves as a third scratch row) each lane counts, for its elements,
// how many of the ray's S depths are smaller -- S warp-broadcast shared-memory reads and S*E independent compares,
// no shuffles, so it overlaps with itself far better than the bitonic network's dependent compare-exchange chain
// (the ray warps are latency bound) -- and scatters (depth, sigma, index) to that rank.  Two equal depths would
// get the same rank; a rank-sum check detects that and falls back to the network with its index tie-break.
// ER: rows of 32 elements the rank count looks at (ceil(S/32) <= ER <= E; E itself must be a power of two for the network).
// `pre_ranked`: tmp / om already hold the ray sorted (pair_rank_scatter below did the first half of this function).
template <int E, bool kScatter, int ER = E>
__device__ __forceinline__ void warp_sort_and_weights(const float* z, const float* sg, float* om, int* oi, int S, int lane,
                                                      float& wsum_out, float& dnum_out, float& mn, float& mx,
                                                      float* tmp = nullptr, bool pre_ranked = false) {
  float key[E]; int idx[E];
  float sgm[E];
  bool ranked = false;
  if (pre_ranked) {
    const float* zs = tmp; const float* ss = tmp + S; const int* is = reinterpret_cast<const int*>(om);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = lane * E + e;
      key[e] = p < S ? zs[p] : __int_as_float(0x7f800000);
      sgm[e] = p < S ? ss[p] : 0.0f;
      idx[e] = p < S ? is[p] : p;
    }
    __syncwarp();                                // `om` is rewritten below
    ranked = true;
  } else if (tmp != nullptr) {
    // element p = e*32 + lane (strided: conflict-free reads and writes)
    float ze[ER]; int cnt[ER];
#pragma unroll
    for (int e = 0; e < ER; ++e) { const int p = e * 32 + lane; ze[e] = p < S ? z[p] : __int_as_float(0x7f800000); cnt[e] = 0; }
    int j = 0;
    if ((reinterpret_cast<uintptr_t>(z) & 15) == 0) {
#pragma unroll 2
      for (; j + 4 <= S; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(z + j);
#pragma unroll
        for (int e = 0; e < ER; ++e)
          cnt[e] += (int)(v.x < ze[e]) + (int)(v.y < ze[e]) + (int)(v.z < ze[e]) + (int)(v.w < ze[e]);
      }
    }
    for (; j < S; ++j) {
      const float v = z[j];
#pragma unroll
      for (int e = 0; e < ER; ++e) cnt[e] += (int)(v < ze[e]);
    }
    int rsum = 0;
#pragma unroll
    for (int e = 0; e < ER; ++e) if (e * 32 + lane < S) rsum += cnt[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(kFull, rsum, o);
    ranked = rsum == S * (S - 1) / 2;            // a tie (or a NaN) makes the sum fall short
    if (ranked) {
      float* zs = tmp; float* ss = tmp + S; int* is = reinterpret_cast<int*>(om);
#pragma unroll
      for (int e = 0; e < ER; ++e) {
        const int p = e * 32 + lane;
        if (p < S) { zs[cnt[e]] = ze[e]; ss[cnt[e]] = sg[p]; is[cnt[e]] = p; }
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int p = lane * E + e;
        key[e] = p < S ? zs[p] : __int_as_float(0x7f800000);
        sgm[e] = p < S ? ss[p] : 0.0f;
        idx[e] = p < S ? is[p] : p;
      }
      __syncwarp();                              // `om` is rewritten below
    }
  }
  if (!ranked) {
    bool sorted = true;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      int p = lane * E + e;
      key[e] = p < S ? z[p] : __int_as_float(0x7f800000);
      idx[e] = p;
      if (e > 0) sorted &= !(key[e] < key[e - 1]);
    }
    {
      float prev = __shfl_up_sync(kFull, key[E - 1], 1);
      if (lane > 0) sorted &= !(key[0] < prev);
    }
    if (!__all_sync(kFull, sorted)) warp_bitonic_sort<E>(key, idx, lane);
    // sorted, blocked: position p = lane*E + e
#pragma unroll
    for (int e = 0; e < E; ++e) sgm[e] = (lane * E + e) < S ? sg[idx[e]] : 0.0f;
  }
  const float nk = __shfl_down_sync(kFull, key[0], 1), ns = __shfl_down_sync(kFull, sgm[0], 1);
  float al[E], dm[E];
  float prod = 1.0f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = lane * E + e;
    const float d1 = e + 1 < E ? key[(e + 1) % E] : nk, s1 = e + 1 < E ? sgm[(e + 1) % E] : ns;
    if (p + 1 < S) {
      al[e] = interval_alpha(key[e], d1, sgm[e], s1);
      dm[e] = (key[e] + d1) * 0.5f;
      prod *= (1.0f - al[e] + 1e-10f);
    } else { al[e] = 0.0f; dm[e] = 0.0f; }
  }
  float T = warp_excl_prod(prod, lane);
  float wsum = 0.f, dnum = 0.f, w[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    w[e] = al[e] * T;
    T *= (1.0f - al[e] + 1e-10f);
    wsum += w[e];
    dnum = fmaf(w[e], dm[e], dnum);
  }
  wsum_out = warp_sum(wsum);
  dnum_out = warp_sum(dnum);
  float wprev = __shfl_up_sync(kFull, w[E - 1], 1);
  if (lane == 0) wprev = 0.0f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = lane * E + e;
    if (p < S) {
      const float omega = 0.5f * ((e == 0 ? wprev : w[(e + E - 1) % E]) + w[e]);
      if (kScatter) om[idx[e]] = omega;
      else { om[p] = omega; oi[p] = idx[e]; }
    }
  }
  // min / max of this ray's depths for the global clamp (VR/ray_marcher.py:50)
  float lmn = key[0], lmx = -__int_as_float(0x7f800000);
#pragma unroll
  for (int e = 0; e < E; ++e) if (lane * E + e < S) lmx = key[e];
  lmn = warp_min((lane * E) < S ? lmn : __int_as_float(0x7f800000));
  lmx = warp_max(lmx);
  mn = fminf(mn, lmn); mx = fmaxf(mx, lmx);
}

// Rank count of a ray's S depths by a PAIR of warps (h = 0, 1), for the kernels whose groups have half as many rays as
// there are ray warps (R = 4: 96+96 samples, the gen_videos.py:127-128 case, where the single-warp count above --
// S * ceil(S/32) compares per lane with half the ray warps idle -- was the critical path of the whole kernel).
// Two savings: (1) the coarse depths z[0..Dc) are already ascending (VR/renderer.py:169-192: linspace + jitter below one
// bin; checked here, not assumed), so a coarse sample's rank is its index plus the number of SMALLER FINE samples and only
// the fine samples are compared against everything; (2) each warp of the pair counts over half of every range and
// warp 1 hands its partial counts to warp 0 through `om`.  Warp 0 then scatters (depth, sigma, index) to the ranks:
// tmp[0..S) depths, tmp[S..2S) sigmas, om[0..S) original indices -- what warp_sort_and_weights(pre_ranked) reads.
// Returns (to warp 0; warp 1's return value is meaningless) whether the ranks form a permutation; ties, NaNs or
// a non-ascending coarse row make it false and the caller falls back to the single-warp path.
// Needs Dc % 32 == 0 and Df % 8 == 0.  `bar`: a named barrier id private to the pair.
template <int ER, int E0>
__device__ __forceinline__ void pair_count_coarse(const float* zc, int n, const float (&ze)[ER], int (&cnt)[ER]) {
#pragma unroll 2
  for (int j = 0; j < n; j += 4) {
    const float4 v = *reinterpret_cast<const float4*>(zc + j);
#pragma unroll
    for (int e = E0; e < ER; ++e)
      cnt[e] += (int)(v.x < ze[e]) + (int)(v.y < ze[e]) + (int)(v.z < ze[e]) + (int)(v.w < ze[e]);
  }
}
template <int ER>
__device__ __forceinline__ bool pair_rank_scatter(const float* z, const float* sg, float* om, float* tmp, int S, int Dc,
                                                  int lane, int h, int bar) {
  const int Df = S - Dc, e0 = Dc >> 5;
  float ze[ER]; int cnt[ER];
  bool asc = true;
#pragma unroll
  for (int e = 0; e < ER; ++e) {
    const int p = e * 32 + lane;
    ze[e] = p < S ? z[p] : __int_as_float(0x7f800000);
    cnt[e] = 0;
    if (p + 1 < Dc) asc &= z[p] <= z[p + 1];
  }
  {   // every row against this warp's half of the fine samples
    const float* zf = z + Dc + h * (Df >> 1);
    const int n = Df >> 1;
#pragma unroll 2
    for (int j = 0; j < n; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(zf + j);
#pragma unroll
      for (int e = 0; e < ER; ++e)
        cnt[e] += (int)(v.x < ze[e]) + (int)(v.y < ze[e]) + (int)(v.z < ze[e]) + (int)(v.w < ze[e]);
    }
  }
  {   // the fine rows (e >= e0) against this warp's half of the coarse samples
    const float* zc = z + h * (Dc >> 1);
    const int n = Dc >> 1;
    if (e0 == 3 && ER > 3) pair_count_coarse<ER, (ER > 3 ? 3 : 0)>(zc, n, ze, cnt);
    else if (e0 == 2 && ER > 2) pair_count_coarse<ER, (ER > 2 ? 2 : 0)>(zc, n, ze, cnt);
    else if (e0 == 4 && ER > 4) pair_count_coarse<ER, (ER > 4 ? 4 : 0)>(zc, n, ze, cnt);
    else {
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        const float v = zc[j];
#pragma unroll
        for (int e = 0; e < ER; ++e) cnt[e] += (int)(e >= e0 && v < ze[e]);
      }
    }
  }
  int* part = reinterpret_cast<int*>(om);
  if (h == 1) {
#pragma unroll
    for (int e = 0; e < ER; ++e) if (e * 32 + lane < S) part[e * 32 + lane] = cnt[e];
  }
  asm volatile("bar.sync %0, 64;" ::"r"(bar) : "memory");
  if (h == 1) return false;
  int rsum = 0;
#pragma unroll
  for (int e = 0; e < ER; ++e) {
    const int p = e * 32 + lane;
    if (p < S) { cnt[e] += part[p] + (p < Dc ? p : 0); rsum += cnt[e]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(kFull, rsum, o);
  const bool ok = __all_sync(kFull, asc) && rsum == S * (S - 1) / 2;
  if (ok) {
    float* zs = tmp; float* ss = tmp + S; int* is = reinterpret_cast<int*>(om);
    float sv[ER];
#pragma unroll
    for (int e = 0; e < ER; ++e) sv[e] = e * 32 + lane < S ? sg[e * 32 + lane] : 0.0f;
    __syncwarp();                                // every lane has read its partial counts out of `om`
#pragma unroll
    for (int e = 0; e < ER; ++e) {
      const int p = e * 32 + lane;
      if (p < S) { zs[cnt[e]] = ze[e]; ss[cnt[e]] = sv[e]; is[cnt[e]] = p; }
    }
    __syncwarp();
  }
  return ok;
}

// ---------------------------------------------------------------------------------------------------------------
// Sorting a ray's S = Dc + Df depths as a MERGE (R = 4 groups of the warp-specialised kernel, 96+96 samples: there the
// rank count above -- even shared by two warps -- was 13 of the ray warps' 41 k cycles per group and the critical path):
//   * the Dc coarse depths are ascending already (VR/renderer.py:169-192; checked, not assumed);
//   * an importance depth is a monotone function of its uniform draw (the inverse CDF, VR/renderer.py:240-252), so the
//     order of the Df importance depths among themselves is the order of their draws u -- known a whole pipeline step
//     before the depths exist.  warp_rank_draws ranks the draws (the four ray warps that have no ray to resample do
//     it while the other four resample); warp_merge_scatter then needs, per importance depth, one binary search in the
//     coarse row (c = number of smaller coarse depths; rank = rank of its draw + c) and, per coarse depth, a prefix sum
//     of the histogram of c (number of smaller importance depths).  ~200 instructions per warp instead of ~900.
// Float rounding can break the monotonicity by an ulp at a bin edge, so the scattered row is verified to be ascending;
// if it is not (or two draws are equal), the caller falls back to the rank count.
// ---------------------------------------------------------------------------------------------------------------
// rk[j] = number of draws smaller than u[j]; rk[0] = -1 if two draws are equal.  Df % 4 == 0, u 16-byte aligned.
template <int ROWS>
__device__ __forceinline__ void warp_rank_draws_rows(const float* u, int* rk, int Df, int lane) {
  float ue[ROWS]; int cnt[ROWS];
#pragma unroll
  for (int e = 0; e < ROWS; ++e) { const int j = e * 32 + lane; ue[e] = j < Df ? u[j] : __int_as_float(0x7f800000); cnt[e] = 0; }
#pragma unroll 2
  for (int j = 0; j < Df; j += 4) {
    const float4 v = *reinterpret_cast<const float4*>(u + j);
#pragma unroll
    for (int e = 0; e < ROWS; ++e)
      cnt[e] += (int)(v.x < ue[e]) + (int)(v.y < ue[e]) + (int)(v.z < ue[e]) + (int)(v.w < ue[e]);
  }
  int rsum = 0;
#pragma unroll
  for (int e = 0; e < ROWS; ++e) { const int j = e * 32 + lane; if (j < Df) { rk[j] = cnt[e]; rsum += cnt[e]; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(kFull, rsum, o);
  __syncwarp();
  if (lane == 0 && rsum != Df * (Df - 1) / 2) rk[0] = -1;
}
__device__ __forceinline__ void warp_rank_draws(const float* u, int* rk, int Df, int lane) {
  const int rows = (Df + 31) >> 5;
  if (rows <= 1) warp_rank_draws_rows<1>(u, rk, Df, lane);
  else if (rows == 2) warp_rank_draws_rows<2>(u, rk, Df, lane);
  else if (rows == 3) warp_rank_draws_rows<3>(u, rk, Df, lane);
  else if (rows == 4) warp_rank_draws_rows<4>(u, rk, Df, lane);
  else warp_rank_draws_rows<7>(u, rk, Df, lane);
}
// Scatter (depth, sigma, index) of the ray into sorted order: tmp[0..S) depths, tmp[S..2S) sigmas, om[0..S) indices (what
// warp_sort_and_weights(pre_ranked) reads).  hist: Dc + 1 ints of scratch.  Returns whether the result is a sorted row.
__device__ __forceinline__ bool warp_merge_scatter(const float* z, const float* sg, const int* rk, float* om, float* tmp,
                                                   int* hist, int S, int Dc, int lane) {
  const int Df = S - Dc;
  float* zs = tmp; float* ss = tmp + S; int* is = reinterpret_cast<int*>(om);
  bool ok = true;
  for (int k = lane; k <= Dc; k += 32) { hist[k] = 0; if (k + 1 < Dc) ok &= z[k] <= z[k + 1]; }
  __syncwarp();
  for (int j = lane; j < Df; j += 32) {
    const float f = z[Dc + j];
    int lo = 0, hi = Dc;                           // c = number of coarse depths < f
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (z[mid] < f) lo = mid + 1; else hi = mid; }
    atomicAdd(hist + lo, 1);
    const int r = rk[j] + lo;
    zs[r] = f; ss[r] = sg[Dc + j]; is[r] = Dc + j;
  }
  __syncwarp();
  {   // coarse depth k goes to k + (number of importance depths with c <= k)
    const int per = (Dc + 31) >> 5, k0 = lane * per;
    int loc = 0;
    for (int e = 0; e < per; ++e) if (k0 + e < Dc) loc += hist[k0 + e];
    int run = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, run, o); if (lane >= o) run += t; }
    run -= loc;
    for (int e = 0; e < per; ++e) {
      const int k = k0 + e;
      if (k < Dc) { run += hist[k]; const int r = k + run; zs[r] = z[k]; ss[r] = sg[k]; is[r] = k; }
    }
  }
  __syncwarp();
  for (int p = lane; p + 1 < S; p += 32) ok &= zs[p] <= zs[p + 1];
  return __all_sync(kFull, ok);
}

// one warp: coarse weights -> smoothed pdf -> CDF -> Df inverse-CDF draws (VR/renderer.py:194-253).
// z, sg: the ray's coarse depths / densities (S-strided row); w, pw, cdf: scratch rows; fine: output row;
// urow: the ray's Df uniform draws (global or shared memory).
__device__ __forceinline__ void warp_resample_ray(const RenderArgs& a, const float* z, const float* sg, float* w, float* pw,
                                                  float* cdf, float* fine, long long g, int lane, const float* urow) {
  const int Dc = a.Dc, Df = a.Df, nb = Dc - 3;
  warp_march_weights(z, sg, w, Dc, lane);
  __syncwarp();
  warp_smooth_weights(w, pw, nb, lane);
  __syncwarp();
  warp_cdf(pw, cdf, nb, lane);
  __syncwarp();
  for (int j = lane; j < Df; j += 32) {
    int inds;
    float smp = invert_cdf(cdf, nb, urow[j],
                           [&](int i) { return __fmul_rn(0.5f, __fadd_rn(z[i], z[i + 1])); }, inds);
    fine[j] = smp;
    if (a.fine_depths) a.fine_depths[g * Df + j] = smp;
    if (a.fine_inds) a.fine_inds[g * Df + j] = inds;
  }
}

}  // namespace tpr
