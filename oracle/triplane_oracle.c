/*
 * triplane_oracle.c -- plain-C (OpenMP) restatement of the reference's tri-plane render path.
 *
 * TEST INFRASTRUCTURE ONLY.  Like oracle/triplane_oracle.py (same algorithm, same arithmetic
 * contract, see that file's header) this is the checker and the CPU baseline: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product (g-nerf_b200/) never links or calls it.  It exists next to the numpy oracle because
 * bench.py's reference arm must use all host threads; numpy's restatement is single threaded.
 *
 * Parity pinning: checked against the reference-generated fixtures in tests/golden/ by
 * tests/test_c_oracle.py (the reference itself ships no tests).
 *
 * Citations are relative to /root/reference/g_nerf/ ; VR/ = training/volumetric_rendering/.
 * Build: oracle/build_c.py (gcc -O3 -fopenmp -ffp-contract=off; contraction is off so the
 * coordinate arithmetic rounds exactly like the reference's separate float32 ops).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define C_FEAT 32
#define C_HID 64
#define C_OUT 33
#define MAX_S 512

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torch Softplus(beta=1, threshold=20) */
static inline float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

/* a4: grid_sample(bilinear, zeros, align_corners=False) on one channels-last plane
 * (VR/renderer.py:55-65); accumulates the plane's 32 channels into feat[]. */
static void plane_lookup(const float* plane /*[H][W][32]*/, int H, int W, float gu, float gv, float* feat) {
  float ix = ((gu + 1.0f) * (float)W - 1.0f) / 2.0f;
  float iy = ((gv + 1.0f) * (float)H - 1.0f) / 2.0f;
  float x0f = floorf(ix), y0f = floorf(iy);
  float wx1 = ix - x0f, wy1 = iy - y0f;
  float wx0 = (x0f + 1.0f) - ix, wy0 = (y0f + 1.0f) - iy;
  float lim;
  lim = x0f < -2.0f ? -2.0f : x0f; lim = lim > (float)(W + 1) ? (float)(W + 1) : lim; int x0 = (int)lim;
  lim = y0f < -2.0f ? -2.0f : y0f; lim = lim > (float)(H + 1) ? (float)(H + 1) : lim; int y0 = (int)lim;
  const int tx[4] = {x0, x0 + 1, x0, x0 + 1}, ty[4] = {y0, y0, y0 + 1, y0 + 1};
  const float tw[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};      /* nw ne sw se */
  float acc[C_FEAT];
  for (int c = 0; c < C_FEAT; ++c) acc[c] = 0.0f;
  for (int t = 0; t < 4; ++t) {
    if (tx[t] < 0 || tx[t] >= W || ty[t] < 0 || ty[t] >= H) continue;
    const float* v = plane + ((size_t)ty[t] * W + tx[t]) * C_FEAT;
    for (int c = 0; c < C_FEAT; ++c) acc[c] = acc[c] + v[c] * tw[t];
  }
  for (int c = 0; c < C_FEAT; ++c) feat[c] += acc[c];
}

/* a3-a6,a8: run_model for one point: project (plane0<-(x,y), plane1<-(x,z), plane2<-(z,x);
 * VR/renderer.py:29-53), gather, mean over planes (training/triplane.py:126), FC-softplus-FC
 * (:118-122, weights with gains already applied), colour activation (:134). */
static void eval_point(const float* img /*[3][H][W][32]*/, int H, int W, const float* w1, const float* b1,
                       const float* w2, const float* b2, float px, float py, float pz, float* rgb, float* sigma) {
  float f[C_FEAT];
  float p0[C_FEAT] = {0}, p1[C_FEAT] = {0}, p2[C_FEAT] = {0};
  const size_t ps = (size_t)H * W * C_FEAT;
  plane_lookup(img, H, W, px, py, p0);
  plane_lookup(img + ps, H, W, px, pz, p1);
  plane_lookup(img + 2 * ps, H, W, pz, px, p2);
  for (int c = 0; c < C_FEAT; ++c) f[c] = ((p0[c] + p1[c]) + p2[c]) / 3.0f;
  float h[C_HID];
  for (int j = 0; j < C_HID; ++j) {
    float a = 0.0f;
    for (int k = 0; k < C_FEAT; ++k) a += f[k] * w1[j * C_FEAT + k];
    h[j] = softplus_f(a + b1[j]);
  }
  for (int o = 0; o < C_OUT; ++o) {
    float a = 0.0f;
    for (int j = 0; j < C_HID; ++j) a += h[j] * w2[o * C_HID + j];
    a += b2[o];
    if (o == 0) *sigma = a;
    else rgb[o - 1] = (1.0f / (1.0f + expf(-a))) * 1.002f - 0.001f;
  }
}

/* a9: MipRayMarcher2 weights for S samples in the given order (VR/ray_marcher.py:26-42) */
static void march_weights(const float* d, const float* sg, int S, float* w) {
  float T = 1.0f;
  for (int i = 0; i < S - 1; ++i) {
    float delta = d[i + 1] - d[i];
    float dens = softplus_f((sg[i] + sg[i + 1]) / 2.0f - 1.0f);
    float alpha = 1.0f - expf(-(dens * delta));
    w[i] = alpha * T;
    T = T * (1.0f - alpha + 1e-10f);
  }
}

/* a10/a11: sample_importance + sample_pdf (VR/renderer.py:194-253) for one ray */
static void resample(const float* z, const float* w, int S, const float* u, int K, float* out, int32_t* inds_out) {
  const int nb = S - 3;
  float pw[MAX_S], cdf[MAX_S];
  double tot = 0.0;
  for (int j = 0; j < nb; ++j) {      /* max_pool(2,1,pad 1) -> avg_pool(2,1) -> +0.01, ends dropped */
    float a = w[j] > w[j + 1] ? w[j] : w[j + 1], b = w[j + 1] > w[j + 2] ? w[j + 1] : w[j + 2];
    pw[j] = ((a + b) / 2.0f + 0.01f) + 1e-5f;
    tot += (double)pw[j];
  }
  const float totf = (float)tot;
  double run = 0.0;
  cdf[0] = 0.0f;
  for (int j = 0; j < nb; ++j) { run += (double)(pw[j] / totf); cdf[j + 1] = (float)run; }
  for (int k = 0; k < K; ++k) {
    int lo = 0, hi = nb + 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u[k]) lo = mid + 1; else hi = mid; }
    int below = lo - 1 > 0 ? lo - 1 : 0, above = lo < nb ? lo : nb;
    float cb = cdf[below], ca = cdf[above];
    float bb = 0.5f * (z[below] + z[below + 1]), ba = 0.5f * (z[above] + z[above + 1]);
    float denom = ca - cb;
    if (denom < 1e-5f) denom = 1.0f;
    out[k] = bb + (u[k] - cb) / denom * (ba - bb);
    if (inds_out) inds_out[k] = lo;
  }
}

static float lin_torch(float start, float end, int steps, int i) {   /* torch.linspace, float32 */
  float step = (end - start) / (float)(steps - 1);
  return i < steps / 2 ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

/* a13: ImportanceRenderer.forward (VR/renderer.py:88-140).  planes are [N,3,32,H,W] like the
 * reference receives them; they are transposed to channels-last once (inside the timed call).
 * w1/b1/w2/b2 carry the FullyConnectedLayer gains already (networks_stylegan2.py:118-127). */
int oracle_render(const float* planes, int N, int H, int W, const float* w1, const float* b1, const float* w2,
                  const float* b2, const float* origins, const float* dirs, int M, const float* jitter, const float* u,
                  double ray_start, double ray_end, double box_warp, int Dc, int Df, int disparity, int white_back,
                  float* rgb_out, float* depth_out, float* wsum_out, float* fine_out, int32_t* inds_out) {
  const int S = Dc + Df;
  if (S > MAX_S || Dc < 2 || (Df > 0 && Dc < 4)) return -2;
  const size_t hw = (size_t)H * W;
  float* nhwc = (float*)malloc(sizeof(float) * (size_t)N * 3 * hw * C_FEAT);
  if (!nhwc) return -1;
#pragma omp parallel for schedule(static)
  for (long p = 0; p < (long)N * 3; ++p)
    for (size_t i = 0; i < hw; ++i)
      for (int c = 0; c < C_FEAT; ++c) nhwc[((size_t)p * hw + i) * C_FEAT + c] = planes[((size_t)p * C_FEAT + c) * hw + i];

  const float scale = (float)(2.0 / box_warp);
  const float rs = (float)ray_start, re = (float)ray_end;
  const float jscale = (float)((ray_end - ray_start) / (Dc - 1));
  const float inv_s = (float)(1.0 / ray_start), inv_e = (float)(1.0 / ray_end), dstep = (float)(1.0 / (Dc - 1));
  float gmin = INFINITY, gmax = -INFINITY;

#pragma omp parallel for schedule(dynamic, 16) reduction(min : gmin) reduction(max : gmax)
  for (long g = 0; g < (long)N * M; ++g) {
    const int n = (int)(g / M);
    const float* img = nhwc + (size_t)n * 3 * hw * C_FEAT;
    const float* o = origins + g * 3;
    const float* dir = dirs + g * 3;
    float d[MAX_S], sg[MAX_S], w[MAX_S];
    float (*col)[C_FEAT] = (float (*)[C_FEAT])malloc(sizeof(float) * S * C_FEAT);
    for (int k = 0; k < Dc; ++k) {                       /* a7: sample_stratified (:169-192) */
      float jit = jitter[g * Dc + k];
      if (disparity) {
        float t = lin_torch(0.0f, 1.0f, Dc, k) + jit * dstep;
        d[k] = 1.0f / (inv_s * (1.0f - t) + inv_e * t);
      } else {
        d[k] = lin_torch(rs, re, Dc, k) + jit * jscale;
      }
    }
    for (int k = 0; k < Dc; ++k)
      eval_point(img, H, W, w1, b1, w2, b2, (o[0] + d[k] * dir[0]) * scale, (o[1] + d[k] * dir[1]) * scale,
                 (o[2] + d[k] * dir[2]) * scale, col[k], &sg[k]);
    int order[MAX_S];
    for (int i = 0; i < S; ++i) order[i] = i;
    if (Df > 0) {
      march_weights(d, sg, Dc, w);
      resample(d, w, Dc, u + g * Df, Df, d + Dc, inds_out ? inds_out + g * Df : NULL);
      if (fine_out) memcpy(fine_out + g * Df, d + Dc, sizeof(float) * Df);
      for (int k = Dc; k < S; ++k)
        eval_point(img, H, W, w1, b1, w2, b2, (o[0] + d[k] * dir[0]) * scale, (o[1] + d[k] * dir[1]) * scale,
                   (o[2] + d[k] * dir[2]) * scale, col[k], &sg[k]);
      /* a12: unify_samples = stable sort by depth (:157-167); insertion sort keeps ties in order */
      for (int i = 1; i < S; ++i) {
        int oi = order[i], j = i - 1;
        while (j >= 0 && d[order[j]] > d[oi]) { order[j + 1] = order[j]; --j; }
        order[j + 1] = oi;
      }
    }
    float ds[MAX_S], ss[MAX_S];
    for (int i = 0; i < S; ++i) { ds[i] = d[order[i]]; ss[i] = sg[order[i]]; }
    march_weights(ds, ss, S, w);
    float acc[C_FEAT], wtot = 0.0f, dnum = 0.0f;
    for (int c = 0; c < C_FEAT; ++c) acc[c] = 0.0f;
    for (int i = 0; i < S - 1; ++i) {
      const float* c0 = col[order[i]];
      const float* c1 = col[order[i + 1]];
      for (int c = 0; c < C_FEAT; ++c) acc[c] += w[i] * ((c0[c] + c1[c]) / 2.0f);
      wtot += w[i];
      dnum += w[i] * ((ds[i] + ds[i + 1]) / 2.0f);
    }
    for (int c = 0; c < C_FEAT; ++c) {
      float v = acc[c];
      if (white_back) v = v + 1.0f - wtot;              /* VR/ray_marcher.py:52-53 */
      rgb_out[g * C_FEAT + c] = v * 2.0f - 1.0f;        /* :55 */
    }
    depth_out[g] = dnum / wtot;                         /* NaN -> inf and the global clamp below (:49-50) */
    wsum_out[g] = wtot;
    if (ds[0] < gmin) gmin = ds[0];
    if (ds[S - 1] > gmax) gmax = ds[S - 1];
    free(col);
  }
#pragma omp parallel for schedule(static)
  for (long g = 0; g < (long)N * M; ++g) {
    float v = depth_out[g];
    if (v != v) v = INFINITY;
    v = v < gmin ? gmin : v;
    depth_out[g] = v > gmax ? gmax : v;
  }
  free(nhwc);
  return 0;
}
