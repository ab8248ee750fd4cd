// Stand-alone entry points for the helper functions of the reference's renderer module that the fused forward never
// calls but a user of the module may: sample_from_planes (VR/renderer.py:55-65), sort_samples / unify_samples
// (:150-167), sample_from_3dgrid (:67-80) and the density_noise term of run_model (:146).
// Citations relative to /root/reference/g_nerf/ (VR/ = training/volumetric_rendering/).
#include <cuda_runtime.h>
#include <stdint.h>

#include "triplane_b200.h"
#include "tpr_device.cuh"
#include "tpr_host.h"

namespace tpr {

// ---------------------------------------------------------------------------------------------------------
// a4: one 8-lane group per (point, plane): four whole 128-byte texels in flight per lane, blended and written as
// one 128-byte row of features [N,3,P,32].  HBM-bound on the 384 bytes it writes per point.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_planes_kernel(const float* __restrict__ planes, int H, int W,
                                                            const float* __restrict__ xyz, long long n_img, long long n_pts,
                                                            float box_scale, float* __restrict__ feat) {
  const int sub = threadIdx.x & 7;
  const long long total = n_img * 3 * n_pts;
  const size_t plane_floats = (size_t)H * W * kC;
  for (long long it = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3; it < total;
       it += ((long long)gridDim.x * blockDim.x) >> 3) {
    const long long n = it / (3 * n_pts), rem = it - n * 3 * n_pts;
    const int p = (int)(rem / n_pts);
    const long long q = rem - (long long)p * n_pts;
    const float* c = xyz + (n * n_pts + q) * 3;
    const float px = __fmul_rn(c[0], box_scale), py = __fmul_rn(c[1], box_scale), pz = __fmul_rn(c[2], box_scale);
    Taps t;
    plane_taps(p == 2 ? pz : px, p == 0 ? py : (p == 1 ? pz : px), H, W, t);      // (x,y) (x,z) (z,x)
    const float* img = planes + ((size_t)n * 3 + p) * plane_floats + sub * 4;
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = ldg128(img + t.off[i]);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) fma4(a, t.w[i], v[i]);
    *reinterpret_cast<float4*>(feat + it * kC + sub * 4) = a;
  }
}

__global__ void add_density_noise_kernel(float* __restrict__ sigma, const float* __restrict__ noise, long long n, float dn) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    sigma[i] = __fadd_rn(sigma[i], __fmul_rn(noise[i], dn));
}

// ---------------------------------------------------------------------------------------------------------
// a12 / a15: one warp per ray.  The S depths are sorted with the bitonic network of the fused kernels (ties broken by
// input index, i.e. a stable sort); colours follow through the sorted index, 128 bytes per lane group at a time.
// ---------------------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(256) sort_samples_kernel(const float* __restrict__ z, const float* __restrict__ col,
                                                           const float* __restrict__ sg, long long n_rays, int S, int C,
                                                           float* __restrict__ zs, float* __restrict__ cols,
                                                           float* __restrict__ sgs) {
  __shared__ int order[8][TPR_MAX_SAMPLES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long g = blockIdx.x * 8LL + warp; g < n_rays; g += gridDim.x * 8LL) {
    float key[E]; int idx[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = lane * E + e;
      key[e] = p < S ? z[g * S + p] : __int_as_float(0x7f800000);
      idx[e] = p;
    }
    warp_bitonic_sort<E>(key, idx, lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = lane * E + e;
      if (p < S) { zs[g * S + p] = key[e]; sgs[g * S + p] = sg[g * S + idx[e]]; order[warp][p] = idx[e]; }
    }
    __syncwarp();
    const float* src = col + g * (long long)S * C;
    float* dst = cols + g * (long long)S * C;
    for (int i = lane; i < S * C; i += 32) {
      const int p = i / C, c = i - p * C;
      dst[i] = src[(long long)order[warp][p] * C + c];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// a15: trilinear lookup, one thread per (point, channel); grid [G,C,D,H,W]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_3dgrid_kernel(const float* __restrict__ grid, long long n_grids, int C, int D,
                                                            int H, int W, const float* __restrict__ coords, long long n_batch,
                                                            long long n_pts, float* __restrict__ out) {
  const long long total = n_batch * n_pts * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pt = i / C;
    const int c = (int)(i - pt * C);
    const long long n = pt / n_pts;
    const float* q = coords + pt * 3;
    // unnormalise (align_corners=False): ((g + 1) * size - 1) / 2
    const float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(q[0], 1.0f), (float)W), -1.0f), 0.5f);
    const float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(q[1], 1.0f), (float)H), -1.0f), 0.5f);
    const float iz = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(q[2], 1.0f), (float)D), -1.0f), 0.5f);
    const float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    const float fx = ix - x0f, fy = iy - y0f, fz = iz - z0f;
    const int x0 = (int)fminf(fmaxf(x0f, -2.0f), (float)(W + 1)), y0 = (int)fminf(fmaxf(y0f, -2.0f), (float)(H + 1)),
              z0 = (int)fminf(fmaxf(z0f, -2.0f), (float)(D + 1));
    const float* g = grid + ((n_grids == 1 ? 0 : n) * C + c) * (long long)D * H * W;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
      const int x = x0 + dx, y = y0 + dy, zc = z0 + dz;
      if (x < 0 || x >= W || y < 0 || y >= H || zc < 0 || zc >= D) continue;       // zero padding
      const float w = (dx ? fx : 1.0f - fx) * (dy ? fy : 1.0f - fy) * (dz ? fz : 1.0f - fz);
      acc = fmaf(w, __ldg(g + ((long long)zc * H + y) * W + x), acc);
    }
    out[i] = acc;
  }
}

}  // namespace tpr

using namespace tpr;

#define TPR_LAUNCHED(what)                                                  \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) return cuda_fail(e__, what);                    \
  } while (0)

extern "C" {

int tpr_sample_planes(const float* planes_packed, int64_t n_img, int32_t height, int32_t width, const float* xyz, int64_t n_pts,
                      double box_warp, float* features, void* stream) {
  if (!planes_packed || !xyz || !features) return fail(TPR_E_NULL, "tpr_sample_planes: NULL pointer");
  if (n_img <= 0 || n_pts <= 0 || height <= 0 || width <= 0 || (int64_t)height * width > (1 << 24))
    return fail(TPR_E_SHAPE, "tpr_sample_planes: bad shape");
  if (!(box_warp > 0.0)) return fail(TPR_E_OPTION, "tpr_sample_planes: box_warp must be > 0");
  DeviceInfo di = device_info();
  if (!di.ok) return fail(TPR_E_DEVICE, "tpr_sample_planes: no CUDA device");
  const long long groups = (long long)n_img * 3 * n_pts;
  sample_planes_kernel<<<grid_for(groups * 8, 256, di.sms, 16), 256, 0, (cudaStream_t)stream>>>(
      planes_packed, height, width, xyz, n_img, n_pts, (float)(2.0 / box_warp), features);
  TPR_LAUNCHED("sample_planes_kernel");
  return 0;
}

int tpr_add_density_noise(float* sigma, const float* noise, int64_t n, double density_noise, void* stream) {
  if (!sigma || !noise) return fail(TPR_E_NULL, "tpr_add_density_noise: NULL pointer");
  if (n <= 0) return fail(TPR_E_SHAPE, "tpr_add_density_noise: n <= 0");
  DeviceInfo di = device_info();
  if (!di.ok) return fail(TPR_E_DEVICE, "tpr_add_density_noise: no CUDA device");
  add_density_noise_kernel<<<grid_for(n, 256, di.sms, 8), 256, 0, (cudaStream_t)stream>>>(sigma, noise, n, (float)density_noise);
  TPR_LAUNCHED("add_density_noise_kernel");
  return 0;
}

int tpr_sort_samples(const float* depths, const float* colours, const float* densities, int64_t n_rays, int32_t n_samples,
                     int32_t n_channels, float* depths_sorted, float* colours_sorted, float* densities_sorted, void* stream) {
  if (!depths || !colours || !densities || !depths_sorted || !colours_sorted || !densities_sorted)
    return fail(TPR_E_NULL, "tpr_sort_samples: NULL pointer");
  if (n_rays <= 0 || n_samples <= 0 || n_samples > TPR_MAX_SAMPLES || n_channels <= 0)
    return fail(TPR_E_SHAPE, "tpr_sort_samples: need n_rays > 0, 0 < n_samples <= 256, n_channels > 0");
  DeviceInfo di = device_info();
  if (!di.ok) return fail(TPR_E_DEVICE, "tpr_sort_samples: no CUDA device");
  const int grid = grid_for(n_rays, 8, di.sms, 8);
  cudaStream_t st = (cudaStream_t)stream;
#define TPR_SORT(E) sort_samples_kernel<E><<<grid, 256, 0, st>>>(depths, colours, densities, n_rays, n_samples, n_channels, \
                                                                 depths_sorted, colours_sorted, densities_sorted)
  if (n_samples <= 32) TPR_SORT(1);
  else if (n_samples <= 64) TPR_SORT(2);
  else if (n_samples <= 128) TPR_SORT(4);
  else TPR_SORT(8);
#undef TPR_SORT
  TPR_LAUNCHED("sort_samples_kernel");
  return 0;
}

int tpr_sample_3dgrid(const float* grid, int64_t n_grids, int32_t channels, int32_t depth, int32_t height, int32_t width,
                      const float* coords, int64_t n_batch, int64_t n_pts, float* features, void* stream) {
  if (!grid || !coords || !features) return fail(TPR_E_NULL, "tpr_sample_3dgrid: NULL pointer");
  if (n_batch <= 0 || n_pts <= 0 || channels <= 0 || depth <= 0 || height <= 0 || width <= 0 || (n_grids != 1 && n_grids != n_batch))
    return fail(TPR_E_SHAPE, "tpr_sample_3dgrid: bad shape (the grid's batch must be 1 or the coordinates' batch)");
  DeviceInfo di = device_info();
  if (!di.ok) return fail(TPR_E_DEVICE, "tpr_sample_3dgrid: no CUDA device");
  const long long total = (long long)n_batch * n_pts * channels;
  sample_3dgrid_kernel<<<grid_for(total, 256, di.sms, 16), 256, 0, (cudaStream_t)stream>>>(
      grid, n_grids, channels, depth, height, width, coords, n_batch, n_pts, features);
  TPR_LAUNCHED("sample_3dgrid_kernel");
  return 0;
}

}  // extern "C"
