// Launcher of the warp-specialised fused forward (kernel: tpr_render_ws.cuh; instantiations: tpr_render_ws_*_{a,b}.cu).
#include "tpr_render_ws.cuh"

namespace tpr {
namespace ws {

template <int MODE>
static Kernel pick_kernel(int S, bool prof, bool train) {
  return S <= 96 ? kernel_small<MODE>(S, prof, train) : kernel_large<MODE>(S, prof);
}

}  // namespace ws

// Rays per group (8 or 4) the warp-specialised kernel would use, or 0 if (Dc, Df) does not fit its slot ring:
// the pipelined job order keeps the coarse slots of two groups and the fine slots of one alive at once.
// mode: 1 = bf16, 2 = 2xFP16 (the template parameter MODE)
int ws_rays_per_group(int Dc, int Df, int mode) {
  const int ns = mode == 1 ? ws::Cols<1>::ns : ws::Cols<2>::ns;
  for (int R = 8; R >= 4; R >>= 1) {
    const int dpt = ws::kRows / R;
    const int nc = (Dc + dpt - 1) / dpt, nf = Df > 0 ? (Df + dpt - 1) / dpt : 0;
    if (2 * nc + nf <= ns) return R;
  }
  return 0;
}

// Sample counts for which a TRAIN instantiation exists (the reference's training depths, 48+48: train.py:312-313).
bool ws_keeps_samples(int Dc, int Df) { const int S = Dc + Df; return S > 64 && S <= 96 && ws_rays_per_group(Dc, Df, 2) == 8; }

// Launch; returns cudaError_t (0 = ok), or -1 if the configuration does not fit (the caller falls back).
int launch_render_ws(RenderArgs a, int mode, int sms, int smem_optin, long long n_img, long long n_rays, cudaStream_t st) {
  const int S = a.Dc + a.Df;
  a.R = ws_rays_per_group(a.Dc, a.Df, mode);
  if (a.R == 0) return -1;
  if (a.col_w > 0 && (a.col_w % a.R != 0 || (long long)a.col_w * a.col_w != n_rays)) a.col_w = 0;
  a.tiles_per_img = (n_rays + a.R - 1) / a.R;
  a.n_tiles = a.tiles_per_img * n_img;
  if (a.n_tiles >= (1ll << 31) || n_rays >= (1ll << 31)) return -1;      // group_geom works in 32 bits
  const bool train = a.sample_colours != nullptr && a.sample_sigma != nullptr;
  if (train && !ws_keeps_samples(a.Dc, a.Df)) return -1;
  ws::Kernel k = mode == 1 ? ws::pick_kernel<1>(S, a.dbg != nullptr, train) : ws::pick_kernel<2>(S, a.dbg != nullptr, train);
  const size_t smem = mode == 1 ? ws::smem_bytes<1>(a.R, S, a.Df) : ws::smem_bytes<2>(a.R, S, a.Df);
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, k);
  if (e != cudaSuccess) return (int)e;
  if ((int)smem > smem_optin - (int)fa.sharedSizeBytes) return -1;
  e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const long long grid = a.n_tiles < sms ? a.n_tiles : sms;      // one CTA per SM: each owns all 512 TMEM columns
  k<<<(unsigned)grid, ws::kThreads, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace tpr
