// Instantiations of render_ws_kernel (tpr_render_ws.cuh) for one decoder mode and one sample-count class.
#include "tpr_render_ws.cuh"

namespace tpr {
namespace ws {
template <> Kernel kernel_large<1>(int S, bool prof) {
  if (prof && S > 128 && S <= 192) return render_ws_kernel<1, 8, 6, true>;       // (192 samples: TPR_PT_DEPTH=96)
#ifdef TPR_DEV_BUILD
  return render_ws_kernel<1, 8, 6, false>;
#else
  return S <= 128 ? render_ws_kernel<1, 4, 4, false> : S <= 192 ? render_ws_kernel<1, 8, 6, false> : render_ws_kernel<1, 8, 8, false>;
#endif
}
}  // namespace ws
}  // namespace tpr
