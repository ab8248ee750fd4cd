// Instantiations of render_ws_kernel (tpr_render_ws.cuh) for one decoder mode and one sample-count class.
#include "tpr_render_ws.cuh"

namespace tpr {
namespace ws {
template <> Kernel kernel_small<2>(int S, bool prof, bool train) {
  if (train && !prof && S > 64) return render_ws_kernel<2, 4, 3, false, true>;
  if (prof && S > 64) return render_ws_kernel<2, 4, 3, true>;                    // TPR_PHASE_TIMING=1 (profiles/phase_timing.py)
#ifdef TPR_DEV_BUILD          // A/B builds (build.py --alt): only the 48+48 and 96+96 instantiations
  return render_ws_kernel<2, 4, 3, false>;
#else
  return S <= 64 ? render_ws_kernel<2, 2, 2, false> : render_ws_kernel<2, 4, 3, false>;
#endif
}
}  // namespace ws
}  // namespace tpr
