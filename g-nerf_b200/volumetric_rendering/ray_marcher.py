"""MipRayMarcher2 with the reference's interface (training/volumetric_rendering/ray_marcher.py:20-62).

Inside ImportanceRenderer.forward the march is fused into the render kernel; this stand-alone
module exists so code that calls the marcher directly keeps working."""
import ctypes

import torch

from .. import _lib


class MipRayMarcher2(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def run_forward(self, colors, densities, depths, rendering_options):
        """colors [N,M,S,C], densities [N,M,S,1], depths [N,M,S,1] ->
        (composite_rgb [N,M,C], composite_depth [N,M,1], weights [N,M,S-1,1])."""
        if rendering_options['clamp_mode'] != 'softplus':
            assert False, "MipRayMarcher only supports `clamp_mode`=`softplus`!"          # ray_marcher.py:35
        for t, name in ((colors, 'colors'), (densities, 'densities'), (depths, 'depths')):
            if not t.is_cuda:
                raise RuntimeError(f'{name} is on {t.device}: the B200 ray marcher has no CPU path')
            if t.dtype != torch.float32:
                raise RuntimeError(f'{name} must be float32')
            if torch.is_grad_enabled() and t.requires_grad:
                raise NotImplementedError('the B200 ray marcher is forward-only')
        colors, densities, depths = colors.contiguous(), densities.contiguous(), depths.contiguous()
        n, m, s, c = colors.shape
        dev = colors.device
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            rgb = torch.empty((n, m, c), device=dev, dtype=torch.float32)
            depth = torch.empty((n, m, 1), device=dev, dtype=torch.float32)
            weights = torch.empty((n, m, s - 1, 1), device=dev, dtype=torch.float32)
            rng = torch.empty(4, device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().tpr_ray_march(p(colors), p(densities), p(depths), n * m, s, c,
                                                int(bool(rendering_options.get('white_back', False))),
                                                p(rgb), p(depth), p(weights), p(rng), 1,
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       'tpr_ray_march')
        return rgb, depth, weights

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)
