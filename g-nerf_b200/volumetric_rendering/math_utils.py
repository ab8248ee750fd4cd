"""The part of training/volumetric_rendering/math_utils.py the hot path touches:
get_ray_limits_box (:46-98), used only when ray_start == ray_end == 'auto' (VR/renderer.py:91-97)."""
import ctypes

import torch

from .. import _lib


def get_ray_limits_box(rays_o: torch.Tensor, rays_d: torch.Tensor, box_side_length):
    """Slab-method ray / axis-aligned-box intersection.  rays_o, rays_d [...,3] ->
    (t_min [...,1], t_max [...,1]); rays that miss get (-1, -2)."""
    if not rays_o.is_cuda:
        raise RuntimeError('get_ray_limits_box: the B200 path has no CPU implementation')
    o = rays_o.detach().reshape(-1, 3).contiguous().float()
    d = rays_d.detach().reshape(-1, 3).contiguous().float()
    n = o.shape[0]
    with torch.cuda.device(o.device):
        tmin = torch.empty(n, device=o.device, dtype=torch.float32)
        tmax = torch.empty(n, device=o.device, dtype=torch.float32)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(_lib.lib().tpr_ray_limits_box(p(o), p(d), n, float(box_side_length), p(tmin), p(tmax),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   'tpr_ray_limits_box')
    shape = tuple(rays_o.shape[:-1]) + (1,)
    return tmin.reshape(shape), tmax.reshape(shape)
