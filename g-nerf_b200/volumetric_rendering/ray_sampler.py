"""RaySampler with the reference's interface (training/volumetric_rendering/ray_sampler.py:18-63),
one CUDA kernel instead of ~15 ATen launches."""
import ctypes

import torch

from .. import _lib


class RaySampler(torch.nn.Module):
    def __init__(self):
        super().__init__()
        # attributes the reference defines (ray_sampler.py:21); never read by it either
        self.ray_origins_h, self.ray_directions, self.depths, self.image_coords, self.rendering_options = \
            None, None, None, None, None

    def forward(self, cam2world_matrix, intrinsics, resolution):
        """cam2world_matrix [N,4,4], intrinsics [N,3,3], resolution int ->
        (ray_origins [N,res*res,3], ray_dirs [N,res*res,3]); x is the fastest pixel index (:44)."""
        for t, name, tail in ((cam2world_matrix, 'cam2world_matrix', (4, 4)), (intrinsics, 'intrinsics', (3, 3))):
            if not t.is_cuda:
                raise RuntimeError(f'{name} is on {t.device}: the B200 ray sampler has no CPU path')
            if t.dtype != torch.float32 or tuple(t.shape[1:]) != tail:
                raise RuntimeError(f'{name} must be float32 [N,{tail[0]},{tail[1]}], got {t.dtype} {tuple(t.shape)}')
        c2w, K = cam2world_matrix.contiguous(), intrinsics.contiguous()
        n, res = c2w.shape[0], int(resolution)
        if K.shape[0] != n:
            raise RuntimeError('cam2world_matrix and intrinsics disagree on the batch size')
        dev = c2w.device
        with torch.cuda.device(dev):
            origins = torch.empty((n, res * res, 3), device=dev, dtype=torch.float32)
            dirs = torch.empty((n, res * res, 3), device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().tpr_ray_sample(ctypes.c_void_p(c2w.data_ptr()), ctypes.c_void_p(K.data_ptr()), n, res,
                                                 ctypes.c_void_p(origins.data_ptr()), ctypes.c_void_p(dirs.data_ptr()),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       'tpr_ray_sample')
        return origins, dirs
