"""Import the UNMODIFIED reference (TEST / BASELINE INFRASTRUCTURE ONLY; see build_ref.py) and build its objects.

Nothing here restates the reference: the classes returned are the reference's own, imported from ``oracle/_ref/g_nerf``
(the travelling copy) or ``/root/reference/g_nerf`` (the build container).  Used by tests/, bench.py's baseline legs
(``--impl reference``, ``gpu_baseline``, ``config3``) and smoke(); never by the product.
"""
import importlib
import sys
import types

from . import build_ref


def reference_dir():
    return build_ref.reference_dir()


def import_reference():
    """Put the reference's g_nerf directory on sys.path (once) and return its hot-path modules."""
    root = reference_dir()
    if root is None:
        raise RuntimeError('the reference is not available: neither oracle/_ref/g_nerf (python oracle/build_ref.py in the '
                           'build container) nor /root/reference/g_nerf exists')
    if root not in sys.path:
        sys.path.insert(0, root)
    mods = types.SimpleNamespace(root=root)
    mods.renderer = importlib.import_module('training.volumetric_rendering.renderer')
    mods.ray_sampler = importlib.import_module('training.volumetric_rendering.ray_sampler')
    mods.ray_marcher = importlib.import_module('training.volumetric_rendering.ray_marcher')
    mods.math_utils = importlib.import_module('training.volumetric_rendering.math_utils')
    mods.triplane = importlib.import_module('training.triplane')
    mods.camera_utils = importlib.import_module('camera_utils')
    mods.dnnlib = importlib.import_module('dnnlib')
    return mods


# rendering_kwargs of the FFHQ configuration (train.py:310-335); SURVEY.md appendix
FFHQ_RENDERING_KWARGS = {
    'image_resolution': 512, 'disparity_space_sampling': False, 'clamp_mode': 'softplus',
    'superresolution_module': 'training.superresolution.SuperresolutionHybrid8XDC',
    'c_gen_conditioning_zero': False, 'gpc_reg_prob': True, 'c_scale': 1, 'superresolution_noise_mode': 'none',
    'density_reg': 0.25, 'density_reg_p_dist': 0.004, 'reg_type': 'l1', 'decoder_lr_mul': 1, 'sr_antialias': True,
    'depth_resolution': 48, 'depth_resolution_importance': 48, 'ray_start': 2.25, 'ray_end': 3.3, 'box_warp': 1,
    'avg_camera_radius': 2.7, 'avg_camera_pivot': [0, 0, 0.2]}


def make_generator(seed=0, **rendering_overrides):
    """Random-init TriPlaneGenerator of the FFHQ shape (train.py:239,275-277,302-304,364,375-377,400-401), eval mode,
    requires_grad off (snapshots are saved that way, training_loop.py:522,538).  On the CPU; move it with .to(device)."""
    import torch
    ref = import_reference()
    rk = dict(FFHQ_RENDERING_KWARGS, **rendering_overrides)
    torch.manual_seed(seed)
    E = ref.dnnlib.EasyDict
    G = ref.triplane.TriPlaneGenerator(
        z_dim=512, c_dim=25, w_dim=512, img_resolution=512, img_channels=3, mapping_kwargs=E(num_layers=2),
        channel_base=32768, channel_max=512, fused_modconv_default='inference_only', rendering_kwargs=rk, num_fp16_res=0,
        conv_clamp=None, sr_num_fp16_res=4,
        sr_kwargs=E(channel_base=32768, channel_max=512, fused_modconv_default='inference_only', w_dim=512))
    return G.eval().requires_grad_(False)


def make_decoder(seed=0):
    """Random-init reference OSGDecoder (training/triplane.py:111-136)."""
    import torch
    ref = import_reference()
    torch.manual_seed(seed)
    return ref.triplane.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).requires_grad_(False)
