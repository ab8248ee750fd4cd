"""CPU, wherever the unmodified reference is present (/root/reference here, oracle/_ref on the GPU box): the launcher's pieces --
install(), the load_network_pkl wrapper and the backbone / plane cache -- against the REAL, unmodified reference
TriPlaneGenerator.  CPU tensors keep going to the reference renderer (install.py dispatches on device), so this checks the
host logic either side of the hot path, not the kernels."""
import os
import sys

import pytest
import torch

from oracle import ref_loader

REF = ref_loader.reference_dir()
pytestmark = pytest.mark.skipif(REF is None, reason='reference not present (neither /root/reference nor oracle/_ref)')


@pytest.fixture(scope='module')
def generator():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dnnlib
    from training.triplane import TriPlaneGenerator
    rk = {'image_resolution': 512, 'disparity_space_sampling': False, 'clamp_mode': 'softplus',
          'superresolution_module': 'training.superresolution.SuperresolutionHybrid8XDC',
          'c_gen_conditioning_zero': False, 'gpc_reg_prob': True, 'c_scale': 1, 'superresolution_noise_mode': 'none',
          'density_reg': 0.25, 'density_reg_p_dist': 0.004, 'reg_type': 'l1', 'decoder_lr_mul': 1, 'sr_antialias': True,
          'depth_resolution': 6, 'depth_resolution_importance': 6, 'ray_start': 2.25, 'ray_end': 3.3, 'box_warp': 1,
          'avg_camera_radius': 2.7, 'avg_camera_pivot': [0, 0, 0.2]}                     # train.py:310-335, small depths
    torch.manual_seed(0)
    G = TriPlaneGenerator(z_dim=512, c_dim=25, w_dim=512, img_resolution=512, img_channels=3,
                          mapping_kwargs=dnnlib.EasyDict(num_layers=2), channel_base=32768, channel_max=512,
                          fused_modconv_default='inference_only', rendering_kwargs=rk, num_fp16_res=0, conv_clamp=None,
                          sr_num_fp16_res=4, sr_kwargs=dnnlib.EasyDict(channel_base=32768, channel_max=512,
                          fused_modconv_default='inference_only', w_dim=512)).eval().requires_grad_(False)
    return G


def _camera():
    from camera_utils import LookAtPoseSampler
    pose = LookAtPoseSampler.sample(3.14 / 2, 3.14 / 2, radius=2.7)
    K = torch.tensor([[4.2647, 0, 0.5], [0, 4.2647, 0.5], [0, 0, 1]])
    return torch.cat([pose.reshape(-1, 16), K.reshape(-1, 9)], 1)


def test_plane_cache_skips_the_backbone_and_keeps_results(pkg, generator):
    G = generator
    pkg.install()
    try:
        c = _camera()
        ws = G.mapping(z=torch.randn(1, 512), c=torch.zeros_like(c))
        calls = []
        first_block = next(m for n, m in G.backbone.synthesis.named_children())     # runs once per real backbone pass
        hook = first_block.register_forward_hook(lambda *a: calls.append(1))
        kw = dict(ws=ws, c=c, noise_mode='const', neural_rendering_resolution=64, only_depth=True)
        torch.manual_seed(1); base = G.synthesis(**kw)['image_depth']
        assert len(calls) == 1
        memo = pkg.enable_plane_cache(G)
        assert G.renderer.cache_packed_planes is True
        torch.manual_seed(1); a = G.synthesis(**kw)['image_depth']
        torch.manual_seed(1); b = G.synthesis(**kw)['image_depth']
        assert (memo.misses, memo.hits) == (1, 1) and len(calls) == 2          # the second call never reached the backbone
        torch.testing.assert_close(a, base, rtol=0, atol=0)
        torch.testing.assert_close(b, base, rtol=0, atol=0)
        # a different latent, an in-place edit of the same latent, or a random noise mode must all miss
        G.synthesis(**dict(kw, ws=ws.clone()))
        assert memo.misses == 2
        ws2 = ws.clone(); G.synthesis(**dict(kw, ws=ws2)); ws2.add_(0.1); G.synthesis(**dict(kw, ws=ws2))
        assert memo.misses == 4
        before = (memo.misses, memo.hits)
        G.synthesis(**dict(kw, noise_mode='random'))
        assert (memo.misses, memo.hits) == before and len(calls) == 6
        hook.remove()
        pkg.disable_plane_cache(G)
        assert 'forward' not in G.backbone.synthesis.__dict__ and G.renderer.cache_packed_planes is False
        assert not any('tpr' in k for k in G.state_dict())
    finally:
        pkg.uninstall()


def test_loader_wrapper_configures_generators(pkg, generator):
    import types
    from importlib import import_module
    launch = import_module('g-nerf_b200.launch')
    fake_legacy = types.SimpleNamespace(load_network_pkl=lambda f: {'G_ema': generator, 'training_set_kwargs': {'x': 1}})
    launch.patch_loader(fake_legacy, {'decoder_precision': 'bf16', 'output_layout': 'channels_first'})
    launch.patch_loader(fake_legacy, {})                                       # idempotent
    try:
        data = fake_legacy.load_network_pkl(None)
        G = data['G_ema']
        assert G.rendering_kwargs['decoder_precision'] == 'bf16' and G.rendering_kwargs['output_layout'] == 'channels_first'
        assert '_tpr_memo' in G.backbone.synthesis.__dict__
    finally:
        pkg.disable_plane_cache(generator)
        generator.rendering_kwargs.pop('decoder_precision', None); generator.rendering_kwargs.pop('output_layout', None)


def test_install_patches_the_real_reference_classes(pkg):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from training.volumetric_rendering import renderer as ref_r, ray_sampler as ref_s
    orig = ref_r.ImportanceRenderer.forward
    pkg.install()
    try:
        assert ref_r.ImportanceRenderer.forward is not orig and ref_r.ImportanceRenderer.forward.__wrapped__ is orig
        assert ref_s.RaySampler.forward.__wrapped__ is not None
        # CPU tensors still reach the reference implementation
        o, d = ref_s.RaySampler()(torch.eye(4)[None], torch.tensor([[[4.2647, 0, .5], [0, 4.2647, .5], [0, 0, 1]]]), 4)
        assert o.shape == (1, 16, 3) and not o.is_cuda
    finally:
        pkg.uninstall()
    assert ref_r.ImportanceRenderer.forward is orig
