"""Autograd for ImportanceRenderer.forward: gradients w.r.t. the tri-planes and the OSGDecoder's four tensors.

The reference gets its backward from PyTorch autograd over ~40 ATen ops and ~10 GB of saved activations per forward at
FFHQ training shape; training back-propagates through the renderer at
training/training_loop.py:335,377.  Here the forward is the fused kernel (nothing per-sample is saved) and the backward is
``tpr_render_backward`` (csrc/tpr_backward.cu): it re-evaluates colours / densities at the forward's sample depths and
runs the march, the decoder and the bilinear gather backwards.  Saved between the two: the packed planes, the packed
decoder, the rays, the S sample depths per ray, the depth range and -- written by the forward kernel's composite on its
way -- the 32 colours and sigma of every sample (132 B per sample; eager autograd keeps ~800 B per sample).

The graph is the reference's: importance depths are constants (VR/renderer.py:198,210: no_grad + detach), so is the
jitter; ray origins / directions are not differentiated (they raise if they require grad).
"""
import ctypes

import torch

from .. import _lib
from . import renderer as _r


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, renderer, decoder, options, noise, planes, w1, b1, w2, b2, origins, dirs):
        with torch.no_grad():
            rgb, depth, wsum, aux = renderer._forward_impl(planes.detach(), decoder, origins, dirs, options, noise=noise,
                                                           train=True)
        fc1, fc2 = decoder.net[0], decoder.net[2]
        ctx.gains = (float(fc1.weight_gain), float(fc1.bias_gain), float(fc2.weight_gain), float(fc2.bias_gain))
        ctx.packed, ctx.opts = aux['packed'], aux['options']
        ctx.shape = tuple(origins.shape[:2])
        fine = aux['fine'] if aux['fine'] is not None else torch.empty(0, device=rgb.device)
        ctx.has_fine = aux['fine'] is not None
        ctx.has_saved = aux['saved'] is not None
        s_col, s_sig, s_feat = aux['saved'] if ctx.has_saved else (None, None, None)
        ctx.has_feat = s_feat is not None
        empty = torch.empty(0, device=rgb.device)
        ctx.save_for_backward(aux['packed'].data, aux['dec'], origins.contiguous(), dirs.contiguous(), aux['coarse'], fine,
                              aux['range'], s_col if ctx.has_saved else empty, s_sig if ctx.has_saved else empty,
                              s_feat if ctx.has_feat else empty)
        return rgb, depth, wsum

    @staticmethod
    @torch.autograd.function.once_differentiable          # raw CUDA kernels: a double backward must raise, not detach silently
    def backward(ctx, g_rgb, g_depth, g_wsum):
        packed, dec, origins, dirs, coarse, fine, rng, s_col, s_sig, s_feat = ctx.saved_tensors
        n, m = ctx.shape
        pp, o = ctx.packed, ctx.opts
        dev = packed.device
        L = _lib.lib()
        p, st = _r._ptr, _r._stream
        zeros = lambda shape: torch.zeros(shape, device=dev, dtype=torch.float32)
        g_rgb = g_rgb.contiguous().float() if g_rgb is not None else zeros((n, m, 32))
        g_depth = g_depth.contiguous().float() if g_depth is not None else zeros((n, m, 1))
        g_wsum = g_wsum.contiguous().float() if g_wsum is not None else zeros((n, m, 1))
        need = ctx.needs_input_grad
        want_planes, want_dec = need[4], any(need[5:9])
        with torch.cuda.device(dev):
            g_planes = torch.empty_like(packed) if want_planes else None
            g_dec = torch.empty_like(dec) if want_dec else None
            nbytes = L.tpr_render_backward_scratch_bytes(n, m, o.depth_resolution + o.depth_resolution_importance)
            scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            o2 = _lib.TprOptions.from_buffer_copy(o)
            o2.plane_sets, o2.depth_clamp_group = 0, 0
            _lib.check(L.tpr_render_backward(p(packed), n, pp.height, pp.width, p(dec), p(origins), p(dirs), m, p(coarse),
                                             p(fine) if ctx.has_fine else None, p(rng), ctypes.byref(o2), p(g_rgb), p(g_depth),
                                             p(g_wsum), p(s_col) if ctx.has_saved else None, p(s_sig) if ctx.has_saved else None,
                                             p(s_feat) if ctx.has_feat else None,
                                             p(g_planes), p(g_dec), p(scratch), nbytes, st()),
                       'tpr_render_backward')
            g_w1 = g_b1 = g_w2 = g_b2 = None
            if want_dec:
                g_w1 = torch.empty((64, 32), device=dev, dtype=torch.float32)
                g_b1 = torch.empty(64, device=dev, dtype=torch.float32)
                g_w2 = torch.empty((33, 64), device=dev, dtype=torch.float32)
                g_b2 = torch.empty(33, device=dev, dtype=torch.float32)
                _lib.check(L.tpr_unpack_decoder_grad(p(g_dec), *ctx.gains, p(g_w1), p(g_b1), p(g_w2), p(g_b2), st()),
                           'tpr_unpack_decoder_grad')
            g_nchw = None
            if want_planes:
                # the plane gradient was scattered in the packed layout [N,3,H,W,32]; the backbone wants [N,3,32,H,W] contiguous
                # (autograd would otherwise make that copy itself, strided: 0.22 ms vs 0.08 ms for 201 MB at config 2)
                g_nchw = torch.empty((n, 3, 32, pp.height, pp.width), device=dev, dtype=torch.float32)
                _lib.check(L.tpr_unpack_planes(p(g_planes), n, pp.height, pp.width, p(g_nchw), st()), 'tpr_unpack_planes')
        return (None, None, None, None, g_nchw,
                g_w1 if need[5] else None, g_b1 if need[6] else None, g_w2 if need[7] else None, g_b2 if need[8] else None,
                None, None)


def render_with_grad(renderer, planes, decoder, ray_origins, ray_directions, options, noise=None):
    """ImportanceRenderer.forward with autograd: returns (rgb, depth, weight_sum) attached to ``planes`` and to the
    decoder's parameters."""
    if isinstance(planes, _r.PackedPlanes):
        raise NotImplementedError('autograd needs the planes tensor itself ([N,3,32,H,W]), not PackedPlanes')
    for t, name in ((ray_origins, 'ray_origins'), (ray_directions, 'ray_directions')):
        if isinstance(t, torch.Tensor) and t.requires_grad:
            raise NotImplementedError(f'{name} requires grad: the renderer differentiates w.r.t. planes and decoder only')
    if options.get('output_layout', 'channels_last') != 'channels_last' or int(options.get('depth_clamp_group', 0)) != 0:
        raise NotImplementedError("autograd supports output_layout='channels_last' and one depth-clamp range per call")
    if planes.shape[0] != ray_origins.shape[0]:
        raise NotImplementedError('autograd needs one plane set per camera (no frame batching)')
    _r.pack_decoder(decoder, allow_grad=True)            # validates the decoder's structure before anything is launched
    fc1, fc2 = decoder.net[0], decoder.net[2]
    return _Render.apply(renderer, decoder, options, noise, planes, fc1.weight, fc1.bias, fc2.weight, fc2.bias,
                         ray_origins, ray_directions)
