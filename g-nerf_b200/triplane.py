"""OSGDecoder with the reference's interface (training/triplane.py:111-136) and parameter names
(`net.0.weight`, `net.0.bias`, `net.2.weight`, `net.2.bias`), so state_dicts are interchangeable.

Inside ImportanceRenderer the decoder never runs as a separate step (it is fused behind the plane
gather); calling this module directly decodes an already gathered [N,3,M,32] feature tensor with
the same CUDA decoder."""
import ctypes

import numpy as np
import torch

from . import _lib
from .volumetric_rendering.renderer import pack_decoder, _require_cuda_f32, _forbid_autograd, _mlp_flag


class FullyConnectedLayer(torch.nn.Module):
    """Parameter container matching training/networks_stylegan2.py:103-119 (linear activation only:
    the decoder's two layers are both `activation='linear'`, so the bias_act plugin is never reached)."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        if activation != 'linear':
            raise NotImplementedError('only linear FullyConnectedLayers occur on the renderer hot path')
        self.in_features, self.out_features, self.activation = in_features, out_features, activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class OSGDecoder(torch.nn.Module):
    def __init__(self, n_features, options):
        super().__init__()
        if n_features != 32 or options.get('decoder_output_dim', 32) != 32:
            raise NotImplementedError('the B200 decoder kernel is specialised for 32 features -> 64 -> 1+32')
        self.hidden_dim = 64
        self.net = torch.nn.Sequential(
            FullyConnectedLayer(n_features, self.hidden_dim, lr_multiplier=options['decoder_lr_mul']),
            torch.nn.Softplus(),
            FullyConnectedLayer(self.hidden_dim, 1 + options['decoder_output_dim'], lr_multiplier=options['decoder_lr_mul']))
        self.precision = 'fp32'

    def forward(self, sampled_features, ray_directions=None):
        """sampled_features [N,3,M,32] -> {'rgb': [N,M,32], 'sigma': [N,M,1]}; ray_directions unused
        (as in the reference, training/triplane.py:124-136)."""
        f = _require_cuda_f32(sampled_features, 'sampled_features', (32,))
        if f.dim() != 4 or f.shape[1] != 3:
            raise RuntimeError(f'sampled_features must be [N,3,M,32], got {tuple(f.shape)}')
        _forbid_autograd(f)
        dec = pack_decoder(self)
        n, _, m, _ = f.shape
        dev = f.device
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        with torch.cuda.device(dev):
            rgb = torch.empty((n, m, 32), device=dev, dtype=torch.float32)
            sigma = torch.empty((n, m, 1), device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().tpr_decode(p(f), n, m, p(dec), p(rgb), p(sigma),
                                             _mlp_flag({'decoder_precision': self.precision}),
                                             ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'tpr_decode')
        return {'rgb': rgb, 'sigma': sigma}
