"""GPU: raw tcgen05 decoder plumbing (SS layer 1, TS layer 2) against a float64 matmul."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(pkg, mode, P=1000, seed=0):
    dev = torch.device('cuda:0')
    torch.manual_seed(seed)
    dec = pkg.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).to(dev).requires_grad_(False)
    with torch.no_grad():
        dec.net[0].bias.normal_(); dec.net[2].bias.normal_()
    packed = pkg.pack_decoder(dec)
    x = torch.randn(P, 32, device=dev)
    hidden = torch.full((P, 64), float('nan'), device=dev)
    out = torch.full((P, 48), float('nan'), device=dev)
    L = pkg._lib.bench_lib()            # the raw tcgen05 layer test lives in the measurement library
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = L.tpr_debug_tc_decode(p(x), ctypes.c_int64(P), p(packed), ctypes.c_int32(mode), p(hidden), p(out),
                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    torch.cuda.synchronize()
    pk = packed.double().cpu().numpy()
    w1t = pk[:2048].reshape(32, 64); b1 = pk[2048:2112]
    w2t = pk[2112:2112 + 64 * 36].reshape(64, 36); b2 = pk[4416:4452]
    xd = x.double().cpu().numpy()
    h_pre = xd @ w1t + b1
    h = np.where(h_pre > 20, h_pre, np.log1p(np.exp(np.minimum(h_pre, 20))))
    o = h @ w2t + b2
    eh = np.abs(hidden.cpu().numpy() - h_pre).max()
    eo = np.abs(out.cpu().numpy()[:, :36] - o).max()
    pad = np.abs(out.cpu().numpy()[:, 36:]).max()
    return eh, eo, pad


@pytest.mark.parametrize('mode,tol_h,tol_o', [(0, 5e-3, 1e-2), (1, 5e-6, 1e-5), (2, 5e-2, 8e-2)])
def test_tc_decode_layers(pkg, mode, tol_h, tol_o):
    eh, eo, pad = _run(pkg, mode)
    print(f'mode {mode}: hidden err {eh:.3e}, out err {eo:.3e}, pad {pad:.3e}')
    assert eh < tol_h and eo < tol_o and pad == 0
