"""GPU: ImportanceRenderer.forward_host / tpr_render_host -- the forward with HOST buffers, pipelined image by image
over the library's copy streams -- must return exactly what forward() returns for device tensors."""
import numpy as np
import pytest
import torch

from oracle import triplane_oracle as O
from tests.test_gpu_parity import T, make_decoder, TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('n_img,res,dc,df', [(1, 8, 48, 48), (3, 16, 48, 48), (5, 9, 20, 13), (2, 8, 32, 0)])
def test_forward_host_equals_forward(pkg, mode, n_img, res, dc, df):
    scene = O.synthetic_scene(31 + n_img, n_img, res, 64, dc, df, 0.5)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, decoder_precision=mode)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    noise = (T(scene['jitter']), T(scene['u']))
    want = R(T(scene['planes']), dec, T(scene['origins']), T(scene['dirs']), opts, noise=noise)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    for _ in range(2):                                   # second call reuses the workspace and the copy streams
        got = R.forward_host(pin(scene['planes']), dec, pin(scene['origins']), pin(scene['dirs']), opts, noise=noise)
        torch.cuda.current_stream().synchronize()
        for g, w in zip(got, want):
            assert not g.is_cuda and g.is_pinned()
            torch.testing.assert_close(g, w.cpu(), rtol=0, atol=0)
    if mode == 'fp32':
        ref = O.render(scene['planes'], scene['dec'], scene['origins'], scene['dirs'], dict(O.FFHQ_OPTIONS, depth_resolution=dc,
                       depth_resolution_importance=df), scene['jitter'], scene['u'])
        for g, w in zip(got, ref):
            assert np.abs(g.numpy() - w).max() < TOL


def test_forward_host_rejects_device_tensors_and_auto_limits(pkg):
    scene = O.synthetic_scene(5, 1, 4, 32, 16, 16, 0.5)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=16, depth_resolution_importance=16)
    cpu = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    with pytest.raises(RuntimeError):
        R.forward_host(T(scene['planes']), dec, cpu(scene['origins']), cpu(scene['dirs']), opts)
    with pytest.raises(NotImplementedError):
        R.forward_host(cpu(scene['planes']), dec, cpu(scene['origins']), cpu(scene['dirs']),
                       dict(opts, ray_start='auto', ray_end='auto'))


def test_render_host_argument_errors(pkg):
    L = pkg._lib.lib()
    assert L.tpr_render_host_workspace_bytes(0, 64, 64, 16) == 0
    assert L.tpr_render_host_workspace_bytes(2, 64, 64, 16) > 2 * 2 * 3 * 32 * 64 * 64 * 4
    rc = L.tpr_render_host(None, 1, 64, 64, None, None, None, 16, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b'NULL' in L.tpr_last_error()


def test_forward_host_deferred_depth(pkg):
    """Rays sharded over GPUs: depth stays on the device until the caller supplies the all-reduced range.  With this
    GPU's own range the result must be forward_host's; with a narrower range the clamp must follow it."""
    scene = O.synthetic_scene(41, 3, 12, 64, 48, 48, 0.5)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=48, depth_resolution_importance=48)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    noise = (T(scene['jitter']), T(scene['u']))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    args = (pin(scene['planes']), dec, pin(scene['origins']), pin(scene['dirs']), opts)
    res = R.forward_host(*args, noise=noise)
    torch.cuda.current_stream().synchronize()            # the outputs land asynchronously
    want = tuple(t.clone() for t in res)
    got = R.forward_host(*args, noise=noise, defer_depth=True)
    rng = R.last_depth_range.clone()
    R.finish_host_depth(rng)
    torch.cuda.current_stream().synchronize()
    for g, w in zip(got, want):
        torch.testing.assert_close(g, w, rtol=0, atol=0)
    got = R.forward_host(*args, noise=noise, defer_depth=True)
    narrow = torch.tensor([2.6, 2.9], device=rng.device)
    R.finish_host_depth(narrow)
    torch.cuda.current_stream().synchronize()
    torch.testing.assert_close(got[1], want[1].clamp(2.6, 2.9), rtol=0, atol=0)
    with pytest.raises(RuntimeError):
        R.finish_host_depth(narrow)                      # nothing pending any more
