"""CPU: pins oracle/torch_oracle.py (the torch restatement used as the full-size checker and GPU baseline on the B200
box) against the fixtures the UNMODIFIED reference produced (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from tests.cases import SCALAR_CASES as CASES, load_case
from oracle import torch_oracle as TO

TOL = 1e-5


@pytest.mark.parametrize('name', list(CASES))
def test_torch_oracle_matches_reference_fixture(name):
    scene, opts, gold = load_case(name)
    t = torch.from_numpy
    dec = TO.decoder_tuple(scene['dec'], 'cpu')
    with torch.no_grad():
        (rgb, depth, wsum), st = TO.render(t(scene['planes']), dec, t(gold['origins']), t(gold['dirs']), opts,
                                           t(scene['jitter']), t(scene['u']), return_stages=True)
        pr, ps = TO.decode(TO.gather(t(scene['planes']), t(gold['pts']), opts['box_warp']), dec)
    # same ATen kernels as the reference on the same machine: bit-identical, not merely close
    np.testing.assert_array_equal(rgb.numpy(), gold['rgb'])
    np.testing.assert_array_equal(depth.numpy(), gold['depth'])
    np.testing.assert_array_equal(wsum.numpy(), gold['wsum'])
    np.testing.assert_array_equal(pr.numpy(), gold['pts_rgb'])
    np.testing.assert_array_equal(ps.numpy(), gold['pts_sigma'])
    if opts['depth_resolution_importance'] > 0:
        np.testing.assert_array_equal(st['inds'].numpy(), gold['inds'])
        np.testing.assert_array_equal(st['depths_fine'].numpy(), gold['depths_fine'])
        np.testing.assert_array_equal(st['weights_coarse'].numpy(), gold['weights_coarse'])


def test_torch_oracle_march_matches_reference_fixture():
    _, opts, gold = load_case('ffhq_small')
    t = torch.from_numpy
    rgb, depth, w = TO.march(t(gold['march_colors']), t(gold['march_sigma']), t(gold['march_depths']))
    np.testing.assert_array_equal(rgb.numpy(), gold['march_rgb'])
    np.testing.assert_array_equal(depth.numpy(), gold['march_depth'])
    np.testing.assert_array_equal(w.numpy(), gold['march_w'])


from tests.cases import BWD_CASES, load_bwd_case


@pytest.mark.parametrize('name', list(BWD_CASES))
def test_torch_oracle_gradients_match_reference_fixture(name):
    """autograd through the restatement == autograd through the unmodified reference (tests/golden/bwd_*.npz,
    made by tests/golden/make_golden_backward.py): pins the gradient oracle the CUDA backward is checked against."""
    scene, opts, gold, (A, B, C) = load_bwd_case(name)
    t = torch.from_numpy
    dec = TO.decoder_tuple(scene['dec'], 'cpu')
    (rgb, depth, wsum), grads = TO.render_grads(t(scene['planes']), dec, t(scene['origins']), t(scene['dirs']), opts,
                                                t(scene['jitter']), t(scene['u']), t(A), t(B), t(C))
    np.testing.assert_array_equal(rgb.numpy(), gold['rgb'])
    for g, k in zip(grads, ('g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2')):
        scale = float(np.abs(gold[k]).max())
        assert float(np.abs(g.numpy() - gold[k]).max()) <= 1e-6 * max(scale, 1.0), k
