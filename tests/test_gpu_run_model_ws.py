"""GPU: the warp-specialised tensor-core run_model kernel (tpr_run_model_ws.cu) -- density grids and
TriPlaneGenerator.sample / sample_mixed queries (training/triplane.py:92-104, gen_videos.py:33-55,198-209) --
against the oracle, the reference fixtures and the fp32 FFMA kernel it replaces above 65 536 points."""
import numpy as np
import pytest
import torch

from oracle import triplane_oracle as O
from tests.cases import SCALAR_CASES as CASES, load_case
from tests.test_gpu_parity import T, make_decoder, psnr, dev

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_ws(monkeypatch):
    monkeypatch.setenv('TPR_RM_WS_MIN_POINTS', '1')


@pytest.mark.parametrize('name', list(CASES))
def test_small_queries_through_the_tensor_core_kernel(pkg, force_ws, name):
    scene, opts, gold = load_case(name)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    out = R.run_model(T(scene['planes']), dec, T(gold['pts']), None, opts)
    rgb_o, sig_o = O.run_model(scene['planes'], scene['dec'], gold['pts'], opts['box_warp'])
    for got, want in ((out['rgb'], rgb_o), (out['sigma'], sig_o), (out['rgb'], gold['pts_rgb']), (out['sigma'], gold['pts_sigma'])):
        assert np.abs(got.cpu().numpy() - want).max() < 1e-4            # fp32 mode (2xFP16 operand pairs): the north_star tolerance
    sig_only = R.run_model(T(scene['planes']), dec, T(gold['pts']), None, opts, want_rgb=False)
    assert sig_only['rgb'] is None
    torch.testing.assert_close(sig_only['sigma'], out['sigma'], rtol=0, atol=0)
    bf = R.run_model(T(scene['planes']), dec, T(gold['pts']), None, dict(opts, decoder_precision='bf16'))
    assert psnr(bf['rgb'].cpu().numpy(), rgb_o, 1.0) >= 50.0


@pytest.mark.parametrize('n_img,n_pts', [(1, 1), (1, 127), (1, 128), (1, 129), (3, 1000), (2, 148 * 128 * 3 + 77)])
def test_ragged_point_counts(pkg, force_ws, n_img, n_pts):
    """Partial last tiles, more and fewer tiles than SMs, several images: identical (to fp32 rounding) to the FFMA kernel."""
    rng = np.random.default_rng(n_pts)
    scene = O.synthetic_scene(90, n_img, 4, 40, 8, 8, 0.5)
    pts = rng.uniform(-0.6, 0.6, size=(n_img, n_pts, 3)).astype(np.float32)        # some outside the box (zero padding)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    opts = dict(O.FFHQ_OPTIONS)
    out = R.run_model(T(scene['planes']), dec, T(pts), None, opts)
    ref = R.run_model(T(scene['planes']), dec, T(pts), None, dict(opts, decoder_precision='fp32_ffma'))
    assert out['rgb'].shape == (n_img, n_pts, 32) and out['sigma'].shape == (n_img, n_pts, 1)
    assert float((out['rgb'] - ref['rgb']).abs().max()) < 2e-5
    assert float((out['sigma'] - ref['sigma']).abs().max()) < 5e-5
    # sigma-only queries run a different role layout (24 gather warps in three teams, sixteen rows per warp, no COLOUR role):
    # same tiles, same arithmetic, so the same bits -- also with fewer tiles than teams and a partial last tile
    sig_only = R.run_model(T(scene['planes']), dec, T(pts), None, opts, want_rgb=False)
    assert sig_only['rgb'] is None and sig_only['sigma'].shape == (n_img, n_pts, 1)
    torch.testing.assert_close(sig_only['sigma'], out['sigma'], rtol=0, atol=0)
    if n_pts <= 1000:
        rgb_o, sig_o = O.run_model(scene['planes'], scene['dec'], pts, opts['box_warp'])
        assert np.abs(out['rgb'].cpu().numpy() - rgb_o).max() < 1e-4 and np.abs(out['sigma'].cpu().numpy() - sig_o).max() < 1e-4


def test_density_grid_slab_uses_the_kernel_by_default(pkg):
    """A 64^3 slab of the gen_videos.py:33-55 grid (262 144 points: above the threshold, no env override):
    sigma equals the FFMA kernel's, a subset equals the oracle's, and the call is deterministic."""
    scene = O.synthetic_scene(91, 1, 4, 64, 8, 8, 0.5)
    g = 64
    ax = (np.arange(g, dtype=np.float32) + 0.5) / g - 0.5
    zz, yy, xx = np.meshgrid(ax, ax, ax, indexing='ij')
    pts = np.stack([xx, yy, zz], -1).reshape(1, -1, 3).astype(np.float32)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    opts = dict(O.FFHQ_OPTIONS)
    a = R.run_model(T(scene['planes']), dec, T(pts), None, opts, want_rgb=False)['sigma']
    b = R.run_model(T(scene['planes']), dec, T(pts), None, opts, want_rgb=False)['sigma']
    f = R.run_model(T(scene['planes']), dec, T(pts), None, dict(opts, decoder_precision='fp32_ffma'), want_rgb=False)['sigma']
    assert torch.equal(a, b)
    assert float((a - f).abs().max()) < 5e-5
    sub = np.random.default_rng(0).choice(g ** 3, 2000, replace=False)
    _, sig_o = O.run_model(scene['planes'], scene['dec'], pts[:, sub], opts['box_warp'])
    assert np.abs(a.cpu().numpy()[:, sub] - sig_o).max() < 1e-4
