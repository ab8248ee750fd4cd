"""GPU: the edges of the hot path's input space, each against the oracle on the same seeded inputs (1e-4 gate) or against
the reference's own failure behaviour.

The reference has no tests; these are the shapes its code admits: one ray, the smallest and the
largest sample counts, non-square planes, ray counts that fill no ray group, no rays at all."""
import numpy as np
import pytest
import torch

from oracle import triplane_oracle as O
from tests.test_gpu_parity import T, make_decoder, TOL, dev

pytestmark = pytest.mark.gpu


def _scene(seed, n_img, n_rays, h, w, dc, df):
    """Random planes of any H x W, rays from the orbit cameras (the first n_rays of a 16 x 16 image), seeded draws."""
    rng = np.random.RandomState(seed)
    planes = rng.standard_normal((n_img, 3, 32, h, w)).astype(np.float32)
    dec = O.make_decoder_params(rng, 1.0, 0.5)
    c2w, K = O.orbit_cameras(n_img)
    o, d = O.ray_sample(c2w, K, 16)
    o, d = np.ascontiguousarray(o[:, :n_rays]), np.ascontiguousarray(d[:, :n_rays])
    below_one = np.nextafter(np.float32(1), np.float32(0))
    jitter = np.minimum(rng.random_sample((n_img, n_rays, dc, 1)).astype(np.float32), below_one)
    u = np.minimum(rng.random_sample((n_img * n_rays, max(df, 1))).astype(np.float32), below_one)[:, :df]
    return planes, dec, o, d, jitter, np.ascontiguousarray(u)


def _check(pkg, planes, dec, o, d, jitter, u, opts, modes=('fp32',)):
    want = O.render(planes, dec, o, d, opts, jitter, u)
    for mode in modes:
        got = pkg.ImportanceRenderer()(T(planes), make_decoder(pkg, dec), T(o), T(d), dict(opts, decoder_precision=mode),
                                       noise=(T(jitter), T(u) if u.size else None))
        errs = [float(np.abs(g.cpu().numpy() - w).max()) for g, w in zip(got, want)]
        assert max(errs) < (TOL if mode != 'bf16' else 5e-2), (mode, errs)
    return want


@pytest.mark.parametrize('n_rays', [1, 3, 7, 9, 250])
def test_ray_counts_that_fill_no_group(pkg, n_rays):
    """A ray group is 8 (or 4) rays; M = 1, 3, 7, 9 and 250 leave partial groups, M = 1 a single warp's worth of work."""
    opts = dict(O.FFHQ_OPTIONS)
    _check(pkg, *_scene(31 + n_rays, 2 if n_rays < 16 else 1, n_rays, 32, 32, 48, 48), opts, modes=('fp32', 'fp32_ffma'))


@pytest.mark.parametrize('hw', [(16, 24), (40, 8), (1, 64), (7, 5)])
def test_non_square_and_tiny_planes(pkg, hw):
    """grid_sample takes any H x W (VR/renderer.py:55-65); so do the taps, including one-texel-high planes where every row tap
    but one is padding."""
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=12, depth_resolution_importance=12)
    _check(pkg, *_scene(41, 1, 64, hw[0], hw[1], 12, 12), opts, modes=('fp32', 'fp32_ffma'))


@pytest.mark.parametrize('dc,df', [(2, 0), (4, 1), (128, 128), (255, 0), (200, 56)])
def test_smallest_and_largest_sample_counts(pkg, dc, df):
    """Dc = 2 is the smallest stratified row (VR/renderer.py:183-188 divides by Dc - 1); 256 samples per ray is the largest
    row the fused kernels sort.  Whatever kernel the dispatcher picks, the result is the oracle's."""
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df)
    _check(pkg, *_scene(51 + dc, 1, 6, 32, 32, dc, df), opts)


@pytest.mark.parametrize('dc', [2, 3])
def test_importance_sampling_needs_four_coarse_samples(pkg, dc):
    """sample_importance drops the first and last of the Dc - 1 smoothed weights (VR/renderer.py:204-208): with Dc < 4 there
    is no bin left to sample from and the reference fails inside sample_pdf; here the call is refused up front."""
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=2)
    planes, dec, o, d, jitter, u = _scene(55, 1, 8, 16, 16, dc, 2)
    with pytest.raises(RuntimeError):
        pkg.ImportanceRenderer()(T(planes), make_decoder(pkg, dec), T(o), T(d), opts, noise=(T(jitter), T(u)))


def test_more_than_256_samples_per_ray_is_refused_loudly(pkg):
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=200, depth_resolution_importance=100)
    planes, dec, o, d, jitter, u = _scene(61, 1, 8, 16, 16, 200, 100)
    with pytest.raises(RuntimeError):
        pkg.ImportanceRenderer()(T(planes), make_decoder(pkg, dec), T(o), T(d), opts, noise=(T(jitter), T(u)))


def test_no_rays_raises_like_the_reference(pkg):
    """With M = 0 the reference fails in the ray marcher's depth.min() of an empty tensor (VR/ray_marcher.py:50); here the
    call is refused before any launch.  Either way: an exception, not an empty result."""
    planes, dec, _, _, _, _ = _scene(71, 1, 8, 16, 16, 8, 8)
    empty = torch.zeros((1, 0, 3), device=dev())
    with pytest.raises((RuntimeError, ValueError, AssertionError)):
        pkg.ImportanceRenderer()(T(planes), make_decoder(pkg, dec), empty, empty, dict(O.FFHQ_OPTIONS))


def test_run_model_single_point_and_points_far_outside_the_box(pkg):
    """One query point; points far outside the box sample only padding (zeros): sigma / rgb are the decoder's response to a zero
    feature (VR/renderer.py:55-65, padding_mode='zeros')."""
    planes, dec, _, _, _, _ = _scene(81, 1, 8, 16, 16, 8, 8)
    opts = dict(O.FFHQ_OPTIONS)
    R = pkg.ImportanceRenderer()
    for pts in (np.array([[[0.1, -0.2, 0.05]]], np.float32), np.full((1, 5, 3), 40.0, np.float32)):
        out = R.run_model(T(planes), make_decoder(pkg, dec), T(pts), None, opts)
        rgb_o, sig_o = O.run_model(planes, dec, pts, opts['box_warp'])
        assert float(np.abs(out['rgb'].cpu().numpy() - rgb_o).max()) < TOL
        assert float(np.abs(out['sigma'].cpu().numpy() - sig_o).max()) < TOL
