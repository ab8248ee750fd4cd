"""CPU: the numpy oracle against known-answer tensors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import numpy as np
import pytest

from tests.cases import CASES, density_noise_draws, load_case
from oracle import triplane_oracle as O

TOL = 1e-5      # oracle vs reference (both fp32 CPU; differences are summation order / libm)


@pytest.mark.parametrize('name', list(CASES))
def test_ray_sampler_matches_reference(name):
    scene, _, gold = load_case(name)
    np.testing.assert_array_equal(scene['origins'], gold['origins'])
    np.testing.assert_allclose(scene['dirs'], gold['dirs'], atol=2e-7, rtol=0)


@pytest.mark.parametrize('name', list(CASES))
def test_render_matches_reference(name):
    scene, opts, gold = load_case(name)
    draws = density_noise_draws(name)[:2] if opts.get('density_noise', 0) > 0 else None
    (rgb, depth, wsum), st = O.render(scene['planes'], scene['dec'], gold['origins'], gold['dirs'], opts,
                                      scene['jitter'], scene['u'], return_stages=True, density_noise_draws=draws)
    assert np.abs(rgb - gold['rgb']).max() < TOL
    assert np.abs(depth - gold['depth']).max() < TOL
    assert np.abs(wsum - gold['wsum']).max() < TOL
    if opts['depth_resolution_importance'] > 0:
        assert np.abs(st['weights_coarse'] - gold['weights_coarse']).max() < TOL
        assert np.abs(st['depths_fine'] - gold['depths_fine']).max() < TOL
        # searchsorted indices from the full pipeline (weights differ in the last ulp, so allow the
        # rare flip; the stage-wise test below is exact)
        flips = int((st['inds'] != gold['inds']).sum())
        assert flips <= max(1, 1e-4 * gold['inds'].size), flips       # (one flip in a 6400-draw case is 1.6e-4)


@pytest.mark.parametrize('name', [n for n in CASES if CASES[n][5] > 0])
def test_sample_pdf_indices_bit_exact_given_identical_inputs(name):
    """Feed the oracle the reference's own (bins, weights, u): indices must be identical and samples
    equal to the last ulp (bit-exact only makes sense stage-wise)."""
    scene, _, gold = load_case(name)
    samples, inds = O.sample_pdf(gold['pdf_bins'], gold['pdf_weights'], scene['u'])
    np.testing.assert_array_equal(inds, gold['inds'])
    ref = gold['depths_fine'].reshape(samples.shape)
    assert np.abs(samples - ref).max() <= 4.8e-7            # 2 ulp at depth ~3


@pytest.mark.parametrize('name', list(CASES))
def test_run_model_matches_reference(name):
    scene, opts, gold = load_case(name)
    dn = opts.get('density_noise', 0)
    rgb, sigma = O.run_model(scene['planes'], scene['dec'], gold['pts'], opts['box_warp'], dn,
                             density_noise_draws(name)[2] if dn > 0 else None)
    assert np.abs(rgb - gold['pts_rgb']).max() < TOL
    assert np.abs(sigma - gold['pts_sigma']).max() < 2e-5


@pytest.mark.parametrize('name', ['ffhq_small', 'white_back'])
def test_marcher_matches_reference(name):
    _, opts, gold = load_case(name)
    rgb, depth, w = O.march(gold['march_colors'], gold['march_sigma'], gold['march_depths'],
                            bool(opts.get('white_back', False)))
    assert np.abs(rgb - gold['march_rgb']).max() < TOL
    assert np.abs(depth - gold['march_depth']).max() < TOL
    assert np.abs(w - gold['march_w']).max() < TOL


def test_ray_limits_box_and_auto_limits_match_reference():
    """a14: get_ray_limits_box (VR/math_utils.py:46-98) on the case's camera rays (a box the outer rays miss) and on random
    rays around a unit box, bit for bit; the 'auto' replacement of the misses (VR/renderer.py:93-96) through render above."""
    scene, opts, gold = load_case('auto_limits')
    tmin, tmax = O.ray_limits_box(gold['origins'], gold['dirs'], opts['box_warp'])
    np.testing.assert_array_equal(tmin, gold['box_tmin'])
    np.testing.assert_array_equal(tmax, gold['box_tmax'])
    misses = int((gold['box_tmin'] == -1).sum())
    assert 0 < misses < tmin.size                      # the fixture exercises both kinds of ray
    tmin, tmax = O.ray_limits_box(gold['box2_origins'], gold['box2_dirs'], 1.0)
    np.testing.assert_array_equal(tmin, gold['box2_tmin'])
    np.testing.assert_array_equal(tmax, gold['box2_tmax'])
    rs, re = O.auto_ray_limits(gold['origins'], gold['dirs'], opts['box_warp'])
    assert (re[gold['box_tmin'] == -1] == gold['box_tmin'][gold['box_tmin'] != -1].max()).all()


def test_cdf_contract_is_order_independent():
    """The float64-accumulated CDF is exact, hence identical for any summation order -- the property the
    GPU's parallel scan relies on to be bit-identical to this oracle."""
    rng = np.random.RandomState(0)
    w = (rng.random_sample((200, 45)).astype(np.float32) * 0.99 + 0.01)
    cdf = O.pdf_to_cdf(w)
    wr = (w + np.float32(1e-5)).astype(np.float32)
    tot = wr[:, ::-1].astype(np.float64).sum(-1, keepdims=True).astype(np.float32)     # reversed order
    pdf = (wr / tot).astype(np.float32).astype(np.float64)
    # pairwise (tree) order
    tree = np.zeros_like(pdf)
    for j in range(45):
        parts = [pdf[:, :j + 1][:, k::4].sum(-1) for k in range(4)]
        tree[:, j] = (parts[0] + parts[2]) + (parts[1] + parts[3])
    np.testing.assert_array_equal(cdf[:, 1:], tree.astype(np.float32))
    assert (cdf[:, 0] == 0).all() and np.all(np.diff(cdf, axis=1) > 0)


def test_softplus_and_edges():
    x = np.array([-100, -20, -1, 0, 1, 19.9, 20, 20.1, 100], np.float32)
    y = O.softplus(x)
    assert y[0] < 1e-40 and y[-1] == 100 and y[-2] == np.float32(20.1)
    assert abs(y[3] - np.log(2)) < 1e-7
    # zero density everywhere -> zero weights -> depth NaN -> +inf -> clamped to max depth
    d = np.linspace(1, 2, 5, dtype=np.float32).reshape(1, 1, 5, 1)
    rgb, depth, w = O.march(np.ones((1, 1, 5, 3), np.float32), np.full((1, 1, 5, 1), -1e4, np.float32), d)
    assert (w == 0).all() and depth[0, 0, 0] == 2.0 and (rgb == -1).all()


def test_stratified_depths_against_reference_fixture():
    """sample_stratified, all three branches (VR/renderer.py:169-192).  The fixture comes from the reference on CPU, whose
    torch.linspace is vectorised (base + lane*step per SIMD vector): <= 2 ulp from the per-element CUDA formula the oracle
    restates; the per-ray branch (math_utils.linspace, plain tensor arithmetic) is bit-exact."""
    import os
    from tests.cases import GOLDEN_DIR
    from tests.stratified_cases import CASES as SC, inputs
    gold = np.load(os.path.join(GOLDEN_DIR, 'stratified.npz'))
    for name, (n, m, d, rs, re, disp) in SC.items():
        jitter, lim = inputs(name)
        want = gold[name]
        got = O.stratified_depths(jitter, rs, re, disp) if lim is None else O.stratified_depths_per_ray(jitter, *lim)
        assert got.shape == want.shape
        if lim is not None:
            np.testing.assert_array_equal(got, want)
        else:
            assert np.abs(got - want).max() <= 2 * np.spacing(np.float32(want.max()))
