"""GPU: same-device parity against the UNMODIFIED reference itself (oracle/_ref/g_nerf, the byte-for-byte copy
oracle/build_ref.py makes in the build container and gpurun ships to the box; see oracle/ref_loader.py).

No noise injection here: the reference and the B200 renderer are both run from the SAME torch CUDA generator state, so
these tests also pin the RNG contract -- the host shim makes the reference's own torch.rand / torch.randn calls with the
reference's shapes in the reference's order (VR/renderer.py:146,190,237), and the generator ends in the same state.

Gates (BASELINE.json north_star): fp32 mode max-abs <= 1e-4 on rgb / depth / weight sum; bf16-MLP mode PSNR >= 50 dB."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import triplane_oracle as O
from tests.test_gpu_parity import T, dev, TOL

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(ref_loader.reference_dir() is None, reason='oracle/_ref (python oracle/build_ref.py) not present')]


@pytest.fixture(scope='module')
def ref():
    torch.backends.cuda.matmul.allow_tf32 = False        # as the reference sets it (training_loop.py:145-146)
    torch.backends.cudnn.allow_tf32 = False
    return ref_loader.import_reference()


def _decoders(pkg, ref, seed=0, bias_scale=0.5):
    """The reference's OSGDecoder and ours with the same parameters (state_dict interchange, both directions)."""
    theirs = ref_loader.make_decoder(seed)
    with torch.no_grad():
        theirs.net[0].bias.normal_(0, bias_scale); theirs.net[2].bias.normal_(0, bias_scale)
    theirs = theirs.to(dev())
    ours = pkg.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    ours.load_state_dict(theirs.state_dict())
    return theirs, ours.to(dev()).requires_grad_(False)


def _psnr(a, b, peak):
    return 10 * np.log10(peak * peak / max(float(((a.double() - b.double()) ** 2).mean()), 1e-30))


def _both(pkg, ref, planes, theirs, ours, o, d, opts, seed=1):
    """(reference outputs, our outputs, RNG states after each) from the same generator state."""
    with torch.no_grad():
        torch.manual_seed(seed)
        want = ref.renderer.ImportanceRenderer()(planes, theirs, o, d, dict(opts))
        state_ref = torch.cuda.get_rng_state(dev())
        torch.manual_seed(seed)
        got = pkg.ImportanceRenderer()(planes, ours, o, d, dict(opts))
        state_ours = torch.cuda.get_rng_state(dev())
    return want, got, state_ref, state_ours


CASES = {
    # name: (n_img, res, plane_res, options on top of FFHQ)
    'config2_pair': (2, 128, 256, {}),
    'inference_96': (1, 64, 256, {'depth_resolution': 96, 'depth_resolution_importance': 96}),      # gen_videos.py:127-128
    'white_back':   (1, 32, 64, {'white_back': True, 'depth_resolution': 24, 'depth_resolution_importance': 24}),
    'auto_limits':  (2, 32, 64, {'ray_start': 'auto', 'ray_end': 'auto', 'box_warp': 0.5}),
    'density_noise': (1, 32, 64, {'density_noise': 0.5}),
    'coarse_only':  (1, 32, 64, {'depth_resolution_importance': 0}),
    'disparity':    (1, 32, 64, {'disparity_space_sampling': True}),
}


@pytest.mark.parametrize('name', list(CASES))
def test_forward_matches_the_reference_on_the_same_gpu(pkg, ref, name):
    n, res, pres, extra = CASES[name]
    opts = dict(O.FFHQ_OPTIONS, **extra)
    g = torch.Generator(device='cpu').manual_seed(sum(map(ord, name)))
    planes = torch.randn((n, 3, 32, pres, pres), generator=g).to(dev())
    theirs, ours = _decoders(pkg, ref)
    c2w, K = O.orbit_cameras(n)
    o_ref, d_ref = ref.ray_sampler.RaySampler()(T(c2w), T(K), res)
    o, d = pkg.RaySampler()(T(c2w), T(K), res)
    assert torch.equal(o, o_ref) and float((d - d_ref).abs().max()) <= 2e-7
    want, got, s_ref, s_ours = _both(pkg, ref, planes, theirs, ours, o_ref, d_ref, opts)
    assert torch.equal(s_ref, s_ours), 'the CUDA generator must advance exactly as the reference advances it'
    errs = {k: float((a - b).abs().max()) for k, a, b in zip(('rgb', 'depth', 'wsum'), got, want)}
    print(f'{name}: {n} x {res}^2 rays, fp32 max-abs vs the reference on this GPU:', errs)
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.isfinite(a).all()
    assert max(errs.values()) < TOL, errs
    # bf16-MLP mode, same generator state
    with torch.no_grad():
        torch.manual_seed(1)
        got16 = pkg.ImportanceRenderer()(planes, ours, o_ref, d_ref, dict(opts, decoder_precision='bf16'))
    peak_d = float(want[1].max() - want[1].min()) if isinstance(opts['ray_start'], str) else opts['ray_end'] - opts['ray_start']
    ps = {'rgb': _psnr(got16[0], want[0], 2.0), 'depth': _psnr(got16[1], want[1], peak_d)}
    print(f'{name}: bf16-MLP PSNR (dB):', ps)
    assert min(ps.values()) >= 50.0, ps


def test_run_model_matches_the_reference_on_a_density_grid_slab(pkg, ref):
    """Config 5's query (gen_videos.py:33-55,198-209): a z-slab of the 256^3 grid through run_model, rgb and sigma of every
    point against the reference's run_model; with density_noise both draw randn_like from the same generator state."""
    g = 256
    ax = (torch.arange(g, device=dev(), dtype=torch.float32) + 0.5) / g - 0.5
    zz, yy, xx = torch.meshgrid(ax[96:104], ax, ax, indexing='ij')
    xyz = torch.stack([xx, yy, zz], -1).reshape(1, -1, 3).contiguous()
    planes = torch.randn((1, 3, 32, 256, 256), generator=torch.Generator().manual_seed(3)).to(dev())
    theirs, ours = _decoders(pkg, ref)
    R_ref = ref.renderer.ImportanceRenderer()
    R_ref.plane_axes = R_ref.plane_axes.to(dev())           # (forward does this at VR/renderer.py:89; run_model relies on it)
    for extra in ({}, {'density_noise': 0.25}):
        opts = dict(O.FFHQ_OPTIONS, **extra)
        with torch.no_grad():
            torch.manual_seed(4)
            want = R_ref.run_model(planes, theirs, xyz, None, opts)
            torch.manual_seed(4)
            got = pkg.ImportanceRenderer().run_model(planes, ours, xyz, None, opts)
        assert float((got['sigma'] - want['sigma']).abs().max()) < TOL
        assert float((got['rgb'] - want['rgb']).abs().max()) < TOL


def test_module_level_helpers_match_the_reference(pkg, ref):
    """a2 / a3 / a4 / a12 / a15: generate_planes, project_onto_planes, sample_from_planes, sample_from_3dgrid, sort_samples
    and unify_samples -- the public names of VR/renderer.py that forward() does not call -- against the reference's own."""
    from importlib import import_module
    mine = import_module('g-nerf_b200.volumetric_rendering.renderer')
    theirs = ref.renderer
    assert torch.equal(mine.generate_planes(), theirs.generate_planes())
    axes = theirs.generate_planes().to(dev())
    rng = torch.Generator().manual_seed(7)
    xyz = (torch.rand((2, 501, 3), generator=rng) - 0.5).mul(1.3).to(dev())           # some outside the box
    torch.testing.assert_close(mine.project_onto_planes(axes, xyz), theirs.project_onto_planes(axes, xyz), rtol=0, atol=1e-6)
    planes = torch.randn((2, 3, 32, 40, 56), generator=rng).to(dev())                 # H != W
    want = theirs.sample_from_planes(axes, planes, xyz, padding_mode='zeros', box_warp=1.0)
    got = mine.sample_from_planes(axes, planes, xyz, padding_mode='zeros', box_warp=1.0)
    assert got.shape == want.shape == (2, 3, 501, 32)
    assert float((got - want).abs().max()) < 2e-5
    # the decoder on those features == run_model
    theirs_dec, ours_dec = _decoders(pkg, ref)
    a = ours_dec(got, None)
    b = pkg.ImportanceRenderer().run_model(planes, ours_dec, xyz, None, {'box_warp': 1.0, 'decoder_precision': 'fp32_ffma'})
    assert float((a['rgb'] - b['rgb']).abs().max()) < 2e-5 and float((a['sigma'] - b['sigma']).abs().max()) < 2e-5
    # 3-D grid, shared and per-batch, points beyond the borders
    for gshape in ((1, 5, 6, 7, 8), (2, 3, 4, 9, 5)):
        grid = torch.randn(gshape, generator=rng).to(dev())
        q = (torch.rand((2, 333, 3), generator=rng) * 2.4 - 1.2).to(dev())
        want = theirs.sample_from_3dgrid(grid, q)
        got = mine.sample_from_3dgrid(grid, q)
        assert got.shape == want.shape and float((got - want).abs().max()) < 1e-5
    # sort_samples / unify_samples: 48 + 48 samples with a few exact ties
    d1 = torch.rand((2, 37, 48, 1), generator=rng).to(dev()) + 2
    d2 = torch.rand((2, 37, 48, 1), generator=rng).to(dev()) + 2
    c1, c2 = torch.randn((2, 37, 48, 32), generator=rng).to(dev()), torch.randn((2, 37, 48, 32), generator=rng).to(dev())
    s1, s2 = torch.randn((2, 37, 48, 1), generator=rng).to(dev()), torch.randn((2, 37, 48, 1), generator=rng).to(dev())
    R_mine, R_theirs = pkg.ImportanceRenderer(), theirs.ImportanceRenderer()
    for got, want in zip(R_mine.unify_samples(d1, c1, s1, d2, c2, s2), R_theirs.unify_samples(d1, c1, s1, d2, c2, s2)):
        assert torch.equal(got, want)
    for got, want in zip(R_mine.sort_samples(d1, c1, s1), R_theirs.sort_samples(d1, c1, s1)):
        assert torch.equal(got, want)
    d1[:, :, 5] = d1[:, :, 4]                            # ties: torch.sort is unstable, so compare depths and the multiset only
    ds, cs, ss = R_mine.sort_samples(d1, c1, s1)
    wd, _, _ = R_theirs.sort_samples(d1, c1, s1)
    assert torch.equal(ds, wd) and torch.equal(ss.sort(dim=-2).values, s1.sort(dim=-2).values)


def test_marcher_and_resampling_match_the_reference(pkg, ref):
    rng = torch.Generator().manual_seed(9)
    n, m, s = 2, 500, 48
    col = torch.rand((n, m, s, 32), generator=rng).to(dev())
    sig = (torch.randn((n, m, s, 1), generator=rng) * 3).to(dev())
    dep = (torch.rand((n, m, s, 1), generator=rng) + 2).sort(dim=-2).values.to(dev())
    for white in (False, True):
        opts = dict(O.FFHQ_OPTIONS, white_back=white)
        want = ref.ray_marcher.MipRayMarcher2()(col, sig, dep, opts)
        got = pkg.MipRayMarcher2()(col, sig, dep, opts)
        for a, b in zip(got, want):
            assert a.shape == b.shape and float((a - b).abs().max()) < 1e-5
    # sample_importance from the same generator state (its one torch.rand, VR/renderer.py:237)
    w = want[2]
    torch.manual_seed(3)
    fine_ref = ref.renderer.ImportanceRenderer().sample_importance(dep, w, 48)
    torch.manual_seed(3)
    fine = pkg.ImportanceRenderer().sample_importance(dep, w, 48)
    # torch's CUDA cumsum is a float32 tree, ours the exactly-rounded sum: a draw within an ulp of a CDF entry may land in the
    # neighbouring bin; everything else agrees to the last ulps
    diff = (fine - fine_ref).abs()
    assert float((diff > 1e-5).float().mean()) < 1e-4 and float(diff.median()) < 5e-7
