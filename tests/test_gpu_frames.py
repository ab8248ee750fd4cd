"""GPU: the rows either side of the hot path -- many frames of one plane set in one call
(plane_sets / depth_clamp_group), the channels-first feature image, the repacked-plane cache and the stand-alone
sample_stratified -- checked against the per-frame forward() they replace and against the oracle."""
import numpy as np
import pytest
import torch

from oracle import triplane_oracle as O
from tests.cases import CASES, load_case
from tests.test_gpu_parity import T, make_decoder, TOL, dev

pytestmark = pytest.mark.gpu


def _orbit(n_frames):
    c2w, K = O.orbit_cameras(n_frames)
    return c2w.astype(np.float32), K.astype(np.float32)


@pytest.mark.parametrize('mode', ['fp32', 'bf16', 'fp32_ffma'])
@pytest.mark.parametrize('P,F,res,dc,df', [(1, 5, 8, 48, 48), (2, 3, 8, 24, 24), (3, 2, 6, 20, 13), (1, 3, 8, 32, 0)])
def test_render_frames_equals_the_frame_loop(pkg, mode, P, F, res, dc, df):
    """F frames x P identities in one call == F reference-style forwards of a P-batch, bit for bit: same planes per
    identity, per-frame depth clamp, per-frame RNG draws in the reference's order, channels-first feature image."""
    scene = O.synthetic_scene(40 + P, P, res, 48, dc, df, 0.5)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, decoder_precision=mode)
    R, S, dec = pkg.ImportanceRenderer(), pkg.RaySampler(), make_decoder(pkg, scene['dec'])
    planes = T(scene['planes'])
    c2w, K = _orbit(F)
    c2w_t, K_t = T(c2w), T(K)
    # the frame loop (gen_videos.py:153-171): every frame is one forward over the P identities with the same camera
    torch.manual_seed(77)
    loop = []
    for f in range(F):
        o, d = S(c2w_t[f:f + 1].expand(P, -1, -1).contiguous(), K_t[f:f + 1].expand(P, -1, -1).contiguous(), res)
        rgb, depth, wsum = R(planes, dec, o, d, opts)
        feat = rgb.permute(0, 2, 1).reshape(P, 32, res, res).contiguous()               # training/triplane.py:81
        loop.append((feat, depth.permute(0, 2, 1).reshape(P, 1, res, res), wsum.permute(0, 2, 1).reshape(P, 1, res, res)))
    after_loop = torch.rand(3, device=dev())
    torch.manual_seed(77)
    got = pkg.render_frames(R, planes, dec, c2w_t, K_t, res, opts)
    after_batch = torch.rand(3, device=dev())
    torch.testing.assert_close(after_loop, after_batch, rtol=0, atol=0)                  # generator consumed identically
    assert got['feature_image'].shape == (F, P, 32, res, res) and got['feature_image'].is_contiguous()
    for f in range(F):
        torch.testing.assert_close(got['feature_image'][f], loop[f][0], rtol=0, atol=0)
        torch.testing.assert_close(got['depth_image'][f], loop[f][1], rtol=0, atol=0)
        torch.testing.assert_close(got['weights_image'][f], loop[f][2], rtol=0, atol=0)
    # chunked calls give the same frames
    torch.manual_seed(77)
    got2 = pkg.render_frames(R, planes, dec, c2w_t, K_t, res, opts, frames_per_call=2)
    for k in got:
        torch.testing.assert_close(got2[k], got[k], rtol=0, atol=0)


def test_depth_clamp_groups_use_their_own_range(pkg):
    """Empty space (weights 0 -> depth NaN -> inf -> clamped to the max sample depth, VR/ray_marcher.py:49-50): with a
    clamp group per frame every frame lands on ITS OWN maximum, with one group on the call-wide one."""
    scene = O.synthetic_scene(3, 1, 8, 32, 16, 16, 0.5)
    d0 = scene['dec']
    empty = O.DecoderParams(d0.w1 * 0, d0.b1 * 0, d0.w2 * 0, np.concatenate([[-1e4], np.zeros(32)]).astype(np.float32))
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=16, depth_resolution_importance=16)
    R, S, dec = pkg.ImportanceRenderer(), pkg.RaySampler(), make_decoder(pkg, empty)
    c2w, K = _orbit(4)
    o, d = S(T(c2w), T(K), 8)
    g = torch.Generator(device=dev()).manual_seed(5)
    jitter = torch.rand((4, 64, 16, 1), device=dev(), generator=g)
    jitter[1] *= 0.25                                                                  # frame 1 never reaches far
    u = torch.rand((4 * 64, 16), device=dev(), generator=g)
    planes = T(scene['planes'])
    _, depth_all, _ = R(planes, dec, o, d, opts, noise=(jitter, u))
    _, depth_grp, _ = R(planes, dec, o, d, dict(opts, depth_clamp_group=1), noise=(jitter, u))
    per_frame_max = O.stratified_depths(jitter.cpu().numpy(), 2.25, 3.3)[:, :, -1, 0].max(axis=1)
    assert (depth_all == float(per_frame_max.max())).all()
    for f in range(4):
        assert (depth_grp[f] == float(per_frame_max[f])).all()
    assert per_frame_max[1] < per_frame_max[0]
    lo, hi = R.last_depth_range.tolist()                                               # still the call-wide range
    assert hi == float(per_frame_max.max())


@pytest.mark.parametrize('name', ['ffhq_small', 'ragged', 'coarse_only'])
@pytest.mark.parametrize('mode', ['fp32', 'fp32_ffma'])
def test_channels_first_output(pkg, name, mode):
    scene, opts, gold = load_case(name)
    opts = dict(opts, decoder_precision=mode)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    args = (T(scene['planes']), dec, T(scene['origins']), T(scene['dirs']))
    noise = (T(scene['jitter']), T(scene['u']))
    rgb, depth, wsum = R(*args, opts, noise=noise)
    rgb_cf, depth_cf, wsum_cf = R(*args, dict(opts, output_layout='channels_first'), noise=noise)
    n, m, _ = rgb.shape
    assert rgb_cf.shape == (n, m, 32) and rgb_cf.permute(0, 2, 1).is_contiguous() and not rgb_cf.is_contiguous()
    torch.testing.assert_close(rgb_cf, rgb, rtol=0, atol=0)
    torch.testing.assert_close(depth_cf, depth, rtol=0, atol=0)
    # what TriPlaneGenerator.synthesis does next (training/triplane.py:81) is now free: no copy is made
    res = int(round(m ** 0.5))
    if res * res == m:
        img = rgb_cf.permute(0, 2, 1).reshape(n, 32, res, res)
        assert img.is_contiguous() and img.contiguous().data_ptr() == rgb_cf.data_ptr()
    assert np.abs(rgb_cf.cpu().numpy() - gold['rgb']).max() < TOL
    with pytest.raises(RuntimeError, match='output_layout'):
        R(*args, dict(opts, output_layout='nhwc'), noise=noise)


def test_packed_plane_cache(pkg):
    scene, opts, _ = load_case('ragged')
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    planes = T(scene['planes'])
    args = (dec, T(scene['origins']), T(scene['dirs']), opts)
    noise = (T(scene['jitter']), T(scene['u']))
    want = R(planes, *args, noise=noise)
    R.cache_packed_planes = True
    a = R(planes, *args, noise=noise)
    packed = R._plane_cache[2]
    b = R(planes.view(planes.shape), *args, noise=noise)                       # a fresh view of the same storage hits
    assert R._plane_cache[2] is packed
    for x, y, z in zip(a, b, want):
        torch.testing.assert_close(x, z, rtol=0, atol=0)
        torch.testing.assert_close(y, z, rtol=0, atol=0)
    planes.mul_(0.5)                                                           # in-place edit: version counter moves
    c = R(planes, *args, noise=noise)
    assert R._plane_cache[2] is not packed
    want2 = pkg.ImportanceRenderer()(planes, *args, noise=noise)
    for x, y in zip(c, want2):
        torch.testing.assert_close(x, y, rtol=0, atol=0)
    assert not torch.equal(c[0], a[0])


@pytest.mark.parametrize('name', list(CASES))
def test_sample_stratified_bit_exact(pkg, name):
    """The coarse depths of every render case: stand-alone kernel == oracle, bit for bit."""
    scene, opts, gold = load_case(name)
    R = pkg.ImportanceRenderer()
    dc = opts['depth_resolution']
    rs, re = opts['ray_start'], opts['ray_end']
    if isinstance(rs, str):                                                  # 'auto': per-ray tensor limits (VR/renderer.py:91-97)
        rs, re = (T(a) for a in O.auto_ray_limits(scene['origins'], scene['dirs'], opts['box_warp']))
    got = R.sample_stratified(T(scene['origins']), rs, re, dc, opts.get('disparity_space_sampling', False), jitter=T(scene['jitter']))
    from tests.cases import coarse_depths
    want = coarse_depths(scene, opts)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_sample_stratified_all_branches(pkg):
    """Scalar, disparity-space and per-ray limits (VR/renderer.py:169-192): bit-exact against the oracle; within 2 ulp
    of the reference's own torch expressions evaluated on this GPU and of the CPU-generated reference fixture (torch's
    CPU linspace is vectorised, its CUDA one FMA-contracted: see oracle/triplane_oracle.py:torch_linspace)."""
    import os
    from tests.cases import GOLDEN_DIR
    from tests.stratified_cases import CASES as SC, inputs
    gold = np.load(os.path.join(GOLDEN_DIR, 'stratified.npz'))
    R = pkg.ImportanceRenderer()
    for name, (n, m, d, rs, re, disp) in SC.items():
        jitter, lim = inputs(name)
        o = torch.zeros((n, m, 3), device=dev())
        jt = T(jitter)
        if lim is None:
            got = R.sample_stratified(o, rs, re, d, disp, jitter=jt)
            want = O.stratified_depths(jitter, rs, re, disp)
            if disp:                                                                    # VR/renderer.py:175-181 on CUDA
                live = torch.linspace(0, 1, d, device=dev()).reshape(1, 1, d, 1).repeat(n, m, 1, 1)
                live += jt * (1 / (d - 1))
                live = 1. / (1. / rs * (1. - live) + 1. / re * live)
            else:                                                                       # :188-190 on CUDA
                live = torch.linspace(rs, re, d, device=dev()).reshape(1, 1, d, 1).repeat(n, m, 1, 1)
                live += jt * ((re - rs) / (d - 1))
            # torch's CUDA linspace / add(alpha) kernels are compiled with FMA contraction, the CPU ones are not: the two
            # reference targets differ from each other in the last bits, and the kernels follow the unfused arithmetic
            assert np.abs(got.cpu().numpy() - live.cpu().numpy()).max() <= 2 * np.spacing(np.float32(gold[name].max()))
            assert np.abs(got.cpu().numpy() - gold[name]).max() <= 2 * np.spacing(np.float32(gold[name].max()))
        else:
            got = R.sample_stratified(o, T(lim[0]), T(lim[1]), d, jitter=jt)
            want = O.stratified_depths_per_ray(jitter, *lim)
            np.testing.assert_array_equal(got.cpu().numpy(), gold[name])
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    # without `jitter` the draw is torch.rand of the reference's shape, from the current generator
    torch.manual_seed(4)
    a = R.sample_stratified(torch.zeros((2, 9, 3), device=dev()), 2.25, 3.3, 12)
    torch.manual_seed(4)
    b = R.sample_stratified(torch.zeros((2, 9, 3), device=dev()), 2.25, 3.3, 12, jitter=torch.rand((2, 9, 12, 1), device=dev()))
    torch.testing.assert_close(a, b, rtol=0, atol=0)
    with pytest.raises(RuntimeError):
        R.sample_stratified(torch.zeros((1, 4, 3), device=dev()), torch.zeros((1, 4, 1), device=dev()), 3.3, 8)
