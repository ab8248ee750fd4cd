"""GPU: the renderer's backward pass through the C ABI, against
  * the gradient fixtures the UNMODIFIED reference produced under autograd (tests/golden/bwd_*.npz), and
  * autograd through the torch restatement (oracle/torch_oracle.py, pinned to those fixtures on CPU) on the same device
    at sizes the fixtures do not cover.
Tolerance: 2e-4 of the largest gradient entry of each tensor (fp32 accumulation everywhere; the decoder GEMMs run on tcgen05 with fp16 hi + lo
operand pairs -- or as 3xTF32 mma.sync with TPR_BWD_IMPL=hmma --, the plane gradient is accumulated with floating-point atomics, so the summation order differs from torch's)."""
import ctypes

import numpy as np
import pytest
import torch

from tests.cases import BWD_CASES, load_bwd_case
from oracle import triplane_oracle as O
from oracle import torch_oracle as TO
from tests.test_gpu_parity import make_decoder, T, dev

pytestmark = pytest.mark.gpu
REL = 2e-4


def rel_err(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-12))


def run_backward(pkg, scene, opts, A, B, C, keep_samples=True, keep_features=True):
    dec = make_decoder(pkg, scene['dec']).requires_grad_(True)
    planes = T(scene['planes']).requires_grad_(True)
    R = pkg.ImportanceRenderer()
    R.keep_samples = keep_samples
    R.keep_features = keep_features
    rgb, depth, wsum = R(planes, dec, T(scene['origins']), T(scene['dirs']), opts, noise=(T(scene['jitter']), T(scene['u'])))
    loss = (rgb * T(A)).sum() + (depth * T(B)).sum() + (wsum * T(C)).sum()
    loss.backward()
    return (rgb, depth, wsum), (planes.grad, dec.net[0].weight.grad, dec.net[0].bias.grad, dec.net[2].weight.grad,
                                dec.net[2].bias.grad)


@pytest.mark.parametrize('name', list(BWD_CASES))
def test_gradients_match_reference_fixture(pkg, name):
    scene, opts, gold, (A, B, C) = load_bwd_case(name)
    (rgb, depth, wsum), grads = run_backward(pkg, scene, opts, A, B, C)
    assert np.abs(rgb.detach().cpu().numpy() - gold['rgb']).max() < 1e-4
    assert np.abs(depth.detach().cpu().numpy() - gold['depth']).max() < 1e-4
    errs = {}
    for g, k in zip(grads, ('g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2')):
        assert g is not None and tuple(g.shape) == gold[k].shape, k
        assert torch.isfinite(g).all(), k
        errs[k] = rel_err(g.cpu().numpy(), gold[k])
    print(name, {k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < REL, errs


@pytest.mark.parametrize('variant', ['random', 'ties', 'unsorted_coarse', 'odd_counts'])
def test_march_backward_against_autograd(pkg, variant):
    """tpr_march_backward alone: d/d(sigma) and the colour weights omega of random (depth, sigma, colour) rows.  'ties': equal
    depths among the importance samples, between a coarse and an importance sample and inside the coarse row (the kernel's fast
    rank count detects them by its rank sum and falls back to the exact, index-tie-broken count = a stable sort);
    'unsorted_coarse': a coarse row that does not ascend (the general path); 'odd_counts': sample counts that are not multiples
    of four (no vector reads)."""
    torch.manual_seed(5)
    d = dev()
    r, dc, df = (37, 24, 16) if variant != 'odd_counts' else (19, 23, 13)
    s = dc + df
    coarse = torch.sort(torch.rand(r, dc, device=d) * 0.8 + 2.3, -1).values.contiguous()
    fine = (torch.rand(r, df, device=d) * 0.8 + 2.3).contiguous()
    if variant == 'ties':
        fine[:, 3] = fine[:, 7]
        fine[::2, 0] = coarse[::2, 5]
        coarse[1::3, 11] = coarse[1::3, 10]
    if variant == 'unsorted_coarse':
        coarse[::2, [4, 9]] = coarse[::2, [9, 4]]
    sigma = (torch.randn(r, s, device=d) * 3).requires_grad_(True)
    col = torch.rand(r, s, 32, device=d).requires_grad_(True)
    A, B, C = torch.randn(r, 32, device=d), torch.randn(r, device=d), torch.randn(r, device=d)
    for white in (False, True):
        depths = torch.cat([coarse, fine], -1)
        d_all, order = torch.sort(depths, dim=-1, stable=True)      # ties: the earlier sample first (coarse before importance)
        rgb, depth, w = TO.march(torch.gather(col, 1, order.unsqueeze(-1).expand(-1, -1, 32)).unsqueeze(0),
                                 torch.gather(sigma, 1, order).unsqueeze(0).unsqueeze(-1), d_all.unsqueeze(0).unsqueeze(-1), white)
        loss = (rgb[0] * A).sum() + (depth[0, :, 0] * B).sum() + (w.sum(2)[0, :, 0] * C).sum()
        g_sig, g_col = torch.autograd.grad(loss, (sigma, col))
        rng = torch.stack([depths.min(), depths.max()])
        gs = torch.empty(r, s, device=d); om = torch.empty(r, s, device=d)
        P = lambda t: ctypes.c_void_p(t.data_ptr())
        pkg._lib.check(pkg._lib.lib().tpr_march_backward(P(coarse), P(fine), dc, df, P(sigma.detach().contiguous()),
                                                         P(col.detach().contiguous()), P(A), P(B), P(C), P(rng), int(white), r,
                                                         P(gs), P(om), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       'tpr_march_backward')
        torch.cuda.synchronize()
        assert rel_err(gs.cpu().numpy(), g_sig.cpu().numpy()) < 2e-5
        want_col = g_col.cpu().numpy()
        got_col = (2 * A.unsqueeze(1) * om.unsqueeze(-1)).cpu().numpy()
        assert rel_err(got_col, want_col) < 2e-5


@pytest.mark.parametrize('shape', [(1, 24, 64, 48, 48), (2, 16, 48, 96, 96), (1, 20, 40, 32, 0)])
def test_gradients_match_autograd_through_the_torch_oracle(pkg, shape):
    """Sizes the fixtures do not cover (a few thousand rays, 96+96 depths), against same-device autograd."""
    n, res, pres, dc, df = shape
    scene = O.synthetic_scene(31 + dc, n, res, pres, dc, df, 0.5)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df)
    rng = np.random.RandomState(77)
    m = res * res
    A, B, C = (rng.standard_normal((n, m, k)).astype(np.float32) for k in (32, 1, 1))
    (rgb, depth, wsum), grads = run_backward(pkg, scene, opts, A, B, C)
    (rgb_o, depth_o, wsum_o), grads_o = TO.render_grads(T(scene['planes']), TO.decoder_tuple(scene['dec'], dev()),
                                                        T(scene['origins']), T(scene['dirs']), opts, T(scene['jitter']),
                                                        T(scene['u']), T(A), T(B), T(C))
    assert (rgb.detach() - rgb_o).abs().max() < 1e-4
    errs = {k: rel_err(g.cpu().numpy(), go.cpu().numpy()) for k, g, go in zip(('planes', 'w1', 'b1', 'w2', 'b2'), grads, grads_o)}
    print(shape, {k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < REL, errs


def test_backward_is_linear_in_the_upstream_gradient_and_respects_requires_grad(pkg):
    scene, opts, gold, (A, B, C) = load_bwd_case('bwd_ffhq')
    _, g1 = run_backward(pkg, scene, opts, A, B, C)
    _, g2 = run_backward(pkg, scene, opts, 2 * A, 2 * B, 2 * C)
    for a, b in zip(g1, g2):
        assert rel_err((2 * a).cpu().numpy(), b.cpu().numpy()) < 1e-5
    # gradients come at any scale (loss scaling): the backward rescales its fp16 gradient-side operands by a power of two
    # taken from the upstream gradient's magnitude (csrc/tpr_backward_tc.cu: scale_kernel), so 2^-20 x and 2^20 x the upstream
    # gradient give exactly-scaled results, neither flushed to zero nor saturated
    for k in (-20, 20):
        f = float(2.0 ** k)
        _, gk = run_backward(pkg, scene, opts, f * A, f * B, f * C)
        for a, b in zip(g1, gk):
            assert rel_err((f * a).cpu().numpy(), b.cpu().numpy()) < 1e-5, k
    # planes only: the decoder stays frozen and gets no .grad
    dec = make_decoder(pkg, scene['dec'])
    planes = T(scene['planes']).requires_grad_(True)
    rgb, depth, wsum = pkg.ImportanceRenderer()(planes, dec, T(scene['origins']), T(scene['dirs']), opts,
                                                noise=(T(scene['jitter']), T(scene['u'])))
    (rgb * T(A)).sum().backward()
    assert planes.grad is not None and dec.net[0].weight.grad is None
    # rays that require grad are refused, not silently ignored
    with pytest.raises(NotImplementedError):
        pkg.ImportanceRenderer()(planes, dec, T(scene['origins']).requires_grad_(True), T(scene['dirs']), opts)
    # under no_grad the inference path runs as before
    with torch.no_grad():
        out = pkg.ImportanceRenderer()(planes, dec, T(scene['origins']), T(scene['dirs']), opts,
                                       noise=(T(scene['jitter']), T(scene['u'])))
    assert not out[0].requires_grad
    torch.testing.assert_close(out[0], rgb.detach(), rtol=0, atol=0)


@pytest.mark.parametrize('name', ['bwd_ffhq', 'bwd_white'])
def test_reduced_precision_mode_gradients(pkg, name):
    """decoder_precision='bf16' (the caller's >= 50 dB mode): bf16 decoder operands in the forward / point query, one
    TF32 HMMA per product (round-to-nearest operands) in the backward.  No gate is stated for gradients in that mode;
    the test pins the measured accuracy class: 1 % of the largest entry."""
    scene, opts, gold, (A, B, C) = load_bwd_case(name)
    _, grads = run_backward(pkg, scene, dict(opts, decoder_precision='bf16'), A, B, C)
    errs = {k: rel_err(g.cpu().numpy(), gold[k]) for g, k in zip(grads, ('g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2'))}
    print(name, 'bf16 mode', {k: f'{v:.1e}' for k, v in errs.items()})
    assert max(errs.values()) < 1e-2, errs


@pytest.mark.parametrize('name', list(BWD_CASES))
def test_gradients_without_kept_samples(pkg, name):
    """keep_samples=False: the forward keeps nothing per sample and the backward re-evaluates colours / sigma with the
    point-query kernel -- same gate, and the two paths agree closely."""
    scene, opts, gold, (A, B, C) = load_bwd_case(name)
    _, g_keep = run_backward(pkg, scene, opts, A, B, C, keep_samples=True)
    _, g_eval = run_backward(pkg, scene, opts, A, B, C, keep_samples=False)
    _, g_half = run_backward(pkg, scene, opts, A, B, C, keep_samples=True, keep_features=False)    # colours kept, features re-gathered
    for grads in (g_eval, g_half):
        for g, k in zip(grads, ('g_planes', 'g_w1', 'g_b1', 'g_w2', 'g_b2')):
            assert rel_err(g.cpu().numpy(), gold[k]) < REL, k
        for a, b in zip(g_keep, grads):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 5e-5


def test_only_the_requested_gradients_are_computed(pkg):
    """Frozen decoder / frozen planes: the corresponding phase of the backward kernel is skipped (NULL output in the C ABI)
    and the remaining gradient is unchanged."""
    scene, opts, gold, (A, B, C) = load_bwd_case('bwd_ffhq')
    _, full = run_backward(pkg, scene, opts, A, B, C)
    args = (T(scene['origins']), T(scene['dirs']), opts)
    noise = (T(scene['jitter']), T(scene['u']))
    loss = lambda out: (out[0] * T(A)).sum() + (out[1] * T(B)).sum() + (out[2] * T(C)).sum()
    planes = T(scene['planes']).requires_grad_(True)
    loss(pkg.ImportanceRenderer()(planes, make_decoder(pkg, scene['dec']), *args, noise=noise)).backward()
    assert rel_err(planes.grad.cpu().numpy(), full[0].cpu().numpy()) < 1e-5
    dec = make_decoder(pkg, scene['dec']).requires_grad_(True)
    loss(pkg.ImportanceRenderer()(T(scene['planes']), dec, *args, noise=noise)).backward()
    for got, want in zip((dec.net[0].weight.grad, dec.net[0].bias.grad, dec.net[2].weight.grad, dec.net[2].bias.grad), full[1:]):
        assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < 1e-5


def test_gradients_with_disparity_sampling_and_auto_limits(pkg):
    """The other two coarse-depth branches (VR/renderer.py:174-186): disparity-space sampling against same-device autograd
    through the torch oracle; per-ray 'auto' limits (no oracle for that branch): the two backward paths (kept samples /
    re-evaluated samples) must agree and be finite."""
    n, res, pres, dc, df = 1, 12, 40, 32, 32
    scene = O.synthetic_scene(57, n, res, pres, dc, df, 0.5)
    rng = np.random.RandomState(78)
    A, B, C = (rng.standard_normal((n, res * res, k)).astype(np.float32) for k in (32, 1, 1))
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, disparity_space_sampling=True)
    _, grads = run_backward(pkg, scene, opts, A, B, C)
    _, grads_o = TO.render_grads(T(scene['planes']), TO.decoder_tuple(scene['dec'], dev()), T(scene['origins']), T(scene['dirs']),
                                 opts, T(scene['jitter']), T(scene['u']), T(A), T(B), T(C))
    for g, go in zip(grads, grads_o):
        assert rel_err(g.cpu().numpy(), go.cpu().numpy()) < REL
    auto = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, ray_start='auto', ray_end='auto')
    _, g1 = run_backward(pkg, scene, auto, A, B, C, keep_samples=True)
    _, g2 = run_backward(pkg, scene, auto, A, B, C, keep_samples=False)
    for a, b in zip(g1, g2):
        assert torch.isfinite(a).all() and a.abs().max() > 0
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 5e-5
