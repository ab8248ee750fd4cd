"""GPU: BASELINE.json's full config-2 shape, EVERY ray: the whole 8-image batch against the UNMODIFIED reference on the same
GPU from the same generator state (oracle/_ref), and an image pair with injected draws against the torch restatement
(oracle/torch_oracle.py, pinned on CPU against the reference fixtures by tests/test_torch_oracle.py), which also exposes the
intermediate stages (importance depths, searchsorted indices).

Gates (BASELINE.json north_star): fp32 mode max-abs <= 1e-4 on rgb / depth / weight sum; bf16-MLP mode PSNR >= 50 dB;
sample_pdf bin indices bit-exact given identical weights / bins / u."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import triplane_oracle as O
from oracle import torch_oracle as TO
from tests.test_gpu_parity import T, make_decoder, TOL, dev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def full_scene():
    torch.backends.cuda.matmul.allow_tf32 = False        # as the reference sets it (training_loop.py:145-146)
    torch.backends.cudnn.allow_tf32 = False
    scene = O.synthetic_scene(21, 2, 128, 256, 48, 48, 0.5)
    return scene, dict(O.FFHQ_OPTIONS)


def _psnr(a, b, peak):
    return 10 * np.log10(peak * peak / max(float(((a - b) ** 2).mean()), 1e-30))


def test_config2_every_ray_fp32_and_bf16(pkg, full_scene):
    scene, opts = full_scene
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    planes, o, d = T(scene['planes']), T(scene['origins']), T(scene['dirs'])
    jitter, u = T(scene['jitter']), T(scene['u'])
    with torch.no_grad():
        want, st = TO.render(planes, TO.decoder_tuple(scene['dec'], dev()), o, d, opts, jitter, u, return_stages=True)
    R.debug_outputs = True
    got = R(planes, dec, o, d, opts, noise=(jitter, u))
    names = ('rgb', 'depth', 'wsum')
    errs = {k: float((g - w).abs().max()) for k, g, w in zip(names, got, want)}
    print('config-2 image pair, 32768 rays, fp32 max-abs vs torch-on-GPU:', errs)
    assert max(errs.values()) < TOL, errs
    # importance depths / indices of the fused kernel against the restatement's (their coarse weights differ in the last
    # bits, so a draw that lands within an ulp of a CDF entry may legitimately fall in the neighbouring bin)
    fine_d, fine_i = R.last_fine
    flips = int((fine_i.long() != st['inds']).sum())
    print('end-to-end searchsorted index flips:', flips, 'of', fine_i.numel())
    assert flips <= 1e-4 * fine_i.numel()
    same = fine_i.long() == st['inds']
    assert float((fine_d - st['depths_fine'].reshape(fine_d.shape)).abs()[same].max()) < 1e-5
    # bf16-MLP mode
    got16 = R(planes, dec, o, d, dict(opts, decoder_precision='bf16'), noise=(jitter, u))
    psnr = {'rgb': _psnr(got16[0].cpu().numpy(), want[0].cpu().numpy(), 2.0),
            'depth': _psnr(got16[1].cpu().numpy(), want[1].cpu().numpy(), opts['ray_end'] - opts['ray_start'])}
    print('bf16-MLP PSNR vs torch-on-GPU (dB):', psnr)
    assert min(psnr.values()) >= 50.0, psnr


@pytest.mark.skipif(ref_loader.reference_dir() is None, reason='oracle/_ref (python oracle/build_ref.py) not present')
def test_config2_full_batch_against_the_unmodified_reference(pkg):
    """BASELINE configs[1] as the bench runs it -- 8 images x 128^2 rays x (48+48), 3x32x256^2 planes -- through the
    reference's own ImportanceRenderer.forward (VR/renderer.py:88-140) and ours, both drawing their noise from the same CUDA
    generator state.  1e-4 on every ray of every image; bf16-MLP mode >= 50 dB; the generator ends in the same state."""
    from tests.test_gpu_reference import _both, _decoders, _psnr as psnr_t
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref = ref_loader.import_reference()
    g = torch.Generator().manual_seed(5)
    planes = torch.randn((8, 3, 32, 256, 256), generator=g).to(dev())
    scene = O.synthetic_scene(22, 8, 128, 8, 2, 2, 0.5)                      # (only its cameras / rays are used)
    o, d = T(scene['origins']), T(scene['dirs'])
    theirs, ours = _decoders(pkg, ref)
    opts = dict(O.FFHQ_OPTIONS)
    want, got, s_ref, s_ours = _both(pkg, ref, planes, theirs, ours, o, d, opts)
    errs = {k: float((a - b).abs().max()) for k, a, b in zip(('rgb', 'depth', 'wsum'), got, want)}
    print('config 2, 8 images x 16384 rays, fp32 max-abs vs the unmodified reference on the same GPU:', errs)
    assert max(errs.values()) < TOL, errs
    assert torch.equal(s_ref, s_ours)
    torch.manual_seed(1)
    with torch.no_grad():
        got16 = pkg.ImportanceRenderer()(planes, ours, o, d, dict(opts, decoder_precision='bf16'))
    db = {'rgb': psnr_t(got16[0], want[0], 2.0), 'depth': psnr_t(got16[1], want[1], opts['ray_end'] - opts['ray_start'])}
    print('bf16-MLP PSNR vs the reference (dB):', db)
    assert min(db.values()) >= 50.0, db


def test_config2_sample_importance_indices_bit_exact(pkg, full_scene):
    """Given IDENTICAL coarse depths, weights and u, the stand-alone resampling kernel returns exactly the oracle's
    searchsorted indices and samples on every one of the 32768 x 48 draws (the oracle's arithmetic contract: row sum and
    running sum accumulated exactly, rounded to float32 per entry -- oracle/triplane_oracle.py:pdf_to_cdf).  torch itself is
    not one target: its CPU sum is a vectorised float32 cascade (machine dependent), its CUDA cumsum a float32 tree, so the
    two disagree with each other on draws that land within an ulp of a CDF entry; both are counted and bounded."""
    scene, opts = full_scene
    planes, o, d = T(scene['planes']), T(scene['origins']), T(scene['dirs'])
    jitter, u = T(scene['jitter']), T(scene['u'])
    with torch.no_grad():
        _, st = TO.render(planes, TO.decoder_tuple(scene['dec'], dev()), o, d, opts, jitter, u, return_stages=True)
        d_c = TO.coarse_depths(jitter, opts['ray_start'], opts['ray_end'], False)
        w_c = st['weights_coarse']
        fine_cpu, inds_cpu = TO.importance_depths(d_c.cpu(), w_c.cpu(), u.cpu())
        fine_gpu, inds_gpu = TO.importance_depths(d_c, w_c, u)
    got, inds = pkg.ImportanceRenderer().sample_importance(d_c, w_c, 48, u=u, return_inds=True)
    want, want_inds = O.sample_importance(d_c.cpu().numpy(), w_c.cpu().numpy(), scene['u'])
    np.testing.assert_array_equal(inds.cpu().numpy(), want_inds)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    n = inds.numel()
    flips = {'kernel_vs_torch_cpu': int((inds.long().cpu() != inds_cpu).sum()),
             'kernel_vs_torch_cuda': int((inds.long() != inds_gpu).sum()),
             'torch_cpu_vs_torch_cuda': int((inds_cpu != inds_gpu.cpu()).sum())}
    print('searchsorted index differences out of', n, ':', flips)
    assert max(flips.values()) <= 1e-4 * n, flips


def test_config5_density_grid_slab(pkg, full_scene):
    """run_model on a slab of the 256^3 density grid of gen_videos.py:33-55,198-209 (config 5): sigma and rgb of every
    point against the restatement."""
    scene, opts = full_scene
    g = 256
    ax = (torch.arange(g, device=dev(), dtype=torch.float32) + 0.5) / g - 0.5
    zz, yy, xx = torch.meshgrid(ax[100:108], ax, ax, indexing='ij')
    xyz = torch.stack([xx, yy, zz], -1).reshape(1, -1, 3).contiguous()
    planes = T(scene['planes'][:1])
    with torch.no_grad():
        rgb_w, sig_w = TO.decode(TO.gather(planes, xyz, opts['box_warp']), TO.decoder_tuple(scene['dec'], dev()))
    out = pkg.ImportanceRenderer().run_model(planes, make_decoder(pkg, scene['dec']), xyz, None, opts)
    assert float((out['sigma'] - sig_w).abs().max()) < TOL
    assert float((out['rgb'] - rgb_w).abs().max()) < TOL
