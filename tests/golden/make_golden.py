"""Generate the committed known-answer fixtures by running the UNMODIFIED
reference (``/root/reference/g_nerf``, imported, CPU) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by ``oracle.triplane_oracle.synthetic_scene``
(numpy's frozen RandomState stream), so only the reference's OUTPUTS are stored.
The reference's two random draws (torch.rand_like at VR/renderer.py:190 and
torch.rand at :237) are replaced, for the duration of the call, by the seeded
``jitter`` / ``u`` tensors -- the reference code itself is not modified.
"""
import os
import sys
import contextlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/g_nerf')

from oracle import triplane_oracle as O                      # noqa: E402
from training.volumetric_rendering.renderer import ImportanceRenderer   # noqa: E402
from training.volumetric_rendering.ray_sampler import RaySampler         # noqa: E402
from training.volumetric_rendering.ray_marcher import MipRayMarcher2     # noqa: E402
from training.triplane import OSGDecoder                                  # noqa: E402


@contextlib.contextmanager
def injected_noise(jitter, u, normal_draws=()):
    """Make the next torch.rand_like / torch.rand return the seeded draws; successive torch.randn_like calls (the
    density_noise term, VR/renderer.py:146) return ``normal_draws`` in order."""
    real_rand_like, real_rand, real_randn_like = torch.rand_like, torch.rand, torch.randn_like
    pending = list(normal_draws)
    torch.rand_like = lambda t, *a, **k: torch.from_numpy(jitter).reshape(t.shape).clone()
    torch.rand = lambda *a, **k: torch.from_numpy(u).clone()
    torch.randn_like = lambda t, *a, **k: torch.from_numpy(pending.pop(0)).reshape(t.shape).clone()
    try:
        yield
    finally:
        torch.rand_like, torch.rand, torch.randn_like = real_rand_like, real_rand, real_randn_like


def ref_decoder(dec: O.DecoderParams) -> OSGDecoder:
    m = OSGDecoder(32, {'decoder_lr_mul': dec.lr_mul, 'decoder_output_dim': 32})
    with torch.no_grad():
        m.net[0].weight.copy_(torch.from_numpy(dec.w1))
        m.net[0].bias.copy_(torch.from_numpy(dec.b1))
        m.net[2].weight.copy_(torch.from_numpy(dec.w2))
        m.net[2].bias.copy_(torch.from_numpy(dec.b2))
    return m.requires_grad_(False)


from tests.cases import CASES, density_noise_draws                      # noqa: E402


def run_case(name):
    seed, n, res, pres, dc, df, bs, extra = CASES[name]
    sc = O.synthetic_scene(seed, n, res, pres, dc, df, bs)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, **extra)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    dec = ref_decoder(sc['dec'])
    # --- ray sampler
    ro, rd = RaySampler()(t(sc['c2w']), t(sc['K']), res)
    # --- forward, capturing stage tensors through the public sub-calls
    R = ImportanceRenderer()
    cap = {}
    real_si, real_sp = R.sample_importance, R.sample_pdf

    def si(z, w, k):
        cap['weights_coarse'] = w.numpy().copy()
        out = real_si(z, w, k)
        cap['depths_fine'] = out.numpy().copy()
        return out

    def sp(bins, weights, k, **kw):
        cap['pdf_bins'], cap['pdf_weights'] = bins.numpy().copy(), weights.numpy().copy()
        return real_sp(bins, weights, k, **kw)
    R.sample_importance, R.sample_pdf = si, sp
    real_ss = torch.searchsorted

    def ss(cdf, uu, **kw):
        out = real_ss(cdf, uu, **kw)
        cap['inds'] = out.numpy().copy()
        return out
    torch.searchsorted = ss
    noisy = opts.get('density_noise', 0) > 0
    nz_c, nz_f, nz_p = density_noise_draws(name) if noisy else (None, None, None)
    try:
        with injected_noise(sc['jitter'], sc['u'], [nz_c, nz_f] if noisy else []):
            rgb, depth, wsum = R(t(sc['planes']), dec, ro, rd, opts)
    finally:
        torch.searchsorted = real_ss
    # --- run_model on scattered points (some outside the box)
    rng = np.random.RandomState(seed + 1000)
    pts = (rng.random_sample((n, 257, 3)).astype(np.float32) - 0.5) * 1.3 * opts['box_warp']
    with injected_noise(sc['jitter'], sc['u'], [nz_p] if noisy else []):
        rm = ImportanceRenderer().run_model(t(sc['planes']), dec, t(pts), None, opts)
    out = dict(origins=ro.numpy(), dirs=rd.numpy(), rgb=rgb.numpy(), depth=depth.numpy(),
               wsum=wsum.numpy(), pts=pts, pts_rgb=rm['rgb'].numpy(), pts_sigma=rm['sigma'].numpy())
    out.update(cap)
    if opts['ray_start'] == 'auto':
        # a14: the box limits themselves (VR/math_utils.py:46-98), plus a second set of rays aimed at and past a unit box
        from training.volumetric_rendering import math_utils as MU
        tmin, tmax = MU.get_ray_limits_box(ro, rd, box_side_length=opts['box_warp'])
        brng = np.random.RandomState(seed + 2000)
        bo = (brng.standard_normal((3, 97, 3)) * 1.2).astype(np.float32)
        bd = brng.standard_normal((3, 97, 3)).astype(np.float32)
        bd /= np.linalg.norm(bd, axis=-1, keepdims=True)
        bmin, bmax = MU.get_ray_limits_box(t(bo), t(bd), box_side_length=1.0)
        out.update(box_tmin=tmin.numpy(), box_tmax=tmax.numpy(), box2_origins=bo, box2_dirs=bd, box2_tmin=bmin.numpy(),
                   box2_tmax=bmax.numpy())
    # --- stand-alone marcher on random inputs
    s = 17
    mc, ms = rng.random_sample((1, 33, s, 32)).astype(np.float32), rng.standard_normal((1, 33, s, 1)).astype(np.float32) * 3
    md = np.sort(rng.random_sample((1, 33, s, 1)).astype(np.float32) + 2, 2)
    mr = MipRayMarcher2()(t(mc), t(ms), t(md), opts)
    out.update(march_colors=mc, march_sigma=ms, march_depths=md,
               march_rgb=mr[0].numpy(), march_depth=mr[1].numpy(), march_w=mr[2].numpy())
    return sc, opts, out


def main():
    for name in (sys.argv[1:] or CASES):      # optional: only the named cases
        sc, opts, out = run_case(name)
        # cross-check the numpy oracle against the reference before committing
        draws = density_noise_draws(name)[:2] if opts.get('density_noise', 0) > 0 else None
        (rgb, depth, wsum), st = O.render(sc['planes'], sc['dec'], sc['origins'], sc['dirs'], opts,
                                          sc['jitter'], sc['u'], return_stages=True, density_noise_draws=draws)
        err = {k: float(np.abs(a - out[k]).max()) for k, a in
               dict(rgb=rgb, depth=depth, wsum=wsum, origins=sc['origins'], dirs=sc['dirs']).items()}
        if 'inds' in out:
            err['inds_mismatch'] = int((st['inds'] != out['inds']).sum())
            err['fine'] = float(np.abs(st['depths_fine'] - out['depths_fine']).max())
        print(name, {k: (f'{v:.2e}' if isinstance(v, float) else v) for k, v in err.items()})
        np.savez_compressed(os.path.join(HERE, f'{name}.npz'),
                            **{k: v for k, v in out.items()})
        print('  wrote', name, os.path.getsize(os.path.join(HERE, f'{name}.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    torch.set_grad_enabled(False)
    main()
