"""Golden cases shared by the CPU (oracle vs reference fixtures) and GPU (kernels vs oracle) tests and by
tests/golden/make_golden.py, which generated the fixtures from them."""
import os

import numpy as np

from oracle import triplane_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = {
    # name: (seed, n_img, res, plane_res, dc, df, bias_scale, extra options)
    'ffhq_small':  (11, 2, 16, 64, 48, 48, 0.5, {}),
    'white_back':  (12, 1, 12, 32, 24, 24, 0.5, {'white_back': True}),
    'ragged':      (13, 1, 9, 48, 20, 13, 0.5, {}),
    'disparity':   (14, 1, 8, 32, 32, 32, 0.0, {'disparity_space_sampling': True}),
    'coarse_only': (15, 1, 8, 32, 32, 0, 0.5, {}),
    'wide_box':    (16, 1, 10, 40, 16, 16, 0.5, {'box_warp': 0.6}),
    # R = 4 ray groups of the warp-specialised kernel (more than 64 samples per pass): the gen_videos.py:127-128 depths,
    # and two shapes that exercise the pair rank count with 2 / 3 purely-coarse rows
    'inference_96': (17, 1, 12, 48, 96, 96, 0.5, {}),
    'mid_64':       (18, 1, 10, 40, 64, 64, 0.5, {}),
    'uneven_96_40': (19, 1, 8, 32, 96, 40, 0.5, {}),
    # the 'auto' ray limits (VR/renderer.py:91-97 + math_utils.get_ray_limits_box): a box small enough that the outer rays miss it
    'auto_limits':  (24, 2, 12, 48, 32, 32, 0.5, {'ray_start': 'auto', 'ray_end': 'auto', 'box_warp': 0.5}),
    # density_noise (VR/renderer.py:146): sigma += randn_like(sigma) * density_noise after each point query
    'density_noise': (25, 1, 10, 40, 24, 24, 0.5, {'density_noise': 0.5}),
}


def density_noise_draws(name):
    """The standard-normal draws that stand for torch.randn_like at VR/renderer.py:146 in a density_noise case:
    (coarse [N,M*Dc,1], fine [N,M*Df,1], run_model points [N,257,1])."""
    seed, n, res, _, dc, df, _, _ = CASES[name]
    rng = np.random.RandomState(seed + 3000)
    m = res * res
    return (rng.standard_normal((n, m * dc, 1)).astype(np.float32), rng.standard_normal((n, m * df, 1)).astype(np.float32),
            rng.standard_normal((n, 257, 1)).astype(np.float32))


# the cases the scalar-limits-only restatements (oracle/torch_oracle.py, oracle/triplane_oracle.c) cover
SCALAR_CASES = [n for n, c in CASES.items() if not isinstance(c[7].get('ray_start', 0.0), str) and not c[7].get('density_noise', 0)]


def case_noise(name, scene):
    """The random draws of a case as the tuple ImportanceRenderer.forward(noise=...) takes (numpy arrays)."""
    if CASES[name][7].get('density_noise', 0) > 0:
        nz_c, nz_f, _ = density_noise_draws(name)
        return scene['jitter'], scene['u'], nz_c, nz_f
    return scene['jitter'], scene['u']


def oracle_render(name, scene, opts, origins=None, dirs=None, return_stages=True):
    draws = density_noise_draws(name)[:2] if opts.get('density_noise', 0) > 0 else None
    return O.render(scene['planes'], scene['dec'], scene['origins'] if origins is None else origins,
                    scene['dirs'] if dirs is None else dirs, opts, scene['jitter'], scene['u'], return_stages=return_stages,
                    density_noise_draws=draws)


def coarse_depths(scene, opts):
    """The case's coarse depths [N,M,Dc,1] (VR/renderer.py:169-192), whichever branch its options select."""
    if isinstance(opts['ray_start'], str):
        return O.stratified_depths_per_ray(scene['jitter'], *O.auto_ray_limits(scene['origins'], scene['dirs'], opts['box_warp']))
    return O.stratified_depths(scene['jitter'], opts['ray_start'], opts['ray_end'], opts.get('disparity_space_sampling', False))


def depth_peak(opts, gold):
    """Peak value for a depth PSNR: ray_end - ray_start, or the spread of the fixture's depths with 'auto' limits."""
    if isinstance(opts['ray_start'], str):
        return float(gold['depth'].max() - gold['depth'].min())
    return opts['ray_end'] - opts['ray_start']


def load_case(name):
    seed, n, res, pres, dc, df, bs, extra = CASES[name]
    scene = O.synthetic_scene(seed, n, res, pres, dc, df, bs)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, **extra)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')))
    return scene, opts, gold


# Gradient fixtures; must stay in sync with tests/golden/make_golden_backward.py (CASES, loss_weights).
BWD_CASES = {
    'bwd_ffhq':   (21, 2, 8, 20, 48, 48, 0.5, {}),
    'bwd_white':  (22, 1, 7, 24, 20, 13, 0.5, {'white_back': True}),
    'bwd_coarse': (23, 1, 6, 16, 32, 0, 0.5, {'box_warp': 0.6}),
}


def loss_weights(seed, n, m):
    """Upstream gradients (A for rgb, B for depth, C for weight_sum) of L = sum(rgb*A) + sum(depth*B) + sum(wsum*C)."""
    rng = np.random.RandomState(seed + 2000)
    return (rng.standard_normal((n, m, 32)).astype(np.float32), rng.standard_normal((n, m, 1)).astype(np.float32),
            rng.standard_normal((n, m, 1)).astype(np.float32))


def load_bwd_case(name):
    seed, n, res, pres, dc, df, bs, extra = BWD_CASES[name]
    scene = O.synthetic_scene(seed, n, res, pres, dc, df, bs)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, **extra)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')))
    return scene, opts, gold, loss_weights(seed, n, res * res)
