"""Golden cases shared by the CPU (oracle vs reference fixtures) and GPU (kernels vs oracle) tests.
Must stay in sync with tests/golden/make_golden.py:CASES (the fixtures were generated from it)."""
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
}


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
