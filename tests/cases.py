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
}


def load_case(name):
    seed, n, res, pres, dc, df, bs, extra = CASES[name]
    scene = O.synthetic_scene(seed, n, res, pres, dc, df, bs)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, **extra)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')))
    return scene, opts, gold
