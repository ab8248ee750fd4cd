"""Seeded inputs of the sample_stratified fixtures, shared by tests/golden/make_golden_stratified.py (which ran the
reference on them) and the tests (which run the oracle and the kernels on them)."""
import numpy as np

# name: (n_img, n_rays, depth_resolution, ray_start, ray_end, disparity_space_sampling); None limits = per-ray tensors
CASES = {'scalar_48': (2, 37, 48, 2.25, 3.3, False), 'scalar_7': (1, 5, 7, 0.1, 9.0, False),
         'disparity_32': (1, 29, 32, 2.25, 3.3, True), 'disparity_96': (1, 3, 96, 0.5, 4.0, True),
         'per_ray_12': (2, 31, 12, None, None, False), 'per_ray_96': (1, 9, 96, None, None, False)}


def inputs(name):
    n, m, d, rs, re, disp = CASES[name]
    g = np.random.RandomState(sum(map(ord, name)))
    jitter = g.random_sample((n, m, d, 1)).astype(np.float32)
    lim = None
    if rs is None:
        a = g.uniform(1.5, 2.5, (n, m, 1)).astype(np.float32)
        lim = (a, a + g.uniform(0.5, 1.5, (n, m, 1)).astype(np.float32))
    return jitter, lim
