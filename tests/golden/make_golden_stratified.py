"""Known answers for sample_stratified (VR/renderer.py:169-192) from the UNMODIFIED reference, all three branches
(scalar limits, disparity-space, per-ray tensor limits), with the torch.rand_like draw replaced by a seeded jitter.
Build container only:   python tests/golden/make_golden_stratified.py   -> tests/golden/stratified.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, '/root/reference/g_nerf')
from oracle import triplane_oracle as O                                            # noqa: E402
from training.volumetric_rendering.renderer import ImportanceRenderer              # noqa: E402

from tests.stratified_cases import CASES, inputs                                   # noqa: E402


def main():
    out = {}
    real = torch.rand_like
    for name, (n, m, d, rs, re, disp) in CASES.items():
        jitter, lim = inputs(name)
        torch.rand_like = lambda t, *a, **k: torch.from_numpy(jitter).reshape(t.shape).clone()
        try:
            o = torch.zeros(n, m, 3)
            if lim is None:
                got = ImportanceRenderer().sample_stratified(o, rs, re, d, disp)
                want = O.stratified_depths(jitter, rs, re, disp)
            else:
                got = ImportanceRenderer().sample_stratified(o, torch.from_numpy(lim[0]), torch.from_numpy(lim[1]), d, disp)
                want = O.stratified_depths_per_ray(jitter, *lim)
        finally:
            torch.rand_like = real
        out[name] = got.numpy()
        print(name, 'oracle == reference:', bool((want == out[name]).all()), float(np.abs(want - out[name]).max()))
    np.savez_compressed(os.path.join(HERE, 'stratified.npz'), **out)


if __name__ == '__main__':
    torch.set_grad_enabled(False)
    main()
