"""CPU: the C/OpenMP oracle (bench.py's all-threads CPU baseline) against the reference fixtures and the
numpy oracle."""
import numpy as np
import pytest

from tests.cases import SCALAR_CASES as CASES, load_case
from oracle import triplane_oracle as O
from oracle import c_oracle

pytestmark = pytest.mark.skipif(not c_oracle.available(), reason='gcc not available to build the C oracle')


@pytest.mark.parametrize('name', list(CASES))
def test_c_oracle_matches_reference_fixture_and_numpy_oracle(name):
    scene, opts, gold = load_case(name)
    rgb, depth, wsum, st = c_oracle.render(scene, opts, want_stages=True)
    rgb_o, depth_o, wsum_o = O.render(scene['planes'], scene['dec'], scene['origins'], scene['dirs'], opts,
                                      scene['jitter'], scene['u'])
    for got, a, b in ((rgb, rgb_o, gold['rgb']), (depth, depth_o, gold['depth']), (wsum, wsum_o, gold['wsum'])):
        assert np.abs(got - a).max() < 1e-5
        assert np.abs(got - b).max() < 1e-5
    if opts['depth_resolution_importance'] > 0:
        assert np.abs(st['depths_fine'] - gold['depths_fine']).max() < 1e-5
        assert int((st['inds'] != gold['inds']).sum()) <= max(1, 1e-4 * gold['inds'].size)


def test_c_oracle_uses_threads():
    assert c_oracle.num_threads() >= 1
