"""ctypes wrapper of the C oracle (TEST INFRASTRUCTURE ONLY; see triplane_oracle.c)."""
import ctypes
import os
import shutil

import numpy as np

from . import build_c

_lib = None


def _load():
    global _lib
    if _lib is None:
        path = build_c.LIB
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(build_c.SRC):
            if shutil.which('gcc') is None:
                raise RuntimeError('C oracle not built and gcc is not available')
            path = build_c.build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_num_threads.restype = ctypes.c_int
        _lib.oracle_render.restype = ctypes.c_int
    return _lib


def available() -> bool:
    try:
        _load()
        return True
    except Exception:
        return False


def num_threads() -> int:
    return int(_load().oracle_num_threads())


def use_all_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline is supposed to use every host core it may."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    _load().oracle_set_num_threads(int(n))
    return num_threads()


def render(scene: dict, opts: dict, want_stages: bool = False):
    """scene as produced by triplane_oracle.synthetic_scene; returns (rgb, depth, wsum[, stages])."""
    lib = _load()
    planes = np.ascontiguousarray(scene['planes'], np.float32)
    n, _, _, h, w = planes.shape
    w1, b1, w2, b2 = (np.ascontiguousarray(a, np.float32) for a in scene['dec'].effective())
    o = np.ascontiguousarray(scene['origins'], np.float32)
    d = np.ascontiguousarray(scene['dirs'], np.float32)
    m = o.shape[1]
    dc, df = int(opts['depth_resolution']), int(opts['depth_resolution_importance'])
    jit = np.ascontiguousarray(scene['jitter'], np.float32)
    u = np.ascontiguousarray(scene['u'], np.float32)
    rgb = np.empty((n, m, 32), np.float32)
    depth = np.empty((n, m, 1), np.float32)
    wsum = np.empty((n, m, 1), np.float32)
    fine = np.empty((n * m, max(df, 1)), np.float32)
    inds = np.empty((n * m, max(df, 1)), np.int32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.oracle_render(P(planes), n, h, w, P(w1), P(b1), P(w2), P(b2), P(o), P(d), m, P(jit), P(u),
                           ctypes.c_double(opts['ray_start']), ctypes.c_double(opts['ray_end']),
                           ctypes.c_double(opts['box_warp']), dc, df,
                           int(bool(opts.get('disparity_space_sampling', False))), int(bool(opts.get('white_back', False))),
                           P(rgb), P(depth), P(wsum), P(fine) if want_stages else None, P(inds) if want_stages else None)
    if rc != 0:
        raise RuntimeError(f'oracle_render failed: {rc}')
    if want_stages:
        return rgb, depth, wsum, {'depths_fine': fine.reshape(n, m, -1, 1), 'inds': inds}
    return rgb, depth, wsum
