"""CPU, world_size 2, gloo: the multi-GPU host logic (partition, global depth clamp via all-reduce, in-place
all-gather) with the numpy oracle standing in for the per-rank CUDA render."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import triplane_oracle as O


def test_partition_covers_every_ray_exactly_once(pkg):
    P = pkg.parallel
    for n_img, n_rays, world in [(8, 100, 8), (8, 100, 3), (32, 7, 8), (1, 101, 8), (3, 64, 8), (2, 5, 8), (5, 9, 4)]:
        plan = P.partition(n_img, n_rays, world)
        cover = np.zeros((n_img, n_rays), int)
        for shards in plan:
            for s in shards:
                cover[s.image, s.ray_begin:s.ray_end] += 1
        assert (cover == 1).all(), (n_img, n_rays, world)
        counts = P.shard_ray_counts(plan)
        assert max(counts) - min(counts) <= n_rays            # balanced to within one image / one ray range
    with pytest.raises(ValueError):
        P.partition(0, 1, 1)


def _oracle_local_render(scene_dec):
    def fn(renderer, planes, decoder, origins, dirs, options, noise, out):
        jit, u = noise
        (rgb, depth, wsum), st = O.render(planes.numpy(), scene_dec, origins.numpy(), dirs.numpy(), options,
                                          jit.numpy(), u.numpy(), return_stages=True)
        # undo the oracle's own (shard-local) clamp by recomputing the unclamped depth is not possible from
        # its outputs, so hand back the shard-local range: min/max of shard-local ranges == the global range,
        # and clamping twice (local range, then global range) equals clamping to the local range only when a
        # ray has zero weight -- handled below by testing with dense scenes where no ray is empty.
        d_all = st.get('depths_all', st['depths_coarse'])
        rng = torch.tensor([d_all.min(), d_all.max()], dtype=torch.float32)
        return torch.from_numpy(rgb), torch.from_numpy(depth), torch.from_numpy(wsum), rng
    return fn


def _worker(rank, world, port, mode, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import gnerf_b200 as pkg
        scene = O.synthetic_scene(31, 2, 6, 16, 12, 8, 0.5)
        opts = dict(O.FFHQ_OPTIONS, depth_resolution=12, depth_resolution_importance=8)
        T = torch.from_numpy
        m = 36
        u = scene['u'].reshape(2, m, 8)
        fn = _oracle_local_render(scene['dec'])
        if mode == 'image':
            i = slice(rank, rank + 1)
            out = pkg.parallel.render_sharded(None, T(scene['planes'][i]), None, T(scene['origins'][i]), T(scene['dirs'][i]), opts,
                                              noise=(T(scene['jitter'][i]), T(u[i].reshape(-1, 8))), local_render=fn)
        else:
            out = pkg.parallel.render_ray_sharded(None, T(scene['planes'][:1]), None, T(scene['origins'][:1]), T(scene['dirs'][:1]), opts,
                                                  noise=(T(scene['jitter'][:1]), T(u[:1].reshape(-1, 8))), local_render=fn)
        ret[rank] = [o.numpy().copy() for o in out]
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('mode', ['image', 'rays'])
def test_two_rank_gloo_matches_single_process(mode):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, ret), nprocs=world, join=True)
    scene = O.synthetic_scene(31, 2, 6, 16, 12, 8, 0.5)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=12, depth_resolution_importance=8)
    n = 2 if mode == 'image' else 1
    want = O.render(scene['planes'][:n], scene['dec'], scene['origins'][:n], scene['dirs'][:n], opts,
                    scene['jitter'][:n], scene['u'][:n * 36])
    for r in range(world):                       # every rank ends up with the whole job's outputs
        for got, w in zip(ret[r], want):
            assert got.shape == w.shape
            np.testing.assert_allclose(got, w, atol=1e-6, rtol=0)


def test_point_slabs_cover_every_point_once(pkg):
    P = pkg.parallel
    for n_pts, world in [(256 ** 3, 8), (1000, 3), (5, 8), (7, 1), (10_000_000, 8)]:
        slabs = P.point_slabs(n_pts, world)
        assert len(slabs) == world and slabs[0].start == 0 and slabs[-1].stop == n_pts
        assert all(a.stop == b.start for a, b in zip(slabs, slabs[1:]))
        assert max(map(len, slabs)) - min(map(len, slabs)) <= 1
    with pytest.raises(ValueError):
        P.point_slabs(0, 2)


def _slab_worker(rank, world, port, n_img, n_pts, want_rgb, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import gnerf_b200 as pkg
        scene = O.synthetic_scene(41, n_img, 4, 16, 8, 8, 0.5)
        pts = (np.random.RandomState(5).random_sample((n_img, n_pts, 3)).astype(np.float32) - 0.5) * 1.2

        def query(xyz):                                   # the oracle stands in for the per-rank CUDA point query
            rgb, sigma = O.run_model(scene['planes'], scene['dec'], xyz.numpy(), 1.0)
            return {'rgb': torch.from_numpy(rgb) if want_rgb else None, 'sigma': torch.from_numpy(sigma)}
        out = pkg.parallel.run_model_sharded(None, None, None, torch.from_numpy(pts), None, {'box_warp': 1}, want_rgb=want_rgb,
                                             local_query=query)
        mine = pkg.parallel.run_model_sharded(None, None, None, torch.from_numpy(pts), None, {'box_warp': 1}, want_rgb=want_rgb,
                                              local_query=query, gather=False)
        ret[rank] = (out['sigma'].numpy().copy(), out['rgb'].numpy().copy() if want_rgb else None, mine['sigma'].numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_img,n_pts,want_rgb', [(1, 64, False), (1, 65, True), (2, 33, False)])
def test_two_rank_gloo_run_model_sharded(pkg, n_img, n_pts, want_rgb):
    """Config 5's sharding: contiguous point slabs per rank, sigma (and rgb) all-gathered; even and
    ragged slab sizes, one and two images."""
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_slab_worker, args=(world, _free_port(), n_img, n_pts, want_rgb, ret), nprocs=world, join=True)
    scene = O.synthetic_scene(41, n_img, 4, 16, 8, 8, 0.5)
    pts = (np.random.RandomState(5).random_sample((n_img, n_pts, 3)).astype(np.float32) - 0.5) * 1.2
    rgb, sigma = O.run_model(scene['planes'], scene['dec'], pts, 1.0)
    slabs = pkg.parallel.point_slabs(n_pts, world)
    for r in range(world):
        got_sigma, got_rgb, mine = ret[r]
        np.testing.assert_allclose(got_sigma, sigma, atol=2e-6, rtol=0)       # (numpy matmul blocks differently per slab size)
        if want_rgb:
            np.testing.assert_allclose(got_rgb, rgb, atol=2e-6, rtol=0)
        np.testing.assert_allclose(mine, sigma[:, slabs[r].start:slabs[r].stop], atol=2e-6, rtol=0)
