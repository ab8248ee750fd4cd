"""Multi-GPU rendering: one process per GPU, rays are independent, one small exchange at the end.

Partition: by image first -- a GPU renders the images whose planes it holds, so
planes never move -- then by contiguous ray range inside an image when there are fewer images than GPUs.
The only data-path collectives are
  * an all-reduce (MIN, MAX) of two scalars: the global depth range MipRayMarcher2 clamps against
    (training/volumetric_rendering/ray_marcher.py:50 reduces over the whole batch), and
  * an in-place all-gather of rgb [.,32] + depth [.,1] + weight_sum [.,1] (136 B per ray): every rank
    renders straight into its own slice of the gather buffers, so there is no staging copy.
The host logic is backend-agnostic (NCCL on GPUs, gloo in the CPU tests); the per-shard render is the fused
CUDA renderer unless a test injects another callable.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    image: int          # index into the global image batch
    ray_begin: int      # [ray_begin, ray_end) of that image
    ray_end: int

    @property
    def n_rays(self):
        return self.ray_end - self.ray_begin


def partition(n_img: int, n_rays: int, world: int) -> List[List[Shard]]:
    """Work list per rank.  n_img >= world: whole images, as evenly as possible (the first n_img % world
    ranks get one more).  n_img < world: ranks are dealt to images round-robin and each image's rays are cut
    into contiguous ranges over the ranks that share it."""
    if n_img <= 0 or n_rays <= 0 or world <= 0:
        raise ValueError('partition: n_img, n_rays and world must be positive')
    plan: List[List[Shard]] = [[] for _ in range(world)]
    if n_img >= world:
        base, extra = divmod(n_img, world)
        start = 0
        for r in range(world):
            cnt = base + (1 if r < extra else 0)
            plan[r] = [Shard(i, 0, n_rays) for i in range(start, start + cnt)]
            start += cnt
        return plan
    ranks_of = [[r for r in range(world) if r % n_img == i] for i in range(n_img)]
    for i, ranks in enumerate(ranks_of):
        k = len(ranks)
        base, extra = divmod(n_rays, k)
        begin = 0
        for j, r in enumerate(ranks):
            cnt = base + (1 if j < extra else 0)
            if cnt > 0:
                plan[r].append(Shard(i, begin, begin + cnt))
            begin += cnt
    return plan


def shard_ray_counts(plan: Sequence[Sequence[Shard]]) -> List[int]:
    return [sum(s.n_rays for s in shards) for shards in plan]


def _default_local_render(renderer, planes, decoder, origins, dirs, options, noise, out):
    """Render one rank's rays with the fused CUDA renderer, leaving the global depth clamp to the caller."""
    renderer.defer_depth_clamp = True
    try:
        rgb, depth, wsum = renderer(planes, decoder, origins, dirs, options, noise=noise, out=out)
    finally:
        renderer.defer_depth_clamp = False
    return rgb, depth, wsum, renderer.last_depth_range


def render_sharded(renderer, planes, decoder, ray_origins, ray_directions, rendering_options, *,
                   group=None, noise=None, local_render: Optional[Callable] = None, gather: bool = True):
    """Render THIS rank's images and return the whole job's outputs.

    planes [n_local,3,32,H,W], ray_origins / ray_directions [n_local,M,3]: this rank's share (image-sharded:
    every rank holds the same number of images).  Returns (rgb [world*n_local,M,32], depth [...,1],
    weight_sum [...,1]) in rank order when ``gather`` is true, else this rank's slice after the global clamp.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n, m, _ = ray_origins.shape
    dev = ray_origins.device
    local_render = local_render or _default_local_render
    # gather buffers; this rank renders directly into its slice
    bufs = [torch.empty((world, n, m, c), device=dev, dtype=torch.float32) for c in (32, 1, 1)]
    out = tuple(b[rank] for b in bufs)
    rgb, depth, wsum, rng = local_render(renderer, planes, decoder, ray_origins, ray_directions, rendering_options, noise, out)
    for dst, src in zip(out, (rgb, depth, wsum)):
        if dst.data_ptr() != src.data_ptr():          # a local_render that ignored `out`
            dst.copy_(src)
    # the one cross-ray dependency: clamp(depth, min(all depths), max(all depths))
    lo, hi = rng[0:1].clone(), rng[1:2].clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    d = out[1]
    if d.is_cuda:
        _clamp_cuda(d, lo, hi)
    else:                                             # gloo / CPU tests
        d.copy_(torch.clamp(torch.nan_to_num(d, nan=float('inf')), lo, hi))
    if world == 1 or not gather:
        return out
    for b in bufs:
        dist.all_gather_into_tensor(b.view(-1), b[rank].reshape(-1), group=group)      # in place
    return tuple(b.view(world * n, m, -1) for b in bufs)


def _clamp_cuda(depth, lo, hi):
    import ctypes
    from . import _lib
    rr = torch.cat([lo, hi]).contiguous()
    with torch.cuda.device(depth.device):
        _lib.check(_lib.lib().tpr_clamp_depth(ctypes.c_void_p(depth.data_ptr()), depth.numel(),
                                              ctypes.c_void_p(rr.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), 'tpr_clamp_depth')
    return depth


def render_ray_sharded(renderer, planes, decoder, ray_origins, ray_directions, rendering_options, *,
                       group=None, noise=None, local_render: Optional[Callable] = None):
    """Fewer images than GPUs: every rank holds ALL images (planes are 25 MB each) and renders the contiguous
    ray ranges `partition` assigns to it; the full [N,M,.] outputs are assembled on every rank with an
    all-gather of padded per-rank buffers."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n, m, _ = ray_origins.shape
    dev = ray_origins.device
    plan = partition(n, m, world)
    local_render = local_render or _default_local_render
    cap = max(max((s.n_rays for s in shards), default=0) for shards in plan)
    per_rank = max(len(shards) for shards in plan)
    buf = torch.zeros((world, per_rank, cap, 34), device=dev, dtype=torch.float32)
    lo = torch.full((1,), float('inf'), device=dev)
    hi = torch.full((1,), float('-inf'), device=dev)
    for k, s in enumerate(plan[rank]):
        sl = slice(s.ray_begin, s.ray_end)
        nz = None
        if noise is not None:
            jit, u = noise
            df = u.shape[-1]
            nz = (jit[s.image:s.image + 1, sl], u.reshape(n, m, df)[s.image, sl].reshape(-1, df))
        rgb, depth, wsum, rng = local_render(renderer, planes[s.image:s.image + 1], decoder, ray_origins[s.image:s.image + 1, sl],
                                             ray_directions[s.image:s.image + 1, sl], rendering_options, nz, None)
        buf[rank, k, :s.n_rays, :32] = rgb[0]
        buf[rank, k, :s.n_rays, 32:33] = depth[0]
        buf[rank, k, :s.n_rays, 33:34] = wsum[0]
        lo = torch.minimum(lo, rng[0:1]); hi = torch.maximum(hi, rng[1:2])
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        dist.all_gather_into_tensor(buf.view(-1), buf[rank].reshape(-1), group=group)
    rgb = torch.empty((n, m, 32), device=dev); depth = torch.empty((n, m, 1), device=dev); wsum = torch.empty((n, m, 1), device=dev)
    for r, shards in enumerate(plan):
        for k, s in enumerate(shards):
            sl = slice(s.ray_begin, s.ray_end)
            rgb[s.image, sl] = buf[r, k, :s.n_rays, :32]
            depth[s.image, sl] = buf[r, k, :s.n_rays, 32:33]
            wsum[s.image, sl] = buf[r, k, :s.n_rays, 33:34]
    depth = torch.clamp(torch.nan_to_num(depth, nan=float('inf')), lo, hi)
    return rgb, depth, wsum
