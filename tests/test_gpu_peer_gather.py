"""GPU: render + gather in one kernel (tpr_render_peers, parallel.PeerGather).

One GPU: the kernel epilogue's extra stores, with ordinary local buffers standing in for the peer-mapped ones --
every sink must receive bit for bit what the primary outputs receive (before the deferred depth clamp).
Two GPUs (skipped on a one-GPU box): two NCCL ranks, real IPC-mapped peer buffers; the peer-store path must return
exactly what the render + all-gather path returns, on both alternating buffer sets.
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch

from oracle import triplane_oracle as O
from tests.test_gpu_parity import T, make_decoder, dev

pytestmark = pytest.mark.gpu


def _sinks(pkg, tensors):
    s = pkg._lib.TprPeerSinks()
    s.n_peers = len(tensors)
    for i, (rgb, depth, wsum) in enumerate(tensors):
        s.rgb[i], s.depth[i], s.weight_sum[i] = rgb.data_ptr(), depth.data_ptr(), wsum.data_ptr()
    return s


@pytest.mark.parametrize('mode,dc,df', [('fp32', 48, 48), ('bf16', 48, 48), ('fp32', 20, 13), ('fp32', 32, 0),
                                        ('fp32_ffma', 24, 24), ('fp32', 120, 120)])
@pytest.mark.parametrize('layout', ['channels_last', 'channels_first'])
def test_every_sink_receives_the_primary_outputs(pkg, mode, dc, df, layout):
    n, res = 2, 12
    scene = O.synthetic_scene(71, n, res, 48, dc, df, 0.5)
    opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df, decoder_precision=mode,
                output_layout=layout)
    R, dec = pkg.ImportanceRenderer(), make_decoder(pkg, scene['dec'])
    m = res * res
    noise = (T(scene['jitter']), T(scene['u']) if df > 0 else None)

    def bufs():
        rgb = torch.full((n, 32, m) if layout == 'channels_first' else (n, m, 32), float('nan'), device=dev())
        if layout == 'channels_first':
            rgb = rgb.permute(0, 2, 1)
        return rgb, torch.full((n, m, 1), float('nan'), device=dev()), torch.full((n, m, 1), float('nan'), device=dev())

    own, peers = bufs(), [bufs() for _ in range(3)]
    sinks = _sinks(pkg, [(p[0].permute(0, 2, 1) if layout == 'channels_first' else p[0], p[1], p[2]) for p in peers])
    R(T(scene['planes']), dec, T(scene['origins']), T(scene['dirs']), opts, noise=noise, out=own, peer_sinks=sinks)
    # the same call without sinks and with the clamp deferred is the expected content of every buffer
    R.defer_depth_clamp = True
    want = R(T(scene['planes']), dec, T(scene['origins']), T(scene['dirs']), opts, noise=noise)
    R.defer_depth_clamp = False
    for got in [own] + peers:
        for g, w in zip(got, want):
            assert torch.equal(torch.nan_to_num(g, nan=-7.0), torch.nan_to_num(w, nan=-7.0))
    assert torch.isfinite(own[0]).all()


def test_peer_render_rejects_an_in_kernel_clamp_and_bad_counts(pkg):
    lib = pkg._lib.lib()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    o = pkg._lib.TprOptions(ray_start=2.25, ray_end=3.3, box_warp=1.0, depth_resolution=8, depth_resolution_importance=8)
    s = pkg._lib.TprPeerSinks()
    s.n_peers = 16
    rc = lib.tpr_render_peers(p, 1, 8, 8, p, p, p, 4, p, p, None, None, ctypes.byref(o), p, p, p, None, p, 1024,
                              ctypes.byref(s), None)
    assert rc == -2 and b'n_peers' in lib.tpr_last_error()
    s.n_peers = 1                                        # pointer left NULL
    rc = lib.tpr_render_peers(p, 1, 8, 8, p, p, p, 4, p, p, None, None, ctypes.byref(o), p, p, p, None, p, 1024,
                              ctypes.byref(s), None)
    assert rc == -1 and b'peer pointer' in lib.tpr_last_error()
    rc = lib.tpr_render_peers(p, 1, 8, 8, p, p, p, 4, p, p, None, None, ctypes.byref(o), p, p, p, None, p, 1024, None, None)
    assert rc == -1


def test_peer_buffer_is_a_torch_tensor(pkg):
    """tpr_peer_alloc memory wrapped through __cuda_array_interface__: torch reads and writes it in place."""
    lib = pkg._lib.lib()
    ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(pkg._lib.PEER_HANDLE_BYTES)
    pkg._lib.check(lib.tpr_peer_alloc(4096, ctypes.byref(ptr), handle), 'tpr_peer_alloc')
    try:
        t = torch.as_tensor(pkg.parallel._DeviceBlock(ptr.value, 1024), device=dev())
        assert t.data_ptr() == ptr.value and t.is_cuda and t.dtype == torch.float32
        t.fill_(3.0)
        assert float(t.sum()) == 3072.0
        assert any(handle.raw)
        del t
    finally:
        torch.cuda.synchronize()
        pkg._lib.check(lib.tpr_peer_free(ptr), 'tpr_peer_free')


# ---------------------------------------------------------------------------------------------------- two GPUs
def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _two_rank_worker(rank, world, port, q):
    import torch.distributed as dist
    import gnerf_b200 as pkg
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    d = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=d)
    try:
        n, res, dc, df = 2, 16, 48, 48
        m = res * res
        opts = dict(O.FFHQ_OPTIONS, depth_resolution=dc, depth_resolution_importance=df)
        R = pkg.ImportanceRenderer()
        peer = pkg.parallel.PeerGather(n, m)
        assert peer.sets == 3
        held = []                                      # (step, gathered outputs still living in the peer buffers, expected)
        for step in range(5):                          # every buffer set, and the first two again
            scene = O.synthetic_scene(300 + 10 * step + rank, n, res, 48, dc, df, 0.5)
            dec = make_decoder(pkg, O.synthetic_scene(300, n, res, 48, dc, df, 0.5)['dec'], device=d)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)       # noqa: E731
            args = (t(scene['planes']), dec, t(scene['origins']), t(scene['dirs']), opts)
            noise = (t(scene['jitter']), t(scene['u']))
            want = pkg.parallel.render_sharded(R, *args, noise=noise)
            want = tuple(w.clone() for w in want)
            got = pkg.parallel.render_sharded(R, *args, noise=noise, peer=peer)
            torch.cuda.synchronize()
            for g, w in zip(got, want):
                assert g.shape == w.shape and torch.equal(g, w), f'rank {rank} step {step}: peer gather differs'
            assert got[0].shape == (world * n, m, 32)
            # the lifetime contract of PeerGather (three sets): the outputs of call s may still be read after call s+1 --
            # consume step s-1 NOW, after this step's render + gather, while the other rank may already be a step ahead
            if rank == 1:
                torch.cuda._sleep(20_000_000)          # the slower rank: ~10 ms behind its peer
            for hs, hgot, hwant in held:
                if hs == step - 1:
                    for g, w in zip(hgot, hwant):
                        assert torch.equal(g, w), f'rank {rank}: outputs of step {hs} were overwritten before step {step + 1}'
            held = [(step, got, want)]
        peer.close()
        q.put((rank, 'ok'))
    except Exception as e:              # noqa: BLE001
        q.put((rank, f'{type(e).__name__}: {e}'))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_two_ranks_peer_gather_equals_all_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res
