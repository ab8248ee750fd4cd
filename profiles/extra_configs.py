"""Throughput of the non-headline BASELINE.json configs on one B200 (CUDA events, 3 warm-up + 10 timed):
  config 4  stress render: batch 32 x 256^2 rays x (96+96) samples  (per-GPU share of it: --n-img)
  config 5  density grid : 256^3 points through run_model, sigma only (gen_videos.py:33-55,198-209)
Usage: python profiles/extra_configs.py [--n-img 32] [--mode fp32|bf16]"""
import argparse, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
pkg = importlib.import_module('g-nerf_b200')
ap = argparse.ArgumentParser(); ap.add_argument('--n-img', type=int, default=32); ap.add_argument('--mode', default='fp32')
ap.add_argument('--skip4', action='store_true'); ap.add_argument('--skip5', action='store_true')
args = ap.parse_args()
dev = torch.device('cuda:0')
dec = bench.make_decoder(torch, pkg, dev, 0)
R, S = pkg.ImportanceRenderer(), pkg.RaySampler()


def timed(fn, warm=3, n=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = {}
if not args.skip4:
    n = args.n_img
    planes_h, c2w, K = bench.make_inputs(torch, 7, n_img=n)
    planes = planes_h.to(dev); del planes_h
    o, d = S(c2w.to(dev), K.to(dev), 256)
    opts = dict(bench.OPTS, depth_resolution=96, depth_resolution_importance=96, decoder_precision=args.mode)
    ms = timed(lambda: R(planes, dec, o, d, opts))
    samples = n * 256 * 256 * 192
    out['config4'] = {'n_img': n, 'rays': 256 * 256, 'samples_per_ray': 192, 'mode': args.mode, 'ms': ms,
                      'ray_samples_per_s': samples / ms * 1e3, 'gather_GBps_1536B': samples * 1536 / ms / 1e6}
    pp = pkg.pack_planes(planes)
    ms = timed(lambda: R(pp, dec, o, d, opts))
    out['config4_prepacked'] = {'ms': ms, 'ray_samples_per_s': samples / ms * 1e3}
    del planes, pp, o, d
    torch.cuda.empty_cache()
if not args.skip5:
    planes_h, _, _ = bench.make_inputs(torch, 9, n_img=1)
    pp = pkg.pack_planes(planes_h.to(dev))
    g = 256
    # gen_videos.create_samples: voxel centres of a cube of side box_warp, x fastest
    ax = (torch.arange(g, device=dev, dtype=torch.float32) + 0.5) / g - 0.5
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing='ij')
    xyz = torch.stack([xx, yy, zz], -1).reshape(1, -1, 3).contiguous()
    for want_rgb in (False, True):
        ms = timed(lambda: R.run_model(pp, dec, xyz, None, dict(bench.OPTS, decoder_precision=args.mode), want_rgb=want_rgb))
        out['config5' + ('_rgb' if want_rgb else '')] = {'points': g ** 3, 'ms': ms, 'points_per_s': g ** 3 / ms * 1e3,
                                                        'gather_GBps_1536B': g ** 3 * 1536 / ms / 1e6}
    # the same points in random order (no locality between neighbouring queries)
    perm = torch.randperm(g ** 3, device=dev)
    xyz_r = xyz[:, perm].contiguous()
    ms = timed(lambda: R.run_model(pp, dec, xyz_r, None, dict(bench.OPTS, decoder_precision=args.mode), want_rgb=False))
    out['config5_shuffled'] = {'points': g ** 3, 'ms': ms, 'points_per_s': g ** 3 / ms * 1e3}
# config 3 (renderer part): the gen_videos.py orbit of one identity -- 120 frames x 64^2 rays x (96+96) samples
ap3 = dict(bench.OPTS, depth_resolution=96, depth_resolution_importance=96, decoder_precision=args.mode)
planes_h, _, _ = bench.make_inputs(torch, 11, n_img=1)
planes = planes_h.to(dev)
cams = importlib.import_module('g-nerf_b200.camera_utils')
c2w, K = cams.orbit_cameras(120)
c2w, K = torch.from_numpy(c2w).to(dev), torch.from_numpy(K).to(dev)
samples = 120 * 64 * 64 * 192


def frame_loop():                                   # what gen_videos.py does: one forward (and one repack) per frame
    for f in range(120):
        o, d = S(c2w[f:f + 1], K[f:f + 1], 64)
        R(planes, dec, o, d, ap3)


def frame_loop_cached():
    R.cache_packed_planes = True
    for f in range(120):
        o, d = S(c2w[f:f + 1], K[f:f + 1], 64)
        R(planes, dec, o, d, ap3)
    R.cache_packed_planes = False


ms = timed(frame_loop, warm=1, n=3)
out['config3_frame_loop'] = {'ms_per_orbit': ms, 'ray_samples_per_s': samples / ms * 1e3}
ms = timed(frame_loop_cached, warm=1, n=3)
out['config3_frame_loop_plane_cache'] = {'ms_per_orbit': ms, 'ray_samples_per_s': samples / ms * 1e3}
ms = timed(lambda: pkg.render_frames(R, planes, dec, c2w, K, 64, ap3), warm=1, n=5)
out['config3_render_frames'] = {'ms_per_orbit': ms, 'ray_samples_per_s': samples / ms * 1e3}
print(json.dumps(out))
