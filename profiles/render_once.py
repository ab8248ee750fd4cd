"""A few forwards of one shape (for ncu / phase timing): python profiles/render_once.py <n_img> <res> <dc> <df> [mode] [reps]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
pkg = importlib.import_module('g-nerf_b200')
n, res, dc, df = (int(x) for x in sys.argv[1:5])
mode = sys.argv[5] if len(sys.argv) > 5 else 'fp32'
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
dev = torch.device('cuda:0')
planes_h, c2w, K = bench.make_inputs(torch, 100, n_img=n)
dec = bench.make_decoder(torch, pkg, dev, 0)
R, S = pkg.ImportanceRenderer(), pkg.RaySampler()
planes = planes_h.to(dev); o, d = S(c2w.to(dev), K.to(dev), res)
opts = dict(bench.OPTS, decoder_precision=mode, depth_resolution=dc, depth_resolution_importance=df)
pp = pkg.pack_planes(planes)
for _ in range(reps):
    R(pp, dec, o, d, opts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    R(pp, dec, o, d, opts)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'{n} x {res}^2 x ({dc}+{df}) {mode}: {ms:.3f} ms, {n * res * res * (dc + df) / ms / 1e6:.3f} G ray-samples/s')
