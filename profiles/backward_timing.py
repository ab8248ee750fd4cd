"""Renderer forward + backward at the BASELINE config-2 shape (8 x 128^2 rays x 48+48) on one B200: the fused CUDA
path (forward kernel + tpr_render_backward) next to autograd through the torch restatement of the reference
(oracle/torch_oracle.py; per image, because eager autograd saves ~10 GB per image at this shape).
Measurement script, not product code.  Usage: python profiles/backward_timing.py [--n-img 8] [--eager-img 1]"""
import argparse, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import torch_oracle as TO
pkg = importlib.import_module('g-nerf_b200')
ap = argparse.ArgumentParser(); ap.add_argument('--n-img', type=int, default=8); ap.add_argument('--eager-img', type=int, default=1)
ap.add_argument('--reps', type=int, default=10); ap.add_argument('--mode', default='fp32'); ap.add_argument('--no-keep', action='store_true'); ap.add_argument('--no-keep-features', action='store_true')
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
planes_h, c2w, K = bench.make_inputs(torch, 100, n_img=args.n_img)
planes = planes_h.to(dev).requires_grad_(True)
dec = bench.make_decoder(torch, pkg, dev, 0).requires_grad_(True)
o, d = pkg.RaySampler()(c2w.to(dev), K.to(dev), bench.RES)
opts = dict(bench.OPTS, decoder_precision=args.mode)
n, m = o.shape[:2]
R = pkg.ImportanceRenderer()
R.keep_samples = not args.no_keep
R.keep_features = not args.no_keep_features
A, B, C = torch.randn(n, m, 32, device=dev), torch.randn(n, m, 1, device=dev), torch.randn(n, m, 1, device=dev)


def timed(fn, warm=3, reps=args.reps):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def fused_step():
    planes.grad = None
    for p in dec.parameters(): p.grad = None
    rgb, depth, wsum = R(planes, dec, o, d, opts)
    torch.autograd.backward((rgb, depth, wsum), (A, B, C))


def fused_fwd():
    with torch.no_grad():
        R(planes, dec, o, d, opts)


samples = n * m * (bench.DC + bench.DF)
out = {'workload': f'{n} x {bench.RES}^2 rays x ({bench.DC}+{bench.DF}) samples, ' + args.mode}
ms_f = timed(fused_fwd)
torch.cuda.reset_peak_memory_stats()
ms_fb = timed(fused_step)
out['fused'] = {'forward_ms': ms_f, 'forward_backward_ms': ms_fb, 'backward_ms': ms_fb - ms_f,
                'ray_samples_per_s_fwd_bwd': samples / ms_fb * 1e3, 'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9}
# eager autograd baseline on the first --eager-img images
k = args.eager_img
if k > 0:
    pl = planes.detach()[:k].clone()
    tdec = tuple(p.detach().clone() for p in (dec.net[0].weight, dec.net[0].bias, dec.net[2].weight, dec.net[2].bias)) + (1.0,)

    def eager_step():
        jitter = torch.rand((k, m, bench.DC, 1), device=dev)
        u = torch.rand((k * m, bench.DF), device=dev)
        TO.render_grads(pl, tdec, o[:k], d[:k], opts, jitter, u, A[:k], B[:k], C[:k])
    torch.cuda.reset_peak_memory_stats()
    ms_e = timed(eager_step, warm=2, reps=5)
    out['eager_torch_autograd'] = {'images': k, 'forward_backward_ms': ms_e, 'ms_per_image': ms_e / k,
                                   'ray_samples_per_s': k * m * (bench.DC + bench.DF) / ms_e * 1e3,
                                   'peak_mem_GB': torch.cuda.max_memory_allocated() / 1e9}
    out['speedup_fwd_bwd_per_image'] = (ms_e / k) / (ms_fb / n)
print(json.dumps(out))
