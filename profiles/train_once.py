"""A few training steps (forward with kept samples + backward) at config 2, for ncu: python profiles/train_once.py [reps]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
pkg = importlib.import_module('g-nerf_b200')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device('cuda:0')
planes_h, c2w, K = bench.make_inputs(torch, 100)
planes = planes_h.to(dev).requires_grad_(True)
dec = bench.make_decoder(torch, pkg, dev, 0).requires_grad_(True)
o, d = pkg.RaySampler()(c2w.to(dev), K.to(dev), bench.RES)
R = pkg.ImportanceRenderer()
n, m = o.shape[:2]
A, B, C = torch.randn(n, m, 32, device=dev), torch.randn(n, m, 1, device=dev), torch.randn(n, m, 1, device=dev)
for _ in range(reps):
    planes.grad = None
    rgb, depth, wsum = R(planes, dec, o, d, dict(bench.OPTS))
    torch.autograd.backward((rgb, depth, wsum), (A, B, C))
torch.cuda.synchronize()
