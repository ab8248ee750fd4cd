"""One small render through the C ABI, checked against the numpy oracle (debug aid).
Usage: python profiles/one_render.py [fp32|bf16|fp32_ffma] [res] [n_img]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import triplane_oracle as O
pkg = importlib.import_module('g-nerf_b200')
mode = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
res = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n_img = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device('cuda:0')
scene = O.synthetic_scene(seed=7, n_img=n_img, res=res, plane_res=64, dc=48, df=48, bias_scale=0.5)
opts = dict(O.FFHQ_OPTIONS, decoder_precision=mode)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dec = pkg.OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
with torch.no_grad():
    dec.net[0].weight.copy_(torch.from_numpy(scene['dec'].w1)); dec.net[0].bias.copy_(torch.from_numpy(scene['dec'].b1))
    dec.net[2].weight.copy_(torch.from_numpy(scene['dec'].w2)); dec.net[2].bias.copy_(torch.from_numpy(scene['dec'].b2))
dec = dec.to(dev).requires_grad_(False)
o, d = pkg.RaySampler()(T(scene['c2w']), T(scene['K']), res)
rgb, depth, wsum = pkg.ImportanceRenderer()(T(scene['planes']), dec, o, d, opts, noise=(T(scene['jitter']), T(scene['u'])))
torch.cuda.synchronize()
want = O.render(scene['planes'], scene['dec'], scene['origins'], scene['dirs'], dict(O.FFHQ_OPTIONS), scene['jitter'], scene['u'])
print(mode, res, n_img, 'max-abs rgb/depth/wsum vs oracle =', [float(np.abs(g.cpu().numpy() - w).max()) for g, w in zip((rgb, depth, wsum), want)])
