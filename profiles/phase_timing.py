"""Per-role wait / work cycle breakdown of CTA 0 of the render kernels (TPR_PHASE_TIMING=1 makes the kernels
write clock64 deltas into the scratch buffer).  Usage: python profiles/phase_timing.py [fp32] [bf16]"""
import os, sys, importlib
os.environ['TPR_PHASE_TIMING'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
pkg = importlib.import_module('g-nerf_b200')
import bench
dev = torch.device('cuda:0')
planes_h, c2w, K = bench.make_inputs(torch, 100)
dec = bench.make_decoder(torch, pkg, dev, 0)
R, S = pkg.ImportanceRenderer(), pkg.RaySampler()
planes = planes_h.to(dev); o, d = S(c2w.to(dev), K.to(dev), 128)
WS = ['GATHER wait coarse_ready', 'GATHER wait fine_ready', 'GATHER wait a1_free', 'GATHER gather+publish',
      'DECODE wait a1_full (issuer)', 'DECODE issue M1 + slot', 'DECODE wait d1_full', 'DECODE epilogue1', 'DECODE wait a2_full (issuer)',
      'DECODE issue M2', 'DECODE sigma halves exchange (named barrier)', 'DECODE publish sigma (arrive)', 'RAYS setup', 'RAYS wait csig', 'RAYS resample',
      'RAYS wait fsig', 'RAYS sort+march', 'RAYS composite', 'RAYS   (merge / pair rank count at 96+96: part of the sort, not counted in sort+march)']
D = int(os.environ.get('TPR_PT_DEPTH', '48'))          # samples per pass (96 = gen_videos / config 4)
RPG = 8 if D <= 48 else 4                                # rays per group the kernel picks
for mode in sys.argv[1:] or ['fp32']:
    opts = dict(bench.OPTS, decoder_precision=mode, depth_resolution=D, depth_resolution_importance=D)
    for _ in range(3): R(planes, dec, o, d, opts)
    torch.cuda.synchronize()
    t = R.last_scratch[64:64 + 24 * 8].view(torch.int64).cpu().numpy()
    ngroups = 16384 * 8 / RPG / 148
    names = WS
    print(mode, 'CTA0 cycles per group (', int(ngroups), 'groups )')
    for n, v in zip(names, t): print(f'  {n:32s} {int(v / ngroups):7d}')
