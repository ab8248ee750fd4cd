import os, sys, importlib
os.environ['TPR_PHASE_TIMING'] = '1'
sys.path.insert(0, '/root/repo')
import torch, numpy as np
pkg = importlib.import_module('g-nerf_b200')
import bench
dev = torch.device('cuda:0')
planes_h, c2w, K = bench.make_inputs(torch, dev, 100)
dec = bench.make_decoder(torch, pkg, dev, 0)
R, S = pkg.ImportanceRenderer(), pkg.RaySampler()
planes = planes_h.to(dev); o, d = S(c2w.to(dev), K.to(dev), 128)
names = ['setup','G0+sync','issueM1','G(t+1)','wait bar1/2','E1+sync','issueM2','pass-end wait+sigma','resample','sort','composite']
for mode in sys.argv[1:] or ['fp32']:
    opts = dict(bench.OPTS, decoder_precision=mode)
    for _ in range(3): R(planes, dec, o, d, opts)
    torch.cuda.synchronize()
    t = R.last_scratch[64:64+16*8].view(torch.int64).cpu().numpy()
    tot = t.sum(); ngroups = 16384*8/8/148
    print(mode, 'CTA0 total cycles', tot, 'per group', int(tot/ngroups))
    for n, v in zip(names, t): print(f'  {n:22s} {v/tot*100:5.1f}%  {int(v/ngroups):6d} cyc/group')
