"""Training forward (kept samples) at config 2; TPR_TRAIN_DEBUG=1 skips the kept colours, 2 the kept features (timing A/B only)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
pkg = importlib.import_module('g-nerf_b200')
dev = torch.device('cuda:0')
planes_h, c2w, K = bench.make_inputs(torch, 100)
planes = planes_h.to(dev).requires_grad_(True)
dec = bench.make_decoder(torch, pkg, dev, 0).requires_grad_(True)
o, d = pkg.RaySampler()(c2w.to(dev), K.to(dev), bench.RES)
R = pkg.ImportanceRenderer()
def fwd():
    R(planes, dec, o, d, dict(bench.OPTS))
for _ in range(3): fwd()
torch.cuda.synchronize()
ts=[]
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fwd(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get('TPR_TRAIN_DEBUG','0'), 'train forward ms', sorted(ts)[5])
