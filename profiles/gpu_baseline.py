"""The "GPU baseline" of SURVEY.md section 8(d): what the reference's eager-PyTorch renderer costs on the same B200.
/root/reference cannot travel to the GPU box, so this times oracle/torch_oracle.py -- the same ATen calls, pinned
bit-identical to the reference on CPU by tests/test_torch_oracle.py -- at config 2 (fp32, TF32 off as in
training_loop.py:145-146), next to the fused kernel on the same inputs.  Measurement script, not product code.
Usage: python profiles/gpu_baseline.py > gpurun_out/gpu_baseline.json"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import torch_oracle as TO
pkg = importlib.import_module('g-nerf_b200')
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device('cuda:0')
planes_h, c2w, K = bench.make_inputs(torch, 100)
planes = planes_h.to(dev)
dec = bench.make_decoder(torch, pkg, dev, 0)
o, d = pkg.RaySampler()(c2w.to(dev), K.to(dev), bench.RES)
opts = dict(bench.OPTS)
n, m = o.shape[:2]
tdec = (dec.net[0].weight.detach(), dec.net[0].bias.detach(), dec.net[2].weight.detach(), dec.net[2].bias.detach(), 1.0)
R = pkg.ImportanceRenderer()


def timed(fn, warm=3, reps=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def eager():
    with torch.no_grad():
        jitter = torch.rand((n, m, bench.DC, 1), device=dev)
        u = torch.rand((n * m, bench.DF), device=dev)
        return TO.render(planes, tdec, o, d, opts, jitter, u)


samples = n * m * (bench.DC + bench.DF)
torch.cuda.reset_peak_memory_stats()
ms_eager = timed(eager)
peak = torch.cuda.max_memory_allocated()
ms_fused = timed(lambda: R(planes, dec, o, d, opts))
ms_fused16 = timed(lambda: R(planes, dec, o, d, dict(opts, decoder_precision='bf16')))
print(json.dumps({'workload': bench.WORKLOAD,
                  'eager_torch_restatement': {'ms': ms_eager, 'ray_samples_per_s': samples / ms_eager * 1e3,
                                              'peak_mem_GB': peak / 1e9, 'tf32': False},
                  'fused_fp32': {'ms': ms_fused, 'ray_samples_per_s': samples / ms_fused * 1e3},
                  'fused_bf16': {'ms': ms_fused16, 'ray_samples_per_s': samples / ms_fused16 * 1e3},
                  'speedup_fp32': ms_eager / ms_fused, 'speedup_bf16': ms_eager / ms_fused16}))
