"""Ceiling of the backward pass's plane-gradient scatter (tpr_scatter_microbench): red.global.add.v4.f32, eight lanes per
128-byte line, random lines of a buffer the size of config 2's plane gradient (8 x 3 x 256^2 x 32 fp32 = 201 MB) and of one
image's (25 MB, L2 resident).  Measurement script.  Usage: python profiles/scatter_shapes.py > gpurun_out/r02_scatter_shapes.json"""
import ctypes, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module('g-nerf_b200')
L = pkg._lib.bench_lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
sms = torch.cuda.get_device_properties(0).multi_processor_count
out = []
for mb in (25, 201):
    n_lines = mb * (1 << 20) // 128
    buf = torch.zeros(n_lines * 32, device='cuda')
    for threads, ctas_per_sm in ((256, 1), (512, 1), (1024, 1), (1024, 2)):
        ctas = sms * ctas_per_sm
        iters = 64
        best = None
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lines = L.tpr_scatter_microbench(ctypes.c_void_p(buf.data_ptr()), n_lines, ctas, threads, 12, iters, st)
            e1.record(); torch.cuda.synchronize()
            assert lines > 0, lines
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        out.append({'buffer_MB': mb, 'threads_per_sm': threads * ctas_per_sm, 'lines': lines, 'ms': best,
                    'TB_per_s': lines * 128 / best / 1e9, 'G_lines_per_s': lines / best / 1e6,
                    'config2_scatter_floor_ms': 8 * 128 * 128 * 96 * 12 / (lines / best)})
print(json.dumps(out, indent=1))
