"""Which gather shape serves random 128-byte lines fastest on this GPU?  (tpr_gather_microbench_v2: LDG.128 x 8 lanes vs
LDG.256 x 4 lanes per line, burst vs software-pipelined, warps per SM.)  25 MB working set = one image's planes, L2 resident.
Usage: python profiles/gather_shapes.py > gpurun_out/gather_shapes.json"""
import ctypes, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module('g-nerf_b200')
L = pkg._lib.bench_lib()
n_lines = 25 * (1 << 20) // 128
buf = torch.randn(n_lines * 32, device='cuda')
sink = torch.empty(65536, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def measure(threads, vec, fl, pipe, iters):
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lines = L.tpr_gather_microbench_v2(ctypes.c_void_p(buf.data_ptr()), n_lines, 148, threads, vec, fl, pipe, iters,
                                           ctypes.c_void_p(sink.data_ptr()), st)
        e1.record(); torch.cuda.synchronize()
        assert lines > 0, lines
        best = max(best, lines * 128 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return round(best, 1)


out = {}
for threads in (512, 768, 1024):
    row = {}
    for vec, fl, pipe in ((4, 2, 0), (4, 2, 1), (4, 4, 0), (4, 4, 1), (4, 6, 0), (4, 8, 0), (8, 1, 0), (8, 1, 1), (8, 2, 0), (8, 2, 1), (8, 4, 0), (8, 4, 1)):
        row[f'ldg{vec * 32}_x{fl}{"_pipelined" if pipe else ""}'] = measure(threads, vec, fl, pipe, 2 * (600 // (fl * (2 if vec == 8 else 1))))
    out[f'{threads // 32}warps'] = row
print(json.dumps(out, indent=1))
