"""Measure the random-128-byte-line gather bandwidth (L2-resident and DRAM-resident working sets) with
tpr_gather_microbench -- the denominator SURVEY.md section 8(d) asks to report beside the HBM copy peak."""
import ctypes, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module('g-nerf_b200')


def measure(mbytes, ctas=296, threads=512, in_flight=12, iters=400, reps=4):
    L = pkg._lib.bench_lib()
    n_lines = mbytes * (1 << 20) // 128
    buf = torch.randn(n_lines * 32, device='cuda')
    sink = torch.empty(65536, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    best = 0.0
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lines = L.tpr_gather_microbench_ex(ctypes.c_void_p(buf.data_ptr()), n_lines, ctas, threads, in_flight, iters,
                                           ctypes.c_void_p(sink.data_ptr()), st)
        e1.record(); torch.cuda.synchronize()
        assert lines > 0, lines
        best = max(best, lines * 128 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


if __name__ == '__main__':
    out = {}
    if len(sys.argv) > 1 and sys.argv[1] == 'sweep':       # one CTA per SM: warps x lines in flight, 25 MB working set
        for threads in (256, 384, 512, 640, 768, 1024):
            for fl in (4, 6, 12, 24):
                out[f'25MB_148ctas_{threads // 32}warps_{fl}inflight'] = round(measure(25, 148, threads, fl, iters=1200 // fl), 1)
    else:
        for mb in (25, 50, 200, 1600):
            for ctas in (148, 296, 592):
                out[f'{mb}MB_{ctas}ctas_x16warps_12inflight'] = round(measure(mb, ctas), 1)
    print(json.dumps(out, indent=1))
