"""tcgen05.mma issue / completion cost for the decoder's small shapes (tpr_mma_microbench)."""
import ctypes, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module('g-nerf_b200')
L = pkg._lib.bench_lib()
out = torch.zeros(2, dtype=torch.int64, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print('issue  kind  A-from  N    count  issue cyc/MMA  total cyc/MMA')
for tight in (0, 1):
  for bf16 in (0, 1):
    for ts in (0, 1):
        for n in (32, 64, 128, 256):
            for count in (12, 96):
                for _ in range(2):
                    rc = L.tpr_mma_microbench(n, bf16, ts, count, tight, ctypes.c_void_p(out.data_ptr()), st)
                    assert rc == 0, rc
                    torch.cuda.synchronize()
                a, b = out.tolist()
                print(f"{'tight' if tight else 'naive'}  {'bf16' if bf16 else 'tf32'}  {'tmem' if ts else 'smem'}  {n:4d} {count:5d}  {a / count:10.1f}  {b / count:12.1f}")
