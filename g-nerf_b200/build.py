"""Build the two C-ABI shared objects in-tree with nvcc for sm_100a (no JIT cache, no torch extension):
  lib/libtriplane_b200.so        the product: include/triplane_b200.h
  lib/libtriplane_b200_bench.so  measurement / bring-up aids only (gather and tcgen05.mma microbenchmarks, the raw
                                 tcgen05 layer test): include/triplane_b200_bench.h.  Never loaded by the renderer."""
import hashlib
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libtriplane_b200.so')
BENCH_LIB_PATH = os.path.join(LIB_DIR, 'libtriplane_b200_bench.so')
SOURCES = ['triplane_b200.cu', 'tpr_render_ws.cu', 'tpr_render_ws_f16x2_a.cu', 'tpr_render_ws_f16x2_b.cu', 'tpr_render_ws_bf16_a.cu',
           'tpr_render_ws_bf16_b.cu', 'tpr_run_model_ws.cu', 'tpr_backward.cu', 'tpr_backward_tc.cu', 'tpr_standalone.cu']
BENCH_SOURCES = ['tpr_microbench.cu', 'tpr_tc_debug.cu']
HEADERS = [os.path.join(CSRC, 'tpr_device.cuh'), os.path.join(CSRC, 'tpr_render.cuh'), os.path.join(CSRC, 'tpr_tc.cuh'),
           os.path.join(CSRC, 'tpr_ws.cuh'), os.path.join(CSRC, 'tpr_render_ws.cuh'), os.path.join(CSRC, 'tpr_host.h'), os.path.join(ROOT, 'include', 'triplane_b200.h'),
           os.path.join(ROOT, 'include', 'triplane_b200_bench.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']
OBJ_DIR = os.path.join(LIB_DIR, 'obj')


def _digest(paths):
    h = hashlib.sha256()
    for f in paths:
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile what changed since the last build (one object per source, compiled in parallel, then linked into the
    product library and the measurement library); return the product library's path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    jobs, objs = [], {LIB_PATH: [], BENCH_LIB_PATH: []}
    for lib, sources in ((LIB_PATH, SOURCES), (BENCH_LIB_PATH, BENCH_SOURCES)):
        for src in sources:
            path = os.path.join(CSRC, src)
            obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
            stamp = obj + '.sha256'
            dig = _digest([path] + HEADERS)
            objs[lib].append(obj)
            if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
                continue
            cmd = [nvcc] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC] + \
                  (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
            jobs.append((lib, src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    relink = set()
    for lib, src, stamp, dig, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
        if verbose:
            print(out)
        with open(stamp, 'w') as fh:
            fh.write(dig)
        relink.add(lib)
    for lib, lib_objs in objs.items():
        # the link stamp catches a changed source LIST (an object dropped from or added to the library)
        link_stamp = lib + '.sha256'
        want = ' '.join(sorted(os.path.basename(o) for o in lib_objs))
        stale = not os.path.exists(link_stamp) or open(link_stamp).read().strip() != want
        if lib in relink or stale or not os.path.exists(lib):
            res = subprocess.run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + lib_objs + ['-o', lib],
                                 capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
            with open(link_stamp, 'w') as fh:
                fh.write(want)
    return LIB_PATH


def build_alt(defines, name='alt', verbose=False) -> str:
    """Development aid for same-box A/B measurements: the product library compiled with extra -D switches into
    lib/libtriplane_b200_<name>.so (load it with TPR_LIB=<path>; see _lib.py).  Not built by build()."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    out_dir = os.path.join(LIB_DIR, 'obj_' + name)
    os.makedirs(out_dir, exist_ok=True)
    objs, jobs = [], []
    for src in SOURCES:
        obj = os.path.join(out_dir, src[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + list(defines) + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC] + \
              (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
        if verbose:
            print(out)
    lib = os.path.join(LIB_DIR, f'libtriplane_b200_{name}.so')
    res = subprocess.run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + objs + ['-o', lib], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('link failed:\n' + res.stdout + res.stderr)
    return lib


if __name__ == '__main__':
    import sys
    if '--alt' in sys.argv:
        i = sys.argv.index('--alt')
        print(build_alt(sys.argv[i + 2:], name=sys.argv[i + 1], verbose='-v' in sys.argv))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
