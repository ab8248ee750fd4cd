"""Build libtriplane_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension:
the library is a plain C-ABI shared object, see include/triplane_b200.h)."""
import hashlib
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libtriplane_b200.so')
SOURCES = ['triplane_b200.cu', 'tpr_tc_debug.cu', 'tpr_render_tc.cu', 'tpr_render_ws.cu', 'tpr_microbench.cu']
HEADERS = [os.path.join(CSRC, 'tpr_device.cuh'), os.path.join(CSRC, 'tpr_render.cuh'), os.path.join(CSRC, 'tpr_tc.cuh'), os.path.join(ROOT, 'include', 'triplane_b200.h')]

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '--threads', '0']


def _digest():
    h = hashlib.sha256()
    for f in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(f, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the sources changed since the last build; return the library path."""
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, 'libtriplane_b200.sha256')
    dig = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC]
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(stamp, 'w') as fh:
        fh.write(dig)
    return LIB_PATH


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
