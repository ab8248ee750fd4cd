"""Build the C oracle (oracle/triplane_oracle.c) into oracle/_build/ (git-ignored, travels with gpurun).
Checker / CPU-baseline only -- see the header of triplane_oracle.c."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'triplane_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libtriplane_oracle.so')
# x86-64-v3 (AVX2+FMA units for the vectoriser) but contraction OFF: products and sums round separately,
# exactly like the reference's float32 ops.
FLAGS = ['-O3', '-march=x86-64-v3', '-ffp-contract=off', '-fno-math-errno', '-fopenmp', '-shared', '-fPIC']


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    res = subprocess.run(['gcc'] + FLAGS + [SRC, '-o', LIB, '-lm'], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('gcc failed:\n' + res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
