#!/bin/bash
# 2xFP16 vs 3xTF32 on one box with one library: bash profiles/ab_scheme.sh <tag> <lib|main>
tag=$1; lib=$2
mkdir -p gpurun_out
if [ "$lib" != main ]; then export TPR_LIB=$PWD/g-nerf_b200/lib/libtriplane_b200_$lib.so; fi
for rep in 1 2; do
  for scheme in f16x2 tf32x3; do
    TPR_FP32_SCHEME=$scheme python bench.py --steps 60 --warmup 5 --legs none --mode fp32 > gpurun_out/${tag}_${scheme}_$rep.json 2>> gpurun_out/${tag}.err
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${tag}_*_?.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'ms/step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))
    except Exception as e: print(f, 'ERR', e)
PY
