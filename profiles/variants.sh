#!/bin/bash
# Same-box comparison of TPR_WS_VARIANT settings with one library: bash profiles/variants.sh <tag> <lib-name|main> <variant>...
tag=$1; lib=$2; shift 2
mkdir -p gpurun_out
if [ "$lib" != main ]; then export TPR_LIB=$PWD/g-nerf_b200/lib/libtriplane_b200_$lib.so; fi
for rep in 1 2; do
  for v in "$@"; do
    for mode in fp32 bf16; do
      TPR_WS_VARIANT=$v python bench.py --steps 60 --warmup 5 --legs none --mode $mode > gpurun_out/${tag}_v${v}_${mode}_$rep.json 2>> gpurun_out/${tag}.err
    done
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${tag}_v*_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'ms/step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))
    except Exception as e: print(f, 'ERR', e)
PY
