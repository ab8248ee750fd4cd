#!/bin/bash
# Same-box A/B of two builds of the library: bash profiles/ab.sh <tag> <alt-lib-name> ; writes gpurun_out/<tag>_*.json
tag=$1; alt=$2
mkdir -p gpurun_out
for rep in 1 2; do
  for which in main $alt; do
    if [ "$which" = main ]; then unset TPR_LIB; else export TPR_LIB=$PWD/g-nerf_b200/lib/libtriplane_b200_$which.so; fi
    for mode in fp32 bf16; do
      python bench.py --steps 60 --warmup 5 --legs none --mode $mode > gpurun_out/${tag}_${which}_${mode}_$rep.json 2>> gpurun_out/${tag}.err
    done
  done
done
unset TPR_LIB
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${tag}_*_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'ms/step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))
    except Exception as e: print(f, 'ERR', e)
PY
