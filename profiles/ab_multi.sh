#!/bin/bash
# Same-box comparison of several (library, TPR_WS_VARIANT) pairs: bash profiles/ab_multi.sh <tag> lib:variant ...   (lib = main or an alt name)
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for lv in "$@"; do
    lib=${lv%%:*}; v=${lv##*:}
    if [ "$lib" = main ]; then unset TPR_LIB; else export TPR_LIB=$PWD/g-nerf_b200/lib/libtriplane_b200_$lib.so; fi
    for mode in fp32 bf16; do
      TPR_WS_VARIANT=$v python bench.py --steps 60 --warmup 5 --legs none --mode $mode > gpurun_out/${tag}_${lib}_v${v}_${mode}_$rep.json 2>> gpurun_out/${tag}.err
    done
  done
done
unset TPR_LIB
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${tag}_*_v*_*.json')):
    try:
        d = json.load(open(f)); print(f.split('/')[-1], 'ms/step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))
    except Exception as e: print(f, 'ERR', e)
PY
