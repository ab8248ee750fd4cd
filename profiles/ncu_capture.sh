#!/bin/bash
# bash profiles/ncu_capture.sh <tag> [mode] [kernel-regex] [extra bench args]: one `ncu --set full` capture (source on) of the render
# kernel at config 2; the summary and the per-source-line table are extracted on the box (the .ncu-rep is kept only if small).
tag=$1; mode=${2:-fp32}; kre=${3:-render_ws}
mkdir -p gpurun_out
rep=gpurun_out/${tag}_${kre}_${mode}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -o $rep \
  python bench.py --mode $mode --steps 1 --warmup 3 --legs none > gpurun_out/${tag}_ncu_${mode}.log 2>&1
python profiles/ncu_summary.py $rep.ncu-rep > ${rep}_ncu_full_summary.txt 2>&1
python profiles/ncu_lines.py $rep.ncu-rep 70 > ${rep}_ncu_lines.txt 2>&1
ncu -i $rep.ncu-rep --page raw --csv > ${rep}_raw.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --print-source sass --csv > ${rep}_sass.csv 2>/dev/null
gzip -f ${rep}_sass.csv
rm -f $rep.ncu-rep
ls -la gpurun_out/
