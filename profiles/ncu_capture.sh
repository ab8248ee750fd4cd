#!/bin/bash
# bash profiles/ncu_capture.sh <tag> <mode> <kernel-regex> <python script + args ...>: one `ncu --set full` capture (source on) of the
# first matching kernel launch after 3 skipped ones; the summary and the per-source-line table are extracted on the box (the .ncu-rep
# itself is too big to bring back).  Default script: bench.py at config 2.
tag=$1; mode=${2:-fp32}; kre=${3:-render_ws}; shift 3
if [ $# -eq 0 ]; then set -- bench.py --mode $mode --steps 1 --warmup 3 --legs none; fi
mkdir -p gpurun_out
rep=gpurun_out/${tag}_${kre}_${mode}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -o $rep python "$@" > gpurun_out/${tag}_ncu_${mode}.log 2>&1
python profiles/ncu_summary.py $rep.ncu-rep > ${rep}_ncu_full_summary.txt 2>&1
python profiles/ncu_lines.py $rep.ncu-rep 70 > ${rep}_ncu_lines.txt 2>&1
rm -f $rep.ncu-rep
