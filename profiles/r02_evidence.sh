#!/bin/bash
# Round-2 evidence, one box: ncu --set full summaries of the final kernels, launch lists, phase tables, microbenchmarks.
# Everything lands in gpurun_out/ (copied into profiles/ afterwards).  Usage: bash profiles/r02_evidence.sh
mkdir -p gpurun_out
bash profiles/ncu_capture.sh r02 fp32 render_ws
bash profiles/ncu_capture.sh r02 bf16 render_ws
bash profiles/ncu_capture.sh r02_96x96 fp32 render_ws profiles/render_once.py 8 256 96 96 fp32 3
bash profiles/ncu_capture.sh r02 fp32 decode_backward_tc profiles/backward_timing.py --eager-img 0 --reps 1
bash profiles/ncu_capture.sh r02 fp32 march_backward profiles/backward_timing.py --eager-img 0 --reps 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_fp32.csv python bench.py --steps 2 --warmup 3 --legs none > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r02_launches_bench_fp32.csv > gpurun_out/r02_launches_bench_fp32_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_step.csv python profiles/backward_timing.py --eager-img 0 --reps 1 > /dev/null 2>&1
python profiles/launch_summary.py gpurun_out/r02_launches_train_step.csv > gpurun_out/r02_launches_train_step_summary.txt
timeout 120 python profiles/phase_timing.py fp32 bf16 > gpurun_out/r02_phase48.txt 2>&1
TPR_PT_DEPTH=96 timeout 120 python profiles/phase_timing.py fp32 > gpurun_out/r02_phase96.txt 2>&1
TPR_BWD_PROFILE=1 timeout 120 python profiles/backward_timing.py --eager-img 0 --reps 1 2>&1 | grep "bwd profile" | tail -17 > gpurun_out/r02_phase_backward.txt
timeout 120 python profiles/scatter_shapes.py > gpurun_out/r02_scatter_shapes.json 2>/dev/null
timeout 200 python profiles/backward_timing.py --eager-img 1 > gpurun_out/r02_backward_timing.json 2>/dev/null
TPR_BWD_IMPL=hmma timeout 200 python profiles/backward_timing.py --eager-img 0 > gpurun_out/r02_backward_timing_hmma.json 2>/dev/null
timeout 200 python profiles/backward_timing.py --eager-img 0 --mode bf16 > gpurun_out/r02_backward_timing_bf16.json 2>/dev/null
timeout 100 python profiles/render_once.py 8 256 96 96 fp32 5 > gpurun_out/r02_render_96x96.txt 2>&1
timeout 100 python profiles/render_once.py 8 256 96 96 bf16 5 >> gpurun_out/r02_render_96x96.txt 2>&1
timeout 100 python profiles/render_once.py 8 128 48 48 fp32 20 >> gpurun_out/r02_render_96x96.txt 2>&1
timeout 100 python profiles/render_once.py 8 128 48 48 bf16 20 >> gpurun_out/r02_render_96x96.txt 2>&1
ls -la gpurun_out | grep r02_ | tail -40
