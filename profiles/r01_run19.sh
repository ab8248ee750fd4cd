set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/s19_bench_fp32.json 2> gpurun_out/s19_bench_fp32.err; cat gpurun_out/s19_bench_fp32.json; tail -3 gpurun_out/s19_bench_fp32.err
timeout 600 python bench.py --mode bf16 --no-cpu-baseline > gpurun_out/s19_bench_bf16.json 2> gpurun_out/s19_bench_bf16.err; cat gpurun_out/s19_bench_bf16.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s19_bench_ref.json 2>&1; cat gpurun_out/s19_bench_ref.json
TPR_PT_DEPTH=96 timeout 300 python profiles/phase_timing.py fp32 2>&1 | tail -21 | tee gpurun_out/s19_phase96.txt
