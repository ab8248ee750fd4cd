set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r3_pytest.log; tail -3 gpurun_out/r3_pytest.log
timeout 600 python bench.py > gpurun_out/r3_bench_fp32.json 2> gpurun_out/r3_bench_fp32.err; cat gpurun_out/r3_bench_fp32.json
timeout 600 python bench.py --mode bf16 --no-cpu-baseline > gpurun_out/r3_bench_bf16.json 2> gpurun_out/r3_bench_bf16.err; cat gpurun_out/r3_bench_bf16.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3_bench_ref.json 2>&1; cat gpurun_out/r3_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3_ncu_b.log 2>&1
timeout 900 python profiles/extra_configs.py > gpurun_out/r3_extra.json 2> gpurun_out/r3_extra.err; cat gpurun_out/r3_extra.json; tail -3 gpurun_out/r3_extra.err
python __graft_entry__.py smoke 2>&1 | tail -2
ls -la gpurun_out
