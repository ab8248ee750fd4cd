set -x
timeout 900 python -m pytest tests/test_gpu_backward.py -x -q -s 2>&1 | tail -12
timeout 600 python profiles/backward_timing.py --eager-img 0 > gpurun_out/s13_bwd.json 2> gpurun_out/s13_bwd.err; cat gpurun_out/s13_bwd.json; tail -5 gpurun_out/s13_bwd.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s13_bwd_launches.csv python profiles/backward_timing.py --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s13_ncu.err; tail -3 gpurun_out/s13_ncu.err
python profiles/launch_summary.py gpurun_out/s13_bwd_launches.csv > gpurun_out/s13_launch_summary.txt; head -6 gpurun_out/s13_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_backward -c 1 -o gpurun_out/s13_decode_bwd python profiles/backward_timing.py --n-img 2 --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s13_ncu2.err; tail -3 gpurun_out/s13_ncu2.err
