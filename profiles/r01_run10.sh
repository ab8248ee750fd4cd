set -x
timeout 600 python profiles/backward_timing.py > gpurun_out/s10_bwd.json 2> gpurun_out/s10_bwd.err; cat gpurun_out/s10_bwd.json; tail -5 gpurun_out/s10_bwd.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s10_bwd_launches.csv python profiles/backward_timing.py --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s10_ncu.err; tail -3 gpurun_out/s10_ncu.err
python profiles/launch_summary.py gpurun_out/s10_bwd_launches.csv | tail -20
