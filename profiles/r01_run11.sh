set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_backward -c 1 -o gpurun_out/s11_decode_bwd python profiles/backward_timing.py --n-img 2 --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s11_ncu.err; tail -3 gpurun_out/s11_ncu.err
ls -la gpurun_out/*.ncu-rep
