set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_backward -c 1 -o gpurun_out/s34_decode_bwd python profiles/backward_timing.py --n-img 2 --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s34_ncu1.err; tail -2 gpurun_out/s34_ncu1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_backward -c 1 -o gpurun_out/s34_march_bwd python profiles/backward_timing.py --n-img 2 --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s34_ncu2.err; tail -2 gpurun_out/s34_ncu2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ws_kernel -c 1 -s 8 -o gpurun_out/s34_render_train python profiles/backward_timing.py --n-img 8 --eager-img 0 --reps 1 > /dev/null 2> gpurun_out/s34_ncu3.err; tail -2 gpurun_out/s34_ncu3.err
ls -la gpurun_out/s34*.ncu-rep
