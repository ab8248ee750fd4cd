set -x
for d in 0 1 2 3; do
TPR_BWD_DEBUG=$d timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_backward -c 2 --csv python profiles/backward_timing.py --eager-img 0 --reps 1 2>/dev/null | grep decode_backward | tail -1
done
