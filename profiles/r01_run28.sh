set -x
for i in 1 2; do
timeout 900 python profiles/extra_configs.py --n-img 8 --skip5 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('merge   ', d['config4']['ms'], d['config3_render_frames']['ms_per_orbit'])"
TPR_WS_VARIANT=3 timeout 900 python profiles/extra_configs.py --n-img 8 --skip5 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('variant3', d['config4']['ms'], d['config3_render_frames']['ms_per_orbit'])"
done
TPR_PT_DEPTH=96 timeout 300 python profiles/phase_timing.py fp32 2>&1 | tail -20
