#!/bin/bash
# full round-end style run: GPU tests, bench, launch lists, full ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; grep -E "bench\]|Error" gpurun_out/bench.err | head -5 | cut -c1-300
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench.json')); print('value', j['value'], 'e2e', j['e2e']['value'], 'roof', j['roofline']['frac'], 'lat', j['single_episode_latency'], 'train', j['train_step'].get('ms_per_step'), j['train_step'].get('launch_mode'))
except Exception as e: print('parse fail', e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref exit $?"; cut -c1-200 gpurun_out/bench_ref.json
SUNB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_run.log 2>&1; echo "ncu list exit $?"
SUNB_BENCH_PROFILE=train SUNB_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_train_run.log 2>&1; echo "ncu train list exit $?"
SUNB_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_slab -s 3 -c 1 -f -o gpurun_out/prof_conv3 python bench.py --steps 1 --warmup 1 > gpurun_out/prof_full.log 2>&1; echo "ncu full exit $?"
