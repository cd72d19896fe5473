#!/bin/bash
# GPU tests + bench + per-launch device times + full ncu capture of the stage-2 GEMMs (pe2, qkv, proj, mlp1, mlp2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench.json')); print('value', j['value'], 'e2e', j['e2e']['value'], 'roof', j['roofline']['frac'], 'train', j['train_step']['ms_per_step'])
except Exception as e: print('parse fail', e)
PY
SUNB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_run.log 2>&1; echo "ncu list exit $?"
SUNB_BENCH_PROFILE=train timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_train_run.log 2>&1; echo "ncu train list exit $?"
SUNB_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 10 -c 5 -f -o gpurun_out/prof_s2 python bench.py --steps 1 --warmup 1 > gpurun_out/prof_full.log 2>&1; echo "ncu full exit $?"
