#!/bin/bash
mkdir -p gpurun_out
SUNB_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "tcgen05" > gpurun_out/pytest_2cta_k.log 2>&1; echo "2cta kernels exit $?"; tail -4 gpurun_out/pytest_2cta_k.log | cut -c1-300
SUNB_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_train.py -q > gpurun_out/pytest_2cta_e.log 2>&1; echo "2cta enc/train exit $?"; tail -4 gpurun_out/pytest_2cta_e.log | cut -c1-300
SUNB_GEMM_2CTA=1 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_2cta.json 2> gpurun_out/bench_2cta.err; echo "bench 2cta exit $?"
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_2cta.json')); print('2CTA value', j['value'], 'e2e', j['e2e']['value'], 'roof', j['roofline']['frac'], 'train', j['train_step'].get('ms_per_step'), j['train_step'].get('launch_mode'))
except Exception as e: print('parse fail', e)
PY
SUNB_GEMM_2CTA=1 SUNB_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_2cta.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_run_2cta.log 2>&1; echo "ncu list exit $?"
