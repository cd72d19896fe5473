#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; grep -E "bench\]|Error" gpurun_out/bench.err | head -5 | cut -c1-300
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench.json')); print('value', j['value'], 'e2e', j['e2e']['value'], 'roof', j['roofline']['frac'], 'train', j['train_step'])
except Exception as e: print('parse fail', e)
PY
timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_train.py -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
