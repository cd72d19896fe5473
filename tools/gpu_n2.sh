#!/bin/bash
mkdir -p gpurun_out
for w in none flat overlap; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/graph_debug_n2.py $w 2>&1 | grep -E "capture OK|FAILED" ; done
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -s > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; grep -E "top offenders|passed|failed" gpurun_out/pytest_multi.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; grep -E "bench\]" gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_n2.json')); print('N2 value', j['value'], 'e2e', j['e2e']['value'], 'train', j['train_step'])
except Exception as e: print('parse fail', e)
PY
