#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -4 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -4 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_n2.json')); print('N2 value', j['value'], 'e2e', j['e2e']['value'], 'train', j['train_step'])
except Exception as e: print('parse fail', e)
PY
SUNB_DDP_OVERLAP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_noov.json 2> gpurun_out/bench_n2_noov.err; echo "bench n2 no-overlap exit $?"
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_n2_noov.json')); print('N2 no-overlap train', j['train_step']['ms_per_step'], j['train_step']['launch_mode'])
except Exception as e: print('parse fail', e)
PY
