#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; grep -E "bench\]" gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/bench_n2.json')); print('N2 value', j['value'], 'e2e', j['e2e']['value'], 'train', j['train_step'])
except Exception as e: print('parse fail', e)
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2>/dev/null; echo "ref arm exit $?"; cut -c1-300 gpurun_out/bench_ref_n2.json
