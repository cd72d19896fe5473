#!/bin/bash
# bench.py on N GPUs of one box (usage: gpurun --gpus N -- bash tools/gpu_nN.sh N), plus the two-rank GPU tests over NCCL
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python - <<PY
import json
try:
    j = json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('N$N value', j['value'], 'e2e', j['e2e']['value'], 'e2e_fp32', j.get('e2e_fp32_input', {}).get('value'))
    print('train', {k: j['train_step'][k] for k in ('ms_per_step', 'launch_mode', 'ms_per_step_eager', 'grad_allreduce')})
    print('sun', {k: j['sun_meta_training_step'][k] for k in ('ms_per_step', 'launch_mode') if k in j['sun_meta_training_step']})
except Exception as e: print('parse fail', e)
PY
