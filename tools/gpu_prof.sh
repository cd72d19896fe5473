#!/bin/bash
# bench + per-launch device times (ncu launch list) for the eval step
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
SUNB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_run.log 2>&1; echo "ncu exit $?"
timeout 300 python -m pytest tests/test_gpu_encoder.py -q -rA -s > gpurun_out/enc_tc.log 2>&1; echo "enc exit $?"; grep -E "rel_l2|passed|failed" gpurun_out/enc_tc.log | tail -40
