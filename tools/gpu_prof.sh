#!/bin/bash
# GPU tests + bench + per-launch device times (ncu launch list) + one full ncu capture of the stem conv3 GEMM
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rA -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "rel_l2|passed|failed|FAILED|argmax" gpurun_out/pytest_gpu.log | tail -50
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
SUNB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_run.log 2>&1; echo "ncu list exit $?"
SUNB_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_conv3 python bench.py --steps 1 --warmup 1 > gpurun_out/prof_full.log 2>&1; echo "ncu full exit $?"
