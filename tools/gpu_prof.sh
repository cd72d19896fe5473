#!/bin/bash
# round-end style run on the GPU box: GPU tests, smoke, bench (product + reference arm), launch lists, full ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref exit $?"
SUNB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_run.log 2>&1; echo "ncu list exit $?"
SUNB_BENCH_PROFILE=train SUNB_TRAIN_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 1 --warmup 1 > gpurun_out/prof_train_run.log 2>&1; echo "ncu train list exit $?"
SUNB_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_slab -s 3 -c 1 -f -o gpurun_out/prof_conv3 python bench.py --steps 1 --warmup 1 > gpurun_out/prof_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:convmlp_tail -s 2 -c 1 -f -o gpurun_out/prof_tail python tools/tail_one.py > /dev/null 2>&1; echo "ncu tail exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 1 -f -o gpurun_out/prof_att100 python tools/att_one.py 100 42 48 > /dev/null 2>&1; echo "ncu att exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_tc -s 2 -c 1 -f -o gpurun_out/prof_attb100 python tools/attb_time.py 480 > /dev/null 2>&1; echo "ncu attb exit $?"
