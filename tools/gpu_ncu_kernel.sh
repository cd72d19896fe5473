#!/bin/bash
# full ncu capture of selected kernels of one eval forward: usage gpu_ncu_kernel.sh <regex> <skip> <count> <outname>
mkdir -p gpurun_out
SUNB_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" -f -o "gpurun_out/$4" python bench.py --steps 1 --warmup 1 > "gpurun_out/$4.log" 2>&1; echo "ncu $4 exit $?"
