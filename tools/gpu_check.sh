#!/bin/bash
# Runs the GPU test tiers in separate processes (a trapped kernel poisons its CUDA context) and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 25 gpurun_out/$name.log; }
SUNB_GEMM=simt run enc_simt python -m pytest tests/test_gpu_encoder.py -q -rA -s
run k_simt python -m pytest tests/test_gpu_kernels.py -q -rA -k "simt"
run k_other python -m pytest tests/test_gpu_kernels.py -q -rA -k "not simt and not tcgen05"
run k_tc python -m pytest tests/test_gpu_kernels.py -q -rA -k "tcgen05"
run enc_tc python -m pytest tests/test_gpu_encoder.py -q -rA -s
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 3 --warmup 3
