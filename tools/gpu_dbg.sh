#!/bin/bash
mkdir -p gpurun_out
for w in eval train_fwd train_fwd_bwd train_dp train_full; do timeout 300 python tools/graph_debug.py $w global 2>&1 | grep -E "capture|FAILED|Error" | head -3; done
for w in train_fwd_bwd train_full; do timeout 300 python tools/graph_debug.py $w thread_local 2>&1 | grep -E "capture|FAILED|Error" | head -3; done
