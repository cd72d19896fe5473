#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -rA -s > gpurun_out/pytest_train.log 2>&1; echo "pytest train exit $?"
grep -E "PASSED|FAILED|ERROR|rel |loss|median|passed|failed|Error|error:" gpurun_out/pytest_train.log | head -80
