#!/bin/bash
# timing experiments of the fused stage-1 block tail (compile-time variants: build them with `tools/build_variants.sh tail`)
python -m pytest tests/test_gpu_kernels.py tests/test_input_path.py -m gpu -q -x -p no:cacheprovider -k "convmlp or preprocess or device_store" 2>&1 | tail -3
python tools/convmlp_time.py
for v in 1 2 4 8 63; do echo -n "dbg $v: "; SUNB200_LIB=$PWD/tools/_dbg/libsunb_dbg$v.so python tools/convmlp_time.py | sed 's/.*fused/fused/'; done
