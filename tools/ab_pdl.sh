#!/bin/bash
# A/B of programmatic dependent launch: product library vs a -DSUNB_NO_PDL build (tools/build_variants.sh -> tools/_ab/libsunb200_nopdl.so)
for lib in "" "tools/_ab/libsunb200_nopdl.so"; do
  echo "=== lib: ${lib:-product}"
  export SUNB200_LIB=$lib; [ -z "$lib" ] && unset SUNB200_LIB
  SUNB_TRAIN_EPISODES=1 SUNB_BENCH_PROFILE=train timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['train_step']; print('train 60 img:', d['ms_per_step'], d['launch_mode'], 'eager', d['ms_per_step_eager'])"
  SUNB_BENCH_PROFILE=train timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['train_step']; print('train 480 img:', d['ms_per_step'], d['launch_mode'], 'eager', d['ms_per_step_eager'])"
  timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('eval:', d['value'], 'e2e', d['e2e']['value'], 'lat', d['single_episode_latency'], '1shot', d['eval_5way_1shot']['value'], 'sun', d['sun_meta_training_step']['ms_per_step'])"
done
