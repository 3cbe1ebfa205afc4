#!/bin/bash
# Launch-bounds sweep of the lane kernels (registers per thread vs resident warps); run on the GPU box.
for cfg in "384 2" "320 2" "256 3" "256 4" "512 1"; do
  set -- $cfg
  BMCTS_NVCC_EXTRA="-DBMCTS_LANE_MAX_THREADS=$1 -DBMCTS_LANE_MIN_BLOCKS=$2" python planner-mcts_b200/build.py --force > /dev/null
  for wl in c2_merge_sbg c5_rsbg c4_coop; do
    python bench.py --workload $wl --rollouts 524288 --steps 5 --warmup 3 --no-cpu-baseline --no-plan 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bounds $1 x $2  $wl', '%.3e' % d['value'], 'ms %.3f' % d['roofline']['kernel_ms'])"
  done
done
python planner-mcts_b200/build.py --force > /dev/null
