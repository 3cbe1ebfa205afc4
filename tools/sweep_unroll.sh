#!/bin/bash
# Unroll factors of the three inner loops of the lane kernel (IDM sub-steps, ego sub-steps,
# segments of a projection); run on the GPU box.
for cfg in "2 1 1" "1 1 1" "5 1 1" "2 2 1" "2 1 2" "2 1 3"; do
  set -- $cfg
  BMCTS_NVCC_EXTRA="-DBMCTS_IDM_UNROLL=$1 -DBMCTS_EGO_UNROLL=$2 -DBMCTS_PROJECT_UNROLL=$3 -Xptxas -v" python planner-mcts_b200/build.py --force --verbose 2>&1 | grep -A1 "rollout_lane_kernelILi1E" | grep -E "spill" | head -1
  for wl in c2_merge_sbg c5_rsbg; do
    python bench.py --workload $wl --rollouts 524288 --steps 5 --warmup 3 --no-cpu-baseline --no-plan 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('idm $1 ego $2 project $3  $wl', '%.3e' % d['value'], 'ms %.3f' % d['roofline']['kernel_ms'])"
  done
done
python planner-mcts_b200/build.py --force > /dev/null
