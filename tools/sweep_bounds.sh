set -e
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "320 2" "256 3" "256 4" "384 2"; do
  set -- $cfg
  BMCTS_NVCC_EXTRA="-DBMCTS_LANE_MAX_THREADS=$1 -DBMCTS_LANE_MIN_BLOCKS=$2" python planner-mcts_b200/build.py --force > /dev/null
  echo "== bounds $1 x $2"
  BMCTS_DEBUG=1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rollouts 100000 2>&1 | grep "^\[bmcts\]" | tail -1
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'])"
done
