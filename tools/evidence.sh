#!/bin/bash
# Round evidence on one B200 (run under gpurun): plain bench lines for every workload, the ncu
# launch list of a short run, and one `ncu --set full` capture per kernel family.
# Usage: bash tools/evidence.sh <tag>        -> gpurun_out/<tag>_*
set -u
tag=${1:-r1}
out=gpurun_out
cd "$(dirname "$0")/.."
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
for w in c1_highway_mdp c3_merge_risk c4_coop c5_rsbg; do
  python bench.py --workload $w --rollouts 524288 --steps 10 --warmup 3 > $out/${tag}_bench_$w.json 2>> $out/${tag}_bench.err
done
# launch list: the plain run first, then the same command under ncu
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-plan > $out/${tag}_bench_short.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-plan > $out/${tag}_ncu_launch.log 2>&1
# full captures (second launch of the kernel: after warm-up)
ncu --set full --clock-control none --import-source on -k regex:rollout_lane -s 1 -c 1 -f -o $out/prof_${tag}_c2 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan > $out/${tag}_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_lane -s 1 -c 1 -f -o $out/prof_${tag}_c5 \
  python bench.py --workload c5_rsbg --rollouts 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-plan > $out/${tag}_ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_lane -s 1 -c 1 -f -o $out/prof_${tag}_c4 \
  python bench.py --workload c4_coop --rollouts 262144 --steps 2 --warmup 3 --no-cpu-baseline --no-plan > $out/${tag}_ncu_c4.log 2>&1
ls -la $out | grep ${tag}
