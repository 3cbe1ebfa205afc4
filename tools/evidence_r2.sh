#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun): the default bench line, the ncu launch list of a
# short run of the same command, and `ncu --set full` captures of the dominant kernel at the
# bench's batch size (c5), of c2 and of the likelihood kernel.
# Usage: bash tools/evidence_r2.sh <tag>        -> gpurun_out/<tag>_*
set -u
tag=${1:-r2}
out=gpurun_out
cd "$(dirname "$0")/.."
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err
# launch list: the plain run first, then the same command under ncu
short="--steps 3 --warmup 3 --no-cpu-baseline --no-plan --no-workloads"
python bench.py $short > $out/${tag}_bench_short.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py $short > $out/${tag}_ncu_launch.log 2>&1
# full captures (a launch after warm-up)
ncu --set full --clock-control none --import-source on -k regex:rollout_lane -s 3 -c 1 -f -o $out/prof_${tag}_c5 \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan --no-workloads > $out/${tag}_ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_lane -s 3 -c 1 -f -o $out/prof_${tag}_c2 \
  python bench.py --workload c2_merge_sbg --rollouts 1048576 --steps 2 --warmup 3 --no-cpu-baseline --no-plan --no-workloads > $out/${tag}_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:likelihood -s 5 -c 1 -f -o $out/prof_${tag}_likelihood \
  python bench.py --rollouts 65536 --steps 1 --warmup 3 --no-cpu-baseline --no-plan > $out/${tag}_ncu_lik.log 2>&1
ls -la $out | grep ${tag}
