#!/bin/bash
# 1 -> 8 GPU scaling series on one box (run under `gpurun --gpus 8`), the way the driver launches it
out=gpurun_out
tag=${1:-r2_scale}
cd "$(dirname "$0")/.."
nproc; nvidia-smi -L | wc -l
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_1gpu.json 2> $out/${tag}_1gpu.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_${n}gpu.json 2> $out/${tag}_${n}gpu.err
done
TAG=$tag python - <<'P'
import json, os
tag = os.environ["TAG"]
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/{tag}_{n}gpu.json").read().strip().split("\n")[-1])
    except Exception as e:
        print(n, "failed", e)
        continue
    if base is None:
        base = d
    p = d.get("plan") or {}
    print(f"N={n}: value {d['value']:.4g} (eff {d['value'] / (n * base['value']):.3f})  e2e {d['e2e']['value']:.4g} "
          f"(eff {d['e2e']['value'] / (n * base['e2e']['value']):.3f})  plan {p.get('iterations_per_s', 0):.4g} it/s "
          f"(eff {p.get('iterations_per_s', 0) / (n * base['plan']['iterations_per_s']):.3f}, "
          f"{p.get('host_threads_per_rank')} threads/rank, merge {p.get('merge_ms', 0):.2f} ms)")
P
