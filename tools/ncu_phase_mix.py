#!/usr/bin/env python
"""Dynamic instruction share per kernel phase.

usage: python tools/ncu_phase_mix.py src_cuda_sass.csv kernel_file.cuh "name:first-last" ...
SASS is laid out roughly in program order; an instruction inlined from a helper header is
attributed to the last line of `kernel_file` seen before it in address order."""
import csv
import sys


def main():
    path, kfile = sys.argv[1], sys.argv[2]
    phases = []
    for spec in sys.argv[3:]:
        name, rng = spec.split(":")
        a, b = rng.split("-")
        phases.append((name, int(a), int(b)))
    cur, hdr, line = None, None, None
    insts = []  # (address, file, line, executed, thread_executed, samples)
    for r in csv.reader(open(path)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
            ix = {n: i for i, n in enumerate(hdr)}
        elif len(r) > 5 and hdr:
            if r[2] == "-":
                line = int(r[0]) if r[0].isdigit() else None
            elif r[2].startswith("0x"):
                insts.append((int(r[2], 16), cur, line, int(r[ix["Instructions Executed"]]),
                              int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])))
    insts.sort()
    tot = sum(i[3] for i in insts)
    tots = sum(i[5] for i in insts)
    agg = {}
    last = None
    for addr, f, ln, ex, th, sm in insts:
        if f == kfile and ln:
            last = ln
        name = "other"
        if last:
            for n, a, b in phases:
                if a <= last <= b:
                    name = n
                    break
        e = agg.setdefault(name, [0, 0, 0, 0])
        e[0] += ex
        e[1] += th
        e[2] += sm
        e[3] += 1
    for n, (ex, th, sm, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{n:14s} {100 * ex / tot:5.1f}% inst {100 * sm / tots:5.1f}% smp  thr/inst {th / max(ex, 1):4.1f}  static {cnt}")


if __name__ == "__main__":
    main()
