#!/usr/bin/env python
"""Per-source-line instruction / stall summary of an ncu report.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_hot_lines.py src.csv [top_n]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    cur, hdr, lines = None, None, []
    for r in csv.reader(open(path)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif len(r) > 5 and hdr and r[2] == "-" and r[0].isdigit():
            lines.append((cur, r))
    ix = {n: i for i, n in enumerate(hdr)}
    I, T, S = ix["Instructions Executed"], ix["Thread Instructions Executed"], ix["# Samples"]
    stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot_i = sum(int(r[I]) for _, r in lines)
    tot_t = sum(int(r[T]) for _, r in lines)
    tot_s = sum(int(r[S]) for _, r in lines)
    print(f"total warp inst {tot_i}  thread inst {tot_t}  samples {tot_s}")
    agg = {}
    for n in stalls:
        agg[n] = sum(int(r[ix[n]]) for _, r in lines)
    print("stall samples:", ", ".join(f"{n[6:]} {100 * v / tot_s:.1f}%" for n, v in
                                       sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    lines.sort(key=lambda fr: -int(fr[1][I]))
    for f, r in lines[:top]:
        top_stall = max(stalls, key=lambda n: int(r[ix[n]]))
        print(f"{100 * int(r[I]) / tot_i:5.2f}% inst {100 * int(r[S]) / tot_s:5.2f}% smp "
              f"thr/inst {int(r[T]) / max(1, int(r[I])):4.1f} {top_stall[6:]:>10s}  {f}:{r[0]:>4s}  {r[1].strip()[:90]}")


if __name__ == "__main__":
    main()
