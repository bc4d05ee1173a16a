#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv [steps] > profiles/rNN_launches.txt

Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes.
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d["Metric Unit"]
            v_us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
            name = d["Kernel Name"]
            ours = ("triad" in name) or ("maxmean" in name) or ("dq_" in name) or ("dv_" in name) or ("nce_" in name)
            a = agg.setdefault((ours, name.split("(")[0][:70]), [0, 0.0])
            a[0] += 1
            a[1] += v_us
    tot_ours = sum(t for (o, _), (_, t) in agg.items() if o)
    tot_all = sum(t for _, (_, t) in agg.items())
    print(f"# {path}: {steps} step(s); library kernels {tot_ours / steps:.1f} us/step, everything {tot_all / steps:.1f} us/step")
    print(f"{'kernel':72s} {'n':>5s} {'us/launch':>11s} {'us/step':>10s} {'share':>7s}")
    for (ours, name), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        tag = "" if ours else "  [torch]"
        print(f"{name + tag:72s} {n:5d} {t / n:11.1f} {t / steps:10.1f} {100 * t / tot_all:6.1f}%")


if __name__ == "__main__":
    main()
