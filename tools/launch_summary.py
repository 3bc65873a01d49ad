"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
Usage: python tools/launch_summary.py gpurun_out/launches.csv [first_id [last_id]]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if lo <= i < hi:
            name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("tsd::<unnamed>::", "").replace("tsd::", "")
            rows.append((i, name, float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
    tot = sum(r[2] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for _, n, t, _, _ in rows:
        agg[n][0] += 1
        agg[n][1] += t
    print(f"{len(rows)} launches, {tot:.1f} us total")
    for n, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{cnt:<5d} avg {t / cnt:8.2f} us  {n}")


if __name__ == "__main__":
    main()
