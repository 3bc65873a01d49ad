"""One CUDA-graph replay (between the last two ddpm_step kernels) of an ncu launch list -> compact CSV + summary.
Usage: python tools/step_launches.py gpurun_out/launches_r01.csv profiles/r01_launches_unet_step_final"""
import collections
import csv
import re
import sys


def main():
    path, out = sys.argv[1], sys.argv[2]
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name).replace("tsd::<unnamed>::", "").replace("tsd::", "").replace("<unnamed>::", "")
        name = name.replace("unnamed>::", "")
        rows.append((int(r["ID"]), name, r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
    idx = [i for i, r in enumerate(rows) if "ddpm_step" in r[1]]
    step = rows[idx[-2] + 1:idx[-1] + 1]
    with open(out + ".csv", "w") as f:
        f.write("id,kernel,grid,block,duration_us\n")
        for r in step:
            f.write(f'{r[0]},{r[1]},"{r[2]}","{r[3]}",{r[4]:.3f}\n')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        agg[r[1]][0] += 1
        agg[r[1]][1] += r[4]
    tot = sum(r[4] for r in step)
    with open(out + "_summary.txt", "w") as f:
        f.write(f"{len(step)} launches, {tot:.1f} us total (ncu gpu__time_duration, serialised, one CUDA-graph replay of a "
                "UNet step + sampler step)\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{c:<4d} avg {t / c:8.2f} us  {n}\n")
    print(open(out + "_summary.txt").read())


if __name__ == "__main__":
    main()
