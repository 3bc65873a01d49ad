"""Condense `ncu -i x.ncu-rep --page raw --csv` into a small JSON + text summary per kernel launch.
Usage: python tools/ncu_summary.py gpurun_out/r01_ncu_gemm_conv320.raw.csv profiles/r01_ncu_gemm_conv320"""
import csv
import json
import re
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__occupancy_limit_shared_mem",
]
EXTRA_PAT = re.compile(r"(pipe_tensor|tmem|utc|tma|pipe_xu_realtime)", re.I)


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]), "grid": r[hdr.index("Grid Size")],
             "block": r[hdr.index("Block Size")], "metrics": {}}
        for i, h in enumerate(hdr):
            if h in KEEP or (EXTRA_PAT.search(h) and r[i] not in ("", "0", "n/a")):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d["metrics"][h] = {"value": v, "unit": units[i]}
        m = d["metrics"]

        def g(k, scale=1.0):
            if k not in m:
                return None
            u = m[k]["unit"].lower()
            mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            return m[k]["value"] * mult * scale

        rd, wr = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
        d["dram_bytes"] = (rd or 0) + (wr or 0)
        d["l2_to_sm_bytes"] = g("l1tex__m_xbar2l1tex_read_bytes.sum")
        out.append(d)
    json.dump(out, open(dst + ".json", "w"), indent=1)
    with open(dst + ".txt", "w") as f:
        for d in out:
            m = d["metrics"]
            f.write(f"{d['kernel']} grid {d['grid']} block {d['block']}\n")
            for k in sorted(m):
                f.write(f"  {k:90s} {m[k]['value']:>16.3f} {m[k]['unit']}\n")
            f.write(f"  => dram bytes {d['dram_bytes']:.0f}  L2->SM bytes {d['l2_to_sm_bytes']}\n\n")
    print("wrote", dst + ".json", dst + ".txt", len(out), "launches")


if __name__ == "__main__":
    main()
