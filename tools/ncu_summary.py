"""Summaries of ncu artefacts for profiles/ (run in the build container on files brought back in gpurun_out/).
  python tools/ncu_summary.py shares  gpurun_out/launches.csv          -> share table of a launch list
  python tools/ncu_summary.py kernel  gpurun_out/x.ncu-rep             -> key metrics of a --set full capture"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]


def shares(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    col = {h: i for i, h in enumerate(rows[hi])}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 2:]:
        if len(r) < len(col):
            continue
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[col["Metric Unit"]]
        us = v / 1000 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1000)
        agg[r[col["Kernel Name"]]][0] += 1
        agg[r[col["Kernel Name"]]][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"total kernel time {tot:.1f} us over {sum(v[0] for v in agg.values())} launches "
          "(per-launch times are cold-cache / serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{100 * v[1] / tot:6.2f}% {v[1]:10.1f} us {v[0]:5d} x  {k[:120]}")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    {"shares": shares, "kernel": kernel}[sys.argv[1]](sys.argv[2])
