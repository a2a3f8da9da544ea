#!/usr/bin/env python
"""Turns `ncu -i X.ncu-rep --page raw --csv` dumps (gpurun_out/r2e_ncu_<workload>_raw.csv) into the
markdown table committed as profiles/r2_ncu_summary.md and refreshes profiles/ncu_traffic.json
(DRAM bytes per launch of every format's kernel, which bench.py reports as roofline.traffic)."""
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
COLS = {
    "ms": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
    "dram_pct": "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit": "lts__t_sector_hit_rate.pct", "l1_hit": "l1tex__t_sector_hit_rate.pct",
    "lts_sectors": "lts__t_sectors.sum", "l1_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "regs": "launch__registers_per_thread", "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "stall_lsb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "inst": "smsp__inst_executed.sum", "grid": "launch__grid_size",
}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return float("nan")


def short(name):
    name = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()
    return name.replace(" ", "")


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": short(d["Kernel Name"])}
        for k, col in COLS.items():
            v = num(d.get(col, "nan"))
            u = units[hdr.index(col)] if col in hdr else ""
            if k == "ms":
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}.get(u, 1e-6)
            if k in ("dram_rd", "dram_wr"):
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            rec[k] = v
        out.append(rec)
    return out


def main():
    alg = json.loads(Path(sys.argv[1]).read_text()) if len(sys.argv) > 1 else {}
    lines = ["# Round 2 -- ncu `--set full --clock-control none` of every SpMV kernel, library defaults",
             "",
             "One capture per workload (`opencl-spmv-algorithms_b200/tools/ncu_target.py`, script "
             "`profiles/r2_scripts/gpu_call5.sh`): each format's SpMV is launched twice, one kernel at a time (a sync "
             "between launches), and ncu profiles every launch with its default cache control (caches flushed before each "
             "replay).  The SECOND launch of each kernel is listed.  Times under ncu are cold-cache and serialised: the "
             "byte counters, hit rates and occupancy are the evidence, the un-profiled times are in the bench files.",
             "`DRAM B` = dram__bytes_read.sum + dram__bytes_write.sum per launch; `alg B` = algorithmic bytes of the "
             "format (SURVEY 8d); `L2 hit` = lts__t_sector_hit_rate; `L1 ld sectors` = l1tex__t_sectors_pipe_lsu_mem_global_op_ld; "
             "`stall LSB` = warps stalled on long_scoreboard per issue.", ""]
    traffic = []
    for w, dt, rows_n in (("rmat", "f32", 1 << 24), ("banded", "f32", 2097152), ("cant", "f64", 62451), ("laplace", "f64", 8000000)):
        path = ROOT / "gpurun_out" / f"r2e_ncu_{w}_raw.csv"
        if not path.exists():
            continue
        recs = load(path)
        seen, keep = {}, []
        for r in recs:
            seen[r["kernel"]] = seen.get(r["kernel"], 0) + 1
        count = {}
        for r in recs:   # the last launch of every distinct kernel
            count[r["kernel"]] = count.get(r["kernel"], 0) + 1
            if count[r["kernel"]] == seen[r["kernel"]]:
                keep.append(r)
        lines += [f"## {w} ({dt})", "",
                  "| kernel | ncu ms | DRAM B | alg B | DRAM/alg | DRAM GB/s (under ncu) | L2 thr % | L1 thr % | L2 hit % | L1 hit % | L1 ld sectors | regs | warps % | stall LSB | grid |",
                  "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
        tb = {}
        for r in keep:
            dram = r["dram_rd"] + r["dram_wr"]
            a = None
            for key, val in alg.get(w, {}).items():
                if r["kernel"].startswith(key):
                    a = val
            lines.append(f"| `{r['kernel'][:70]}` | {r['ms']:.4f} | {dram:.4g} | {a if a else ''} | {dram / a:.3f} |" if a else
                         f"| `{r['kernel'][:70]}` | {r['ms']:.4f} | {dram:.4g} |  |  |")
            lines[-1] += (f" {dram / (r['ms'] * 1e-3) * 1e-9:.0f} | {r['lts_pct']:.1f} | {r['l1_pct']:.1f} | {r['l2_hit']:.1f} | {r['l1_hit']:.1f} | "
                          f"{r['l1_ld_sectors']:.4g} | {r['regs']:.0f} | {r['warps_pct']:.1f} | {r['stall_lsb']:.1f} | {r['grid']:.0f} |")
            tb[r["kernel"]] = int(dram)
        lines.append("")
        traffic.append({"workload": {"laplace": "laplace-iter"}.get(w, w), "dtype": dt, "rows": rows_n, "kernels": tb})
    Path(ROOT / "profiles" / "r2_ncu_summary.md").write_text("\n".join(lines) + "\n")
    Path(ROOT / "profiles" / "r2_ncu_kernels.json").write_text(json.dumps(traffic, indent=1))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
