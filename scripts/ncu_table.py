#!/usr/bin/env python
"""Reads an `ncu --page raw --csv` export and prints / writes the per-kernel table used in profiles/*_summary.md and the
counters file bench.py reads (profiles/<tag>_counters.json).

    ncu -i gpurun_out/r02_prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_table.py /tmp/raw.csv[,/tmp/more.csv] profiles/r02_counters.json zero123g/trained
"""
import csv
import json
import sys

WANT = [("time_ms", "gpu__time_duration.sum", 1e-6 * 1e3), ("regs", "launch__registers_per_thread", 1),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
        ("ipc_per_sm", "sm__inst_issued.avg.per_cycle_active", 1), ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("dram_read_bytes", "dram__bytes_read.sum", 1),
        ("dram_write_bytes", "dram__bytes_write.sum", 1), ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", 1),
        ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1), ("warp_instructions", "smsp__inst_executed.sum", 1),
        ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
        ("pipe_fma_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
        ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1),
        ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1),
        ("sm_cycles_active", "sm__cycles_active.avg", 1)]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    out = []
    for path in sys.argv[1].split(","):
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        parse_rows(rows[2:], idx, units, out)
    report(out)


def parse_rows(rows, idx, units, out):
    for r in rows:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("lgm::", "").replace("<unnamed>::", "").replace("unnamed>::", "").strip()
        d = {"kernel": short, "grid": r[idx["launch__grid_size"]], "block": r[idx["launch__block_size"]]}
        for key, metric, _ in WANT:
            if metric not in idx or r[idx[metric]] == "":
                continue
            v = float(r[idx[metric]].replace(",", ""))
            u = units[idx[metric]]
            if key == "time_ms":
                v = v * UNIT_SCALE.get(u, 1.0) * 1e3
            elif u in UNIT_SCALE and "bytes" in key:
                v = v * UNIT_SCALE[u]
            d[key] = v
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[i])
                  for h, i in idx.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i]}
        d["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
        out.append(d)


def report(out):
    print("| kernel | time | regs | warps active | issue active | DRAM % | DRAM read / write | L1/TEX % | warp instr | top stalls (warps per issue cycle) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for d in out:
        print("| %s | %.3f ms | %d | %.0f %% | %.0f %% | %.1f %% | %.0f MB / %.0f MB | %.0f %% | %.3f G | %s |" % (
            d["kernel"], d["time_ms"], d.get("regs", 0), d.get("warps_active_pct", 0), 100.0 * d.get("ipc_per_sm", 0) / 4.0, d.get("dram_pct", 0),
            d.get("dram_read_bytes", 0) / 1e6, d.get("dram_write_bytes", 0) / 1e6, d.get("l1tex_pct", 0), d.get("warp_instructions", 0) / 1e9,
            ", ".join(f"{k} {v:.1f}" for k, v in d["top_stalls"].items())))
    if len(sys.argv) > 2:
        grp = [d for d in out if any(k in d["kernel"] for k in ("preprocess_fwd", "scan_block", "tile_enumerate", "tile_ranges_scan", "tile_bucket_sort", "tile_group_sort",
                                                                 "coarse_scatter", "tile_scatter_entries"))]
        wi = {}
        for d in out:
            key = ("composite_fwd" if "composite2_fwd" in d["kernel"] or "composite_fwd" in d["kernel"] else
                   "composite_bwd" if "composite2_bwd" in d["kernel"] or "composite_bwd" in d["kernel"] else d["kernel"].split("<")[0])
            wi[key] = wi.get(key, 0) + int(d.get("warp_instructions", 0))
        json.dump({"workload": sys.argv[3] if len(sys.argv) > 3 else "zero123g/trained", "source": "ncu --set full --clock-control none, one step of bench.py (scripts/gpu_profile.sh); see the summary .md beside this file",
                   "hbm_group_dram_bytes": int(sum(d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0) for d in grp)),
                   "hbm_group_kernels": [d["kernel"] for d in grp], "warp_instructions": wi, "kernels": out}, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
