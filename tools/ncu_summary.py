"""Condense an `ncu -i X.ncu-rep --page raw --csv` export into a per-launch JSON summary (profiles/r02_*_ncu_summary.json).

    python tools/ncu_summary.py gpurun_out/r02_ncu_chain.raw.csv profiles/r02_chain_ncu_summary.json
"""
import csv
import json
import re
import sys

WANT = {   # summary key -> substring of the ncu metric column (first match wins)
    "duration_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_gbs": "dram__bytes.sum.per_second",
    "dram_read_pct_of_peak": "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_pipe_active_busiest_sm_pct": "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "registers_per_thread": "launch__registers_per_thread",
    "dynamic_smem_bytes": "launch__shared_mem_per_block_dynamic",
    "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm_active_cycles": "sm__cycles_active.avg",
    "ipc": "sm__inst_executed.avg.per_cycle_elapsed",
}
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main(src, dst):
    rows = list(csv.reader(open(src, newline="")))
    header, units, data = rows[0], rows[1], rows[2:]
    cols = {}
    for key, needle in WANT.items():
        exact = [i for i, name in enumerate(header) if name == needle]
        loose = [i for i, name in enumerate(header) if needle in name and "Triage" not in name]
        if exact or loose:
            cols[key] = (exact or loose)[0]
    name_col, grid_col, block_col = header.index("Kernel Name"), header.index("Grid Size"), header.index("Block Size")
    out = []
    for r in data:
        if len(r) < len(header):
            continue
        rec = {"kernel": re.sub(r"\(.*", "", r[name_col].replace("void <unnamed>::", "").replace("<unnamed>::", "")),
               "grid": r[grid_col], "block": r[block_col]}
        for key, i in cols.items():
            try:
                val = float(r[i].replace(",", ""))
            except ValueError:
                continue
            unit = units[i]
            if (key.endswith("_bytes") or key == "duration_us") and unit in UNIT_SCALE:
                val *= UNIT_SCALE[unit]
            if key == "dram_gbs":
                val *= {"Tbyte/s": 1e3, "Gbyte/s": 1.0, "Mbyte/s": 1e-3, "Kbyte/s": 1e-6, "byte/s": 1e-9}.get(unit, 1.0)
            rec[key] = round(val, 3)
        out.append(rec)
    by_kernel = {}
    for rec in out:
        k = by_kernel.setdefault(rec["kernel"], {"launches": 0, "duration_us": 0.0, "dram_bytes": 0.0, "tensor_pct_sum": 0.0, "dram_pct_sum": 0.0})
        k["launches"] += 1
        k["duration_us"] += rec.get("duration_us", 0.0)
        k["dram_bytes"] += rec.get("dram_read_bytes", 0.0) + rec.get("dram_write_bytes", 0.0)
        k["tensor_pct_sum"] += rec.get("tensor_pipe_active_pct", 0.0)
        k["dram_pct_sum"] += rec.get("dram_gbs", 0.0)
    total = sum(k["duration_us"] for k in by_kernel.values()) or 1.0
    families = {name: {"launches": k["launches"], "total_us": round(k["duration_us"], 2), "share_of_capture": round(k["duration_us"] / total, 4),
                       "avg_us": round(k["duration_us"] / k["launches"], 2), "avg_dram_bytes": round(k["dram_bytes"] / k["launches"]),
                       "avg_tensor_pipe_active_pct": round(k["tensor_pct_sum"] / k["launches"], 2),
                       "avg_dram_gbs": round(k["dram_pct_sum"] / k["launches"], 1)} for name, k in by_kernel.items()}
    chain = [r for r in out if r["kernel"].startswith("decode_chain_kernel")]
    summary = {"source": src, "how": "ncu --set full --clock-control none (cold caches, serialised launches: shares and per-launch traffic, not the bench's timings)",
               "metric_columns": {k: rows[0][i] for k, i in cols.items()}, "families": families, "launches": out}
    if chain:
        summary["avg_dram_bytes_per_launch"] = round(sum(r.get("dram_read_bytes", 0) + r.get("dram_write_bytes", 0) for r in chain) / len(chain))
    json.dump(summary, open(dst, "w"), indent=1)
    for name, fam in sorted(families.items(), key=lambda kv: -kv[1]["total_us"]):
        print(f"{name[:60]:60s} x{fam['launches']:3d} avg {fam['avg_us']:8.2f} us share {fam['share_of_capture']:.3f} dram {fam['avg_dram_bytes'] / 1e6:7.2f} MB "
              f"tensor pipe active {fam['avg_tensor_pipe_active_pct']:5.1f}% of SM-active cycles, dram {fam['avg_dram_gbs']:7.1f} GB/s")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
