#!/usr/bin/env python3
"""Summarise an `ncu --set full` capture and write the per-kernel DRAM traffic bench.py reports.

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_extract.py raw.csv [--json profiles/r02_traffic.json] [--tag r02] > profiles/r02_ncu_summary.txt

The JSON maps the library's launch names (the keys of `kernels_ms_and_launches_per_pass` in bench.py's line) to
{"kernel": demangled name, "dram_read": bytes, "dram_write": bytes, "traffic": read + write, "time_us": duration under
ncu (cold cache, serialised: not a bench value), ...}; the LAST launch of every kernel in the capture is kept (the
first ones may belong to the warm-up path of the library).  bench.py reads `profiles/*_traffic.json` (newest tag) for
`roofline.traffic`: nothing is hard-coded there."""
import csv
import json
import sys

NAMES = [  # substring of the kernel name -> launch name of the library's profiler
    ("k_asm_fans", "asm_rows_p1"), ("k_asm_tiles", "asm_rows_p1"), ("k_rhs_tiles", "rhs_rows"), ("k_rhs_fans", "rhs_rows"),
    ("k_sym_p1_fused", "sym_p1_fused"), ("k_spmv_sell", "cg_spmv_dots"), ("k_p2p_halo", "p2p_halo"), ("k_asm_p2", "asm_rows_p2"),
    ("k_spmv_nodeblock", "cg_spmv_dots_nodeblock"), ("k_cg_update_g", "cg_update_g"), ("k_cg_update_xh", "cg_update_xh"),
    ("k_cg_fused", "cg_persistent"),
]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio"]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    u = unit.lower()
    return f * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


def to_us(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit.lower(), 1.0)


def main():
    args = sys.argv[1:]
    js = None
    tag = ""
    if "--json" in args:
        i = args.index("--json")
        js = args[i + 1]
        del args[i:i + 2]
    if "--tag" in args:
        i = args.index("--tag")
        tag = args[i + 1]
        del args[i:i + 2]
    rows = list(csv.reader(open(args[0])))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[start], rows[start + 1]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[start + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        print("----", name[:150])
        for w in WANT:
            if w in col:
                print(f"  {w:82s} {r[col[w]][:40]} {units[col[w]]}")
        key = next((k for sub, k in NAMES if sub in name), None)
        if key and "dram__bytes_read.sum" in col:
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            out[key] = {"kernel": name.split("(")[0][-60:], "dram_read": rd, "dram_write": wr, "traffic": rd + wr,
                        "time_us_under_ncu": to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]]),
                        "smem_wavefronts": float(r[col["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]].replace(",", "")) if "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum" in col else None,
                        "inst_executed": float(r[col["smsp__inst_executed.sum"]].replace(",", "")) if "smsp__inst_executed.sum" in col else None,
                        "capture": tag}
    if js:
        with open(js, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        print("wrote", js, sorted(out))


if __name__ == "__main__":
    main()
