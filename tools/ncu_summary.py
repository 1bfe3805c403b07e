#!/usr/bin/env python
"""Summarises an .ncu-rep (read on the CPU box): key raw metrics per kernel launch and the
top stall sites from the source page.  Usage: tools/ncu_summary.py file.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_ldgsts.sum"]


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[name_i][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:70s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            break
        body.append(r)
    tot = {s: 0 for s in stalls}
    for r in body:
        for s in stalls:
            try:
                tot[s] += int(r[ci[s]])
            except ValueError:
                pass
    T = max(1, sum(tot.values()))
    print("== stall mix (first launch):", {k: round(v / T, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]})
    print("== instructions:", len(body), " executed (warp-level):", sum(int(r[ci["Instructions Executed"]] or 0) for r in body))
    top = sorted(body, key=lambda r: -int(r[ci["# Samples"]] or 0))[:ntop]
    for r in top:
        st = {s: int(r[ci[s]] or 0) for s in stalls}
        print(f"   {r[ci['# Samples']]:>6s} smp  x{r[ci['Instructions Executed']]:>8s}  {max(st, key=st.get):22s} {r[ci['Source']][:80]}")


if __name__ == "__main__":
    main()
