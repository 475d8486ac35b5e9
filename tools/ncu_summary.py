#!/usr/bin/env python
"""Summarise `ncu --set full` captures for profiles/: python tools/ncu_summary.py out.json out.md rep1.ncu-rep [rep2 ...]
One JSON object per kernel (short name; the launch with the LONGEST duration when a kernel was captured several times, plus
the list of all its durations) with the metrics the roofline discussion in DESIGN.md uses; bench.py reads the DRAM bytes."""
import csv
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "lts__t_requests_srcunit_tex_op_red.sum",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.split("(")[0]
    return name.replace("lvba::", "").replace("nd::", "")


def main():
    out_json, out_md, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    best, durs, stalls = {}, {}, {}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            if len(r) <= ix["Kernel Name"]:
                continue
            k = short(r[ix["Kernel Name"]])
            d = float(r[ix["gpu__time_duration.sum"]])
            durs.setdefault(k, []).append(round(d, 2))
            if k in best and float(best[k]["gpu__time_duration.sum"]) >= d:
                continue
            rec = {"source": rep.split("/")[-1]}
            for m in KEEP:
                if m in ix and r[ix[m]] not in ("", "n/a"):
                    rec[m] = r[ix[m]]
                    rec[m + " [unit]"] = units[ix[m]]
            st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(r[i])) for h, i in ix.items()
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and r[i] not in ("", "n/a")}
            tot = sum(st.values()) or 1
            rec["stall_samples_top"] = {a: round(b / tot, 3) for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:5]}
            best[k] = rec
    for k in best:
        best[k]["durations_us_all_captured_launches"] = durs[k]
    json.dump(best, open(out_json, "w"), indent=1)
    with open(out_md, "w") as f:
        f.write("# ncu --set full, config C on one B200 (tools/dev_e2e.py C): one row per kernel, the longest captured launch\n\n")
        f.write("| kernel | us | DRAM MB (r+w) | DRAM % | L2 % | L1/TEX % | FP64 pipe % | issue % | warps active % | regs | grid x block | top stall reasons |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|\n")
        for k, r in sorted(best.items(), key=lambda kv: -float(kv[1]["gpu__time_duration.sum"])):
            g = lambda m, nd=1: (f"{float(r[m]):.{nd}f}" if m in r else "-")
            dram = float(r.get("dram__bytes_read.sum", 0)) + float(r.get("dram__bytes_write.sum", 0))
            f.write(f"| `{k}` | {g('gpu__time_duration.sum')} | {dram:.1f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} | {g('lts__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                    f"{g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')} | {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active')} | "
                    f"{g('sm__warps_active.avg.pct_of_peak_sustained_active')} | {g('launch__registers_per_thread', 0)} | {g('launch__grid_size', 0)} x {g('launch__block_size', 0)} | "
                    f"{', '.join(f'{a} {100 * b:.0f}%' for a, b in r['stall_samples_top'].items())} |\n")


if __name__ == "__main__":
    main()
