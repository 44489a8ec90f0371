#!/usr/bin/env python
"""Condense `ncu --set full` reports (.ncu-rep) into a small markdown table for profiles/.

    python tools/ncu_summary.py title rep1.ncu-rep [rep2.ncu-rep ...] > profiles/rNN_xxx_ncu.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA bytes L2->smem"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tcgen05 pipe active % (of SM-active)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "HMMA subpipe active cycles"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank-conflict wavefronts"),
]
STALL = "smsp__pcsamp_warps_issue_stalled_"


def main():
    title, reps = sys.argv[1], sys.argv[2:]
    print(f"# {title}\n")
    print("`ncu --set full --clock-control none --import-source on`, one launch each; a number measured under the profiler is not a "
          "bench value (it is ~5-15 % slower than the CUDA-event timing).\n")
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"## {rep}: unreadable\n")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].replace("void ", "").replace("unnamed>::", "")
            print(f"## `{name[:110]}`  ({rep.split('/')[-1]})\n")
            print("| metric | value |\n|---|---|")
            for key, label in METRICS:
                if key in hdr and r[hdr.index(key)] not in ("", "n/a"):
                    print(f"| {label} | {r[hdr.index(key)]} {units[hdr.index(key)]} |")
            stalls = [(h[len(STALL):], float(r[i])) for i, h in enumerate(hdr) if h.startswith(STALL) and not h.endswith("_not_issued") and r[i] not in ("", "0")]
            tot = sum(v for _, v in stalls) or 1.0
            top = sorted(stalls, key=lambda kv: -kv[1])[:6]
            print("| warp-stall samples (top) | " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top) + " |")
            print()


if __name__ == "__main__":
    main()
