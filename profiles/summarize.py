#!/usr/bin/env python
"""Turn ncu output into the markdown summaries kept in this directory.

    python profiles/summarize.py launches  gpurun_out/r01_launches_final.csv
    python profiles/summarize.py full      gpurun_out/r01_edge_fwd_final.ncu-rep [more.ncu-rep ...]

`launches`: per-kernel totals / shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
`full`:     the metrics quoted in DESIGN.md from `ncu --set full` captures (one column per captured launch).
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("duration us", "gpu__time_duration.sum"),
    ("dram read MB", "dram__bytes_read.sum"),
    ("dram write MB", "dram__bytes_write.sum"),
    ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 % of peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1/smem % of peak", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("SM % of peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe % active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("issue slots % busy", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block KB", "launch__shared_mem_per_block_dynamic"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("cycles per issued inst", "smsp__average_warp_latency_per_inst_issued.ratio"),
    ("stall long_scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall barrier / issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall no_instruction / issue", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:78]
        if "spin_kernel" in name:      # torch.cuda._sleep of bench.py's per-kernel timing pass (keeps the host ahead): not work
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v = v / 1000 if r[ix["Metric Unit"]] == "ns" else v * 1000 if r[ix["Metric Unit"]] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "<unnamed>" in k)
    print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"| `{k}` | {a[0]} | {a[1]:.0f} | {a[1] / tot * 100:.1f}% | {a[1] / a[0]:.1f} |")
    print(f"\nTotal {tot:.0f} us over {len(rows) - 1} launches; libgp_b200 kernels = {ours / tot * 100:.1f}% of GPU time.")


def full(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        print(f"\n### `{path.split('/')[-1]}`\n")
        names = [re.sub(r"\(.*", "", d[ix["Kernel Name"]]).replace("void <unnamed>::", "") for d in data]
        print("| metric | " + " | ".join(f"launch {i}: `{n}`" for i, n in enumerate(names)) + " |")
        print("|---|" + "---|" * len(data))
        for label, m in METRICS:
            if m not in ix:
                continue
            vals = []
            for d in data:
                v, u = d[ix[m]], units[ix[m]]
                try:
                    f = float(v.replace(",", ""))
                    if label.startswith("dram") and u in ("byte", "Kbyte", "Mbyte", "Gbyte"):
                        f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
                    if label.startswith("duration"):
                        f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
                    if label.startswith("dyn smem") and u == "byte":
                        f /= 1024
                    vals.append(f"{f:,.2f}" if f < 1e6 else f"{f:,.0f}")
                except ValueError:
                    vals.append(v)
            print(f"| {label} (`{m}`) | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2] if sys.argv[1] == "launches" else sys.argv[2:])
