#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launches csv into the small tracked summaries under profiles/.

    python profiles/summarize.py gpurun_out/prof_step.ncu-rep gpurun_out/launches.csv r1
"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")], "id": r[0]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for i, h in enumerate(hdr)
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued") and r[i] not in ("", "n/a")}
        tot = sum(stalls.values()) or 1.0
        d["stall_sample_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.01}
        out.append(d)
    with open(os.path.join(HERE, f"ncu_full_{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)
    first = max(out, key=lambda d: float(d.get("gpu__time_duration.sum", "0 us").split()[0]))   # the fused step kernel
    def to_b(s):
        v, u = s.split()
        return float(v) * TO_BYTES.get(u, 1)
    with open(os.path.join(HERE, f"ncu_step_kernel_{tag}.json"), "w") as f:
        json.dump({"kernel": first["kernel"], "duration": first.get("gpu__time_duration.sum"),
                   "dram_bytes_read": to_b(first["dram__bytes_read.sum"]), "dram_bytes_write": to_b(first["dram__bytes_write.sum"]),
                   "source": os.path.basename(rep), "note": "one launch, ncu --set full --clock-control none, B=65536 N=8 cpm_entire"}, f, indent=1)
    # launch list -> per-kernel totals and shares
    agg, n = defaultdict(float), defaultdict(int)
    with open(launches) as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        h = next(rd)
        for r in rd:
            name = r[h.index("Kernel Name")].split("(")[0][:90]
            agg[name] += float(r[h.index("Metric Value")])
            n[name] += 1
    tot = sum(agg.values())
    with open(os.path.join(HERE, f"launches_{tag}.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` over bench.py timed steps\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{k}` | {n[k]} | {v / 1e3:.1f} | {100 * v / tot:.1f}% |\n")
    print("wrote", [x for x in os.listdir(HERE) if tag in x])


if __name__ == "__main__":
    main()
