#!/usr/bin/env python
"""Turn the round's ncu reports (gpurun_out/<tag>_*.ncu-rep) and the bench launch list into the tracked summaries under
profiles/: ncu_full_<tag>.json (every kernel family, with the commit the library was built from), launches_<tag>.md,
srcprof_step_<tag>.txt.

    python profiles/summarize.py r2 [commit]
"""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
OUT = os.path.join(REPO, "gpurun_out")
KEYS = {"gpu__time_duration.sum": "duration_us", "launch__grid_size": "grid", "launch__block_size": "block",
        "launch__registers_per_thread": "registers", "launch__shared_mem_per_block_dynamic": "smem_dynamic_kb",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__inst_executed.sum": "warp_inst_per_launch",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_warp_inst",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct"}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def kernels_of(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {}
        for k, short in KEYS.items():
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                d[short] = float(r[hdr.index(k)].replace(",", "")) * SCALE.get(units[hdr.index(k)], 1)
        stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i]) for i, h in enumerate(hdr)
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued") and r[i] not in ("", "n/a")}
        tot = sum(stalls.values()) or 1.0
        d["stall_sample_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot > 0.01}
        yield name, d


def short_name(name):
    m = re.search(r"env_step_kernel<\(?(?:int\))?(\d)[^0-9]+(\d)[^0-9]+(\d)", name)
    if m:
        return f"env_step_kernel<{m.group(1)},{m.group(2)},{m.group(3)}>"
    return re.sub(r"\(.*", "", name).replace("sgb::", "").replace("void ", "").strip()


def main():
    tag = sys.argv[1]
    commit = sys.argv[2] if len(sys.argv) > 2 else subprocess.run(["git", "rev-parse", "HEAD"], cwd=REPO, capture_output=True, text=True).stdout.strip()
    src = open(os.path.join(REPO, "sigmarl_b200", "csrc", "sgb_kernels.cuh"), "rb").read()
    reports = {"step_g4": "cpm_entire 65536 envs x 8 agents (headline shape)", "reset": "cpm_entire 65536 x 8, masked reset of ~26 % done envs",
               "step_g2_n12": "roundabout_2 8192 envs x 12 agents (BASELINE configs[3] per-GPU shape; two lanes per agent)",
               "step_ov1": "cpm_entire 65536 x 8, flag-driven observation writer (steering + neighbours' reference paths, D = 48)"}
    out = {"round": tag, "commit": commit, "sgb_kernels_cuh_sha256_16": hashlib.sha256(src).hexdigest()[:16], "shape_agents": 65536 * 8,
           "how": "ncu --set full --import-source on --clock-control none, one launch each after 8 warm-up launches (profiles/scripts/evidence.sh)",
           "kernels": {}}
    for key, shape in reports.items():
        rep = os.path.join(OUT, f"{tag}_{key}.ncu-rep")
        if not os.path.exists(rep):
            continue
        for name, d in kernels_of(rep):
            sn = short_name(name)
            if key != "step_g4":
                sn = {"step_g2_n12": sn + " @N=12", "step_ov1": sn, "reset": sn}[key]
            d["workload"] = shape
            d["report"] = os.path.basename(rep)
            out["kernels"].setdefault(sn, d)
    with open(os.path.join(HERE, f"ncu_full_{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("kernels:", {k: (round(v.get("duration_us", 0), 1), v.get("registers")) for k, v in out["kernels"].items()})
    # launch list of the bench
    lc = os.path.join(OUT, f"launches_{tag}.csv")
    if os.path.exists(lc):
        rows = list(csv.reader(l for l in open(lc) if l.startswith('"')))
        h = rows[0]
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            try:
                v = float(r[h.index("Metric Value")].replace(",", ""))
            except Exception:
                continue
            u = r[h.index("Metric Unit")]
            us = v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
            a = agg[r[h.index("Kernel Name")]]
            a[0] += 1; a[1] += us
        tot = sum(a[1] for a in agg.values()) or 1
        with open(os.path.join(HERE, f"launches_{tag}.md"), "w") as f:
            f.write(f"# ncu launch list ({tag}, commit {commit[:10]}): `ncu --metrics gpu__time_duration.sum --clock-control none` over "
                    f"`bench.py --steps 20 --warmup 3 --no-rollout`\n\nPer-launch times under ncu are cold-cache and serialised: compare "
                    f"SHARES, not absolutes.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
            for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"| `{k[:90]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |\n")
        subprocess.run(["cp", lc, os.path.join(HERE, f"launches_{tag}.csv")])
    rep = os.path.join(OUT, f"{tag}_step_g4.ncu-rep")
    if os.path.exists(rep):
        txt = subprocess.run([sys.executable, os.path.join(HERE, "srcprof.py"), rep, "60"], capture_output=True, text=True).stdout
        open(os.path.join(HERE, f"srcprof_step_{tag}.txt"), "w").write(
            f"# per-source-line instruction / stall-sample profile of the step + refresh kernels ({tag}, commit {commit[:10]}); lines of sgb_kernels.cuh\n" + txt)


if __name__ == "__main__":
    main()
