#!/usr/bin/env python
"""SASS evidence for the claims DESIGN.md makes about the kernels (run where the library was built; no GPU needed):
per kernel the number of TMA bulk copies (UBLKCP), mbarrier operations (SYNCS), warp shuffles, named barriers, 128-bit
global loads / stores, MUFU ops, and that no tensor-core instruction exists — plus the first occurrence of each, with its
address, as an excerpt.      python profiles/sass_evidence.py [lib.so] > profiles/sass_evidence_r2.txt"""
import os, re, subprocess, sys, tempfile
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sigmarl_b200", "libsigmarl_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
pats = {"UBLKCP (TMA bulk copy global -> shared)": r"\bUBLKCP", "SYNCS (mbarrier arrive / try_wait)": r"\bSYNCS", "SHFL (warp shuffle)": r"\bSHFL",
        "BAR.SYNC (CTA / named barriers)": r"\bBAR\.", "LDG.E.128": r"\bLDG\.E\.128", "STG.E.128": r"\bSTG\.E\.128", "LDG.E.64": r"\bLDG\.E\.64",
        "STG.E.64": r"\bSTG\.E\.64", "LDS.128": r"\bLDS\.128", "LDS.64": r"\bLDS\.64", "MUFU (rcp / sqrt / ...)": r"\bMUFU", "FFMA": r"\bFFMA",
        "tensor core (HMMA / UTCMMA / *MMA)": r"MMA", "local memory (STL / LDL: spills)": r"\b(STL|LDL)\b", "VOTE / MATCH": r"\b(VOTE|MATCH)"}
print(f"# SASS evidence of {os.path.basename(so)} (sm_100a); nvdisasm of the embedded cubin\n")
name, body, fns = None, [], []
for l in txt:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        if name: fns.append((name, body))
        name, body = m.group(1), []
    elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        body.append(l.strip())
if name: fns.append((name, body))
for name, body in fns:
    dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
    print(f"## {dem}   [{len(body)} instructions]")
    for label, pat in pats.items():
        hits = [b for b in body if re.search(pat, b)]
        if hits or "tensor" in label:
            first = re.sub(r"\s+", " ", hits[0])[:110] if hits else "-"
            print(f"  {label:46s} {len(hits):5d}   first: {first}")
    print()
