#!/usr/bin/env python
"""SASS of one kernel grouped by source line (needs -lineinfo):
    python profiles/sass_by_line.py <lib.so> <mangled-name substring> <first line> <last line> [count]
`count` prints only the number of instructions per line."""
import os, re, subprocess, sys, tempfile

so, kname, a, b = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
count_only = len(sys.argv) > 5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(txt) if l.startswith("//----") and kname in l and ".text." in l)
end = next((i for i in range(start + 1, len(txt)) if txt[i].startswith("//----")), len(txt))
cur, per, total = None, {}, 0
for l in txt[start:end]:
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith("sgb_kernels.cuh") else None
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        total += 1
        if cur is not None:
            per.setdefault(cur, []).append(re.sub(r"/\*[0-9a-f]+\*/", "", l).strip())
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "sigmarl_b200", "csrc", "sgb_kernels.cuh")).read().splitlines()
print(f"{kname}: {total} SASS instructions")
for ln in range(a, b + 1):
    if ln in per:
        print(f"L{ln} [{len(per[ln])}] {src[ln - 1].strip()[:100]}")
        if not count_only:
            for i in per[ln]:
                print("        " + i[:80])
