#!/usr/bin/env python
"""Per-source-line instruction / stall-sample breakdown of one kernel from an `ncu --set full --import-source on`
report (kernel compiled with -lineinfo).   python profiles/srcprof.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr and r and r[0].isdigit():
            g = lambda i: int(r[i]) if r[i].lstrip("-").isdigit() else 0  # noqa: E731
            data.append((int(r[0]), r[1], g(iI), g(iT), g(iS)))
    return data


def main():
    data = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    tot = sum(d[2] for d in data) or 1
    tots = sum(d[4] for d in data) or 1
    print(f"total warp-inst {tot}  samples {tots}")
    src = open(__file__.replace("profiles/srcprof.py", "sigmarl_b200/csrc/sgb_kernels.cuh")).read().splitlines()
    # regions = runs of lines between "// @region name" markers, else fixed-size buckets of 25 lines
    marks = [(i + 1, l.split("@region", 1)[1].strip()) for i, l in enumerate(src) if "@region" in l]
    if marks:
        marks.append((10 ** 9, "end"))
        for (a, name), (b, _) in zip(marks, marks[1:]):
            s = sum(d[2] for d in data if a <= d[0] < b)
            t = sum(d[3] for d in data if a <= d[0] < b)
            sm = sum(d[4] for d in data if a <= d[0] < b)
            print(f"{name:34s} L{a:<5d} inst {100 * s / tot:5.1f}%  thr/inst {t / max(s, 1):5.1f}  samples {100 * sm / tots:5.1f}%")
    print()
    for d in sorted(data, key=lambda d: -d[2])[:top]:
        print(f"L{d[0]:<5d} {100 * d[2] / tot:4.1f}% thr {d[3] / max(d[2], 1):4.1f} smp {100 * d[4] / tots:4.1f}%  {d[1].strip()[:100]}")


if __name__ == "__main__":
    main()
