#!/bin/bash
# programmatic dependent launch A/B: bench line (with the rollout record) and kernel timings with and without it
mkdir -p gpurun_out
L=gpurun_out/pdl.log
: > $L
for v in 0 1; do
  echo "== SGB_NO_PDL=$v" >> $L
  SGB_NO_PDL=$v KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 65536 40 >> $L 2>&1
  SGB_NO_PDL=$v KB_WRITE_OBS=1 timeout 120 python profiles/kbench.py 8192 40 >> $L 2>&1
  SGB_NO_PDL=$v timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_pdl$v.json 2>> $L
  python - >> $L <<PY
import json
d=json.load(open("gpurun_out/bench_pdl$v.json")); r=d["rollout"]
print("bench: value %.4e ms/step %.4f kernel_ms %.4f e2e %.3e rollout %.4e collect %.3f ms" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], r["value"], r["breakdown_ms"]["collect"]))
PY
done
timeout 300 python profiles/exact_check.py 32768 12 >> $L 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 >> $L
cat $L
