#!/bin/bash
# A/B of the planner / kernel switches on the bench workload: profiles/ab_bench.sh <tag> "<ENV=.. ENV=..>" ...
tag=$1; shift
for cfg in "$@"; do
  name=$(echo "$cfg" | tr ' =/' '___' | sed 's/.*tensorrl_qas_b200_//')
  env $cfg timeout 180 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - "$cfg" gpurun_out/${tag}_${name}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    ks = " ".join(f"{k['ms'] * 1e3:.0f}" for k in d["roofline"].get("kernels", []))
    print(f"{sys.argv[1]:50s} {d['value']:10.1f} evals/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}  us/kernel [{ks}]")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
