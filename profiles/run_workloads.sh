#!/bin/bash
# bench lines of every workload at N = 1 (profiles/bench_r02_<workload>.json): profiles/run_workloads.sh [steps]
steps=${1:-20}
mkdir -p gpurun_out
for w in C1 C2 C3 C4 C5 C5L C5G; do
  timeout 600 python bench.py --workload $w --steps $steps --warmup 3 > gpurun_out/bench_r02_$w.json 2> gpurun_out/bench_r02_$w.err
  python - $w <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_r02_{w}.json"))
    r = d["roofline"]
    print(f"{w:4s} {d['value']:12.1f} evals/s  {d['ms_per_step']:.4f} ms/step  e2e {d['e2e']['value']:.1f}  dominant {r['kernel']} "
          f"share {r['measured']['share_of_step']:.2f} model-frac {r['measured']['frac_of_peak_model']:.3f}  step-model-frac "
          f"{r['step']['frac_of_peak_model']:.3f} dmma {r['fp64_tensor']['achieved']:.2f}/{r['fp64_tensor']['peak']:.1f} TF  "
          f"cpu {d['cpu_baseline']['value']:.1f} ({d['cpu_baseline']['cores']} cores) dE {d['cpu_baseline']['max_abs_dE_vs_gpu']:.1e}"
          + (f"  B1 {d['latency']['us_per_eval_B1_host_call']:.1f} us" if d.get('latency') else ""))
except Exception as e:
    print(w, "FAILED", e)
    print(open(f"gpurun_out/bench_r02_{w}.err").read()[-1500:])
PY
done
