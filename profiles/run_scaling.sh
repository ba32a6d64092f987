#!/bin/bash
# Multi-GPU bench lines on one box: profiles/run_scaling.sh [all|c5|small8|small24]  (under `gpurun --gpus 8`; small24: --gpus 4)
# C5 weak scaling at N = 1, 2, 4, 8 and strong scaling (global batch 64) at N = 2, 4, 8; C1-C4 weak at N = 8 / at N = 2, 4.
mkdir -p gpurun_out
run() {  # run <tag> <ngpus> <args...>
  tag=$1; n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/scale_r02_$tag.json 2> gpurun_out/scale_r02_$tag.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
      bench.py --gpus $n "$@" > gpurun_out/scale_r02_$tag.json 2> gpurun_out/scale_r02_$tag.err
  fi
  python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/scale_r02_{tag}.json") if l.startswith("{")][-1])
    print(f"{tag:14s} N={d['n_gpus']} {d['scaling']:6s} {d['value']:14.1f} evals/s {d['ms_per_step']:.4f} ms/step e2e {d['e2e']['value']:.1f}")
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/scale_r02_{tag}.err").read()[-1200:])
PY
}
what=${1:-all}
if [ "$what" = "all" ] || [ "$what" = "c5" ]; then     # needs 8 GPUs
  for n in 1 2 4 8; do run C5_weak_$n $n --workload C5 --steps 20 --warmup 3 --no-cpu-baseline; done
  for n in 2 4 8; do run C5_strong_$n $n --workload C5 --scaling strong --batch 64 --steps 20 --warmup 3 --no-cpu-baseline; done
fi
if [ "$what" = "all" ] || [ "$what" = "small8" ]; then  # needs 8 GPUs
  for w in C1 C2 C3 C4; do run ${w}_weak_8 8 --workload $w --steps 20 --warmup 3 --no-cpu-baseline; done
fi
if [ "$what" = "small24" ]; then                        # needs 4 GPUs: the small configurations at N = 2, 4
  for w in C1 C2 C3 C4; do for n in 2 4; do run ${w}_weak_$n $n --workload $w --steps 20 --warmup 3 --no-cpu-baseline; done; done
fi
