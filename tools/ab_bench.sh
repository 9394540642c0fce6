#!/bin/bash
# A/B the experimental library variants on the bench workload: prints ms/step and the per-kernel times of each
for v in "$@"; do
  lib=point-cloud-audio_b200/csrc/libvar_$v.so
  PCAUDIO_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/ab_{v}.json"))
    k = d["kernels"]
    print(f"{v:10s} step {d['ms_per_step']:.4f} ms  apply {k['mab_apply_tc_kernel']['ms_per_step']:.4f}  reduce {k['mab_reduce_tc_kernel']['ms_per_step']:.4f}  pool {k['pma_pool_tc_kernel']['ms_per_step']:.4f}")
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/ab_{v}.err").read()[-400:])
PY
done
