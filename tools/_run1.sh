python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -c 300 gpurun_out/bench_r02c.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r02c.json"))
print(d["ms_per_step"], d["value"], d["e2e"]["value"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
print(d["roofline"])
PY
