python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py tests/test_gpu_varsets.py -q -m gpu -x 2>&1 | tail -2
for b in 256 32; do python tools/train_bench.py --batch $b --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/train_b$b.json; done
python - <<'PY'
import json
a=json.load(open("gpurun_out/train_b256.json")); b=json.load(open("gpurun_out/train_b32.json"))
ka=a["kernels_ms_one_step"]; kb=b["kernels_ms_one_step"]
print("step", a["ms_per_step"], b["ms_per_step"], "kernel sums", sum(ka.values()), sum(kb.values()))
for k in ka:
    print(f"{k:34s} {ka[k]:8.4f} {kb.get(k,0):8.4f}  ratio {ka[k]/max(kb.get(k,1e-9),1e-9):5.2f}")
PY
