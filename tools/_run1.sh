python -m pytest tests/test_gpu_parity.py tests/test_gpu_varsets.py tests/test_gpu_sampling.py -q -m gpu 2>&1 | tail -2
for k in 256 256 256 1024 2048; do TOPK_K=$k python tools/topk_once.py; done
