python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/train_bench.py --steps 20 --warmup 5 2>gpurun_out/train8.err | tail -1 > gpurun_out/train_modelnet_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/train_modelnet_8gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/train_bench.py --steps 20 --warmup 5 2>gpurun_out/train2.err | tail -1 > gpurun_out/train_modelnet_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/train_modelnet_2gpu.json')); print(d['n_gpus'], d['ms_per_step'], d['value'])"
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -2
