#!/bin/bash
# Lean 8-GPU session (charged 8x): multi-device tests, the driver's own torchrun command (strong scaling of ONE 10 M batch,
# gather inside e2e), the tensor variant's e2e (PCIe ceiling), the one-process multi-device entry.
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e_topo_${N}gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/r2e_pytest_multi_${N}gpu.log 2>&1
tail -2 gpurun_out/r2e_pytest_multi_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2e_bench_strong_${N}gpu.json 2> gpurun_out/r2e_bench_strong_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --variant tensor --no-extras --no-cpu-baseline --no-ncu > gpurun_out/r2e_bench_tensor_${N}gpu.json 2> gpurun_out/r2e_bench_tensor_${N}gpu.err
timeout 300 python bench.py --gpus $N --no-extras --no-cpu-baseline --no-ncu > gpurun_out/r2e_bench_threads_${N}gpu.json 2> gpurun_out/r2e_bench_threads_${N}gpu.err
for f in gpurun_out/r2e_bench_*_${N}gpu.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','process_model')}, 'e2e', d.get('e2e',{}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('cores'))
except Exception as e: print('ERR', e)
"; done
tail -3 gpurun_out/r2e_bench_*_${N}gpu.err | tail -30
