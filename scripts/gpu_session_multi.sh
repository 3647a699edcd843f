#!/bin/bash
# GPU session on N GPUs (N = $1): multi-device tests, the torchrun bench (strong scaling, gather inside e2e), the
# single-process multi-device bench, and the tensor variant under torchrun.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2m_topo_${N}gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/r2m_pytest_multi_${N}gpu.log 2>&1
tail -3 gpurun_out/r2m_pytest_multi_${N}gpu.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-extras --no-cpu-baseline > gpurun_out/r2m_bench_strong_1gpu.json 2> gpurun_out/r2m_bench_strong_1gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n > gpurun_out/r2m_bench_strong_${n}gpu.json 2> gpurun_out/r2m_bench_strong_${n}gpu.err
    timeout 600 python bench.py --gpus $n --no-extras > gpurun_out/r2m_bench_threads_${n}gpu.json 2> gpurun_out/r2m_bench_threads_${n}gpu.err
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --variant tensor --no-extras > gpurun_out/r2m_bench_tensor_${n}gpu.json 2> gpurun_out/r2m_bench_tensor_${n}gpu.err
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --impl reference --steps 2 > gpurun_out/r2m_bench_reference_${n}gpu.json 2> gpurun_out/r2m_bench_reference_${n}gpu.err
  fi
done
for f in gpurun_out/r2m_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','process_model')}, 'e2e', d.get('e2e',{}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('cores'))
except Exception as e: print('ERR', e)
"; done
tail -3 gpurun_out/r2m_bench_*.err | tail -40
