#!/bin/bash
# GPU session B (1 GPU): whole GPU suite, default bench line, tensor throughput of every size, tcn phase stamps, latency.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/r2f_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest_gpu.log
cp gpurun_out/parity_maxima.json gpurun_out/r2f_parity_maxima.json
timeout 900 python bench.py > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err
echo "bench exit $?" >> gpurun_out/r2f_bench_1gpu.err
timeout 400 python scripts/tensor_throughput.py > gpurun_out/r2f_tensor_throughput.jsonl 2>&1
for v in tensor tensor_fp32; do for s in xlarge medium; do timeout 100 python scripts/gpu_debug_tcn_stamps.py $v $s; done; done > gpurun_out/r2f_tcn_stamps.txt 2>&1
timeout 300 python scripts/latency_sweep.py > gpurun_out/r2f_latency_sweep.jsonl 2>&1
timeout 200 python scripts/latency_breakdown.py > gpurun_out/r2f_latency_breakdown.txt 2>&1
tail -4 gpurun_out/r2f_pytest_gpu.log
