#!/bin/bash
# GPU session A (1 GPU): the whole GPU test suite, the default bench line, the reference arm, sanitizer passes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --durations=20 -p no:cacheprovider > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err
echo "bench exit $?" >> gpurun_out/r2a_bench_1gpu.err
timeout 600 python bench.py --impl reference > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_target.py > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r2a_sanitizer_$tool.log
done
tail -5 gpurun_out/r2a_pytest_gpu.log
