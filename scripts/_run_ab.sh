timeout 600 python scripts/gpu_debug.py > gpurun_out/acc_silu_fast.txt 2>&1; echo "fast rc=$?"
NFB_LIB_PATH=/root/repo/build_ab/silu_lib.so timeout 600 python scripts/gpu_debug.py > gpurun_out/acc_silu_lib.txt 2>&1; echo "lib rc=$?"
timeout 300 python scripts/size_sweep.py > gpurun_out/grp3_sweep.txt 2>&1; cat gpurun_out/grp3_sweep.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
