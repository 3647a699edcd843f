timeout 300 python scripts/size_sweep.py xxxlarge > gpurun_out/lead_2.txt 2>&1
for v in lead1 lead3; do NFB_LIB_PATH=/root/repo/build_ab/$v.so timeout 300 python scripts/size_sweep.py xxxlarge > gpurun_out/$v.txt 2>&1; done
tail -n 2 gpurun_out/lead*.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
