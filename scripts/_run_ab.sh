NFB_LIB_PATH=/root/repo/build_ab/tnw8.so timeout 300 python scripts/size_sweep.py large xlarge xxlarge xxxlarge > gpurun_out/tnw8.txt 2>&1
cat gpurun_out/tnw8.txt
