timeout 300 python scripts/size_sweep.py xxsmall small medium xlarge xxlarge > gpurun_out/dec_base.txt 2>&1
NFB_LIB_PATH=/root/repo/build_ab/dec1.so timeout 300 python scripts/size_sweep.py xxsmall small medium xlarge xxlarge > gpurun_out/dec1.txt 2>&1
tail -n 6 gpurun_out/dec_base.txt gpurun_out/dec1.txt
