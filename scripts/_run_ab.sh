timeout 300 python scripts/size_sweep.py xxsmall xsmall small medium large xlarge xxlarge > gpurun_out/grp4_sweep.txt 2>&1; cat gpurun_out/grp4_sweep.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
