python bench.py > gpurun_out/bench_groups.json 2> gpurun_out/bench_groups.err; echo "bench rc=$?"
python bench.py --model-size xlarge --no-tensor-extra --no-cpu-baseline > gpurun_out/bench_groups_xlarge.json 2> gpurun_out/bench_groups_xlarge.err; echo "bench xlarge rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nfb_ffma_groups_kernel -s 1 -c 1 -f -o gpurun_out/ffma_xxxlarge_groups_10M python scripts/ncu_target_ffma.py xxxlarge 10000000 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nfb_ffma_groups_kernel -s 1 -c 1 -f -o gpurun_out/ffma_xlarge_groups_final python scripts/ncu_target_ffma.py xlarge 10000000 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:nfb_ffma_groups_kernel -s 1 -c 1 -f -o gpurun_out/ffma_xxsmall_groups_final python scripts/ncu_target_ffma.py xxsmall 10000000 > gpurun_out/ncu3.log 2>&1
