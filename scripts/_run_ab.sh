timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -12
python bench.py > gpurun_out/bench_r1_groups_final.json 2> gpurun_out/bench_r1_groups_final.err; echo "bench rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_groups.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
cat /sys/kernel/mm/transparent_hugepage/enabled; nproc
