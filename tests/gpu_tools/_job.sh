S=$(date +%s)
timeout 400 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "t=$(( $(date +%s)-S ))s"
timeout 200 python bench.py > gpurun_out/bench_cfg4.log 2>&1
timeout 120 python bench.py --workload cfg5 --no-cpu > gpurun_out/bench_cfg5.log 2>&1
timeout 120 python bench.py --workload cfg2 --no-cpu > gpurun_out/bench_cfg2.log 2>&1
echo "t=$(( $(date +%s)-S ))s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_cfg4.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_cfg5.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:pack_cm_kernel -s 2 -c 1 -o gpurun_out/pack_cm_final -f python tests/gpu_tools/pack_bench.py 250000 2 > gpurun_out/pack_cm_ncu.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:pack_rm_kernel -s 2 -c 1 -o gpurun_out/pack_rm_final -f python tests/gpu_tools/pack_bench.py 250000 2 > gpurun_out/pack_rm_ncu.log 2>&1
python tests/gpu_tools/pack_bench.py
echo "t=$(( $(date +%s)-S ))s"
