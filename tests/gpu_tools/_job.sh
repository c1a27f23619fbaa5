timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 84 -c 40 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_cfg4.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 258 -c 40 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_cfg5.log 2>&1
tail -12 gpurun_out/launches_cfg4.csv | cut -c1-200
tail -12 gpurun_out/launches_cfg5.csv | cut -c1-200
