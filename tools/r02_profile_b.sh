# second profiling batch of round 2: the resident-query K2 kernel (cfg4, cfg3, cfg5) + launch lists
mkdir -p gpurun_out/r02
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r02/launches_cfg4_r02.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r02/launches_cfg4.log 2>&1
$NCU --set full --import-source on -k regex:knn_search_resident_kernel -s 6 -c 1 -o gpurun_out/r02/search_resident_cfg4_r02 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu > gpurun_out/r02/search_resident_cfg4.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed -k regex:knn_search -s 4 -c 1 --csv --log-file gpurun_out/r02/search_cfg3_metrics_r02.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-extras --no-cpu > gpurun_out/r02/search_cfg3.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed -k regex:knn_search -s 3 -c 1 --csv --log-file gpurun_out/r02/search_cfg5_metrics_r02.csv python bench.py --workload cfg5 --steps 1 --warmup 2 --no-extras --no-cpu > gpurun_out/r02/search_cfg5.log 2>&1
ALIVE_KNN_RESIDENT=0 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed -k regex:knn_search -s 6 -c 1 --csv --log-file gpurun_out/r02/search_cfg4_noresident_metrics_r02.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu > gpurun_out/r02/search_cfg4_nores.log 2>&1
ls -la gpurun_out/r02
