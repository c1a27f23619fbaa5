mkdir -p gpurun_out/r02
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/r02/launches_cfg4_r02.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r02/launches_cfg4.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02/launches_cfg2_r02.csv python bench.py --workload cfg2 --steps 20 --warmup 3 --no-graph --no-cpu > gpurun_out/r02/launches_cfg2.log 2>&1
$NCU --set full --import-source on -k regex:knn_search_kernel -s 6 -c 1 -o gpurun_out/r02/search_cfg4_r02 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu > gpurun_out/r02/search_cfg4.log 2>&1
$NCU --set full --import-source on -k regex:pack_cm_kernel -s 2 -c 1 -o gpurun_out/r02/pack_cm_r02 python tests/gpu_tools/pack_bench.py 250000 2 > gpurun_out/r02/pack_cm.log 2>&1
$NCU --set full --import-source on -k regex:pack_rm_kernel -s 2 -c 1 -o gpurun_out/r02/pack_rm_r02 python tests/gpu_tools/pack_bench.py 250000 2 > gpurun_out/r02/pack_rm.log 2>&1
K4_N=4000000 $NCU --set full --import-source on -k regex:gather_mean_warp_kernel -s 3 -c 1 -o gpurun_out/r02/gather_r02 python tools/gpu_k4_bench.py one > gpurun_out/r02/gather.log 2>&1
$NCU --set full --import-source on -k regex:knn_search_kernel -s 5 -c 1 -o gpurun_out/r02/collect_refined_r02 python tools/gpu_clustered_once.py bf16 > gpurun_out/r02/collect_refined.log 2>&1
$NCU --set full --import-source on -k regex:refine_prep_kernel -s 2 -c 1 -o gpurun_out/r02/refine_prep_r02 python tools/gpu_clustered_once.py bf16 > gpurun_out/r02/refine_prep.log 2>&1
python tools/gpu_clustered_once.py bf16 > gpurun_out/r02/clustered_bf16.log 2>&1; python tools/gpu_clustered_once.py fp16 > gpurun_out/r02/clustered_fp16.log 2>&1; python tools/gpu_clustered_once.py auto > gpurun_out/r02/clustered_auto.log 2>&1
cat gpurun_out/r02/clustered_*.log
python -m pytest tests/test_gpu_boundary.py -q -k "auto_format" 2>&1 | tail -3
ls -la gpurun_out/r02
