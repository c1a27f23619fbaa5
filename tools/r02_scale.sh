# 8-GPU box: bit-identity of the sharded paths, the scaling bench at N = 8, 4, 2, 1 and the per-kernel profile of one step
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29511 tools/gpu_sharded_check.py > gpurun_out/r02/sharded_check_n8.log 2>&1; echo "check rc=$?"; tail -3 gpurun_out/r02/sharded_check_n8.log
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02/bench_n$n.json 2> gpurun_out/r02/bench_n$n.err; echo "bench n=$n rc=$?"
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extras > gpurun_out/r02/bench_n1.json 2> gpurun_out/r02/bench_n1.err; echo "bench n=1 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29531 tools/gpu_sharded_profile.py > gpurun_out/r02/profile_n8.log 2>&1; echo "profile rc=$?"
EXCHANGE=nccl timeout 300 $TR --nproc-per-node 8 --master-port 29532 tools/gpu_sharded_profile.py > gpurun_out/r02/profile_n8_nccl.log 2>&1; echo "profile nccl rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --exchange nccl > gpurun_out/r02/bench_n8_nccl.json 2> gpurun_out/r02/bench_n8_nccl.err; echo "bench n=8 nccl rc=$?"
for f in gpurun_out/r02/bench_n*.json; do python - "$f" <<'PY'
import sys, json
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], j["n_gpus"], round(j["value"]), round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"]), j["parity"]["ok"], j["roofline"]["achieved"])
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
ls gpurun_out/*.csv
