set -x
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
(time timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/r19_bench_${N}gpu.log 2>&1; grep -E '^\{"metric' gpurun_out/r19_bench_${N}gpu.log | cut -c1-1300; tail -5 gpurun_out/r19_bench_${N}gpu.log | cut -c1-300
