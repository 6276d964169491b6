#!/bin/bash
# 2-GPU check of the bench path (NCCL transport, parity_nccl, z periodic, spare SMs next to the persistent kernels).  usage: gpu_r2e.sh <tag> <ngpus>
TAG=${1:-r4d}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.log 2>&1; echo "rc=$?"; grep -E '^\{"metric' gpurun_out/${TAG}_bench_${N}gpu.log | cut -c1-1500; tail -5 gpurun_out/${TAG}_bench_${N}gpu.log | cut -c1-400
