#!/bin/bash
# 8-GPU check of the bench path at a reduced block size (the driver runs the full size): 2x2x2 blocks, z periodic through
# processorCyclic patches, NCCL transport against the oracle (parity_nccl).   usage: gpu_r2f.sh <tag> <ngpus> <n>
# (the block size travels in LFM_BENCH_N: torchrun's own argument parser trips over `--n`)
TAG=${1:-r4e}; N=${2:-8}; NN=${3:-128}
mkdir -p gpurun_out
LFM_BENCH_N=$NN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-extras > gpurun_out/${TAG}_bench_${N}gpu_n${NN}.log 2>&1; echo "rc=$?"; grep -E '^\{"metric' gpurun_out/${TAG}_bench_${N}gpu_n${NN}.log | cut -c1-700; tail -3 gpurun_out/${TAG}_bench_${N}gpu_n${NN}.log | cut -c1-300
