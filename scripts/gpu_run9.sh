set -x
mkdir -p gpurun_out
nvidia-smi -L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
(time timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --n 128 --no-cpu-baseline) > gpurun_out/r9_bench2_n128.log 2>&1; tail -4 gpurun_out/r9_bench2_n128.log | cut -c1-900
(time timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3) > gpurun_out/r9_bench2_n256.log 2>&1; tail -4 gpurun_out/r9_bench2_n256.log | cut -c1-1200
