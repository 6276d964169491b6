#!/bin/bash
# the round's measurement pass on the final code, every command time-boxed.  usage: gpu_final2.sh <tag>
TAG=${1:-r2b_final}
mkdir -p gpurun_out
LFMGPU_PLAN_STATS=1 timeout 120 python -m lfm_public_b200.tools.tune --n 128 --steps 5 > gpurun_out/${TAG}_tune128_bricks.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/${TAG}_tune128_bricks.log | cut -c1-250
LFMGPU_PLAN_STATS=1 timeout 120 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton > gpurun_out/${TAG}_tune128_morton.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/${TAG}_tune128_morton.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
(time timeout 700 python bench.py --steps 10 --warmup 3) > gpurun_out/${TAG}_bench_fp64.log 2>&1; echo "rc=$?"; grep -E '^\{"metric' gpurun_out/${TAG}_bench_fp64.log | cut -c1-700
timeout 500 ncu --set full --clock-control none --import-source on -k regex:_pipe -s 10 -c 2 -o gpurun_out/${TAG}_ncu256 python -m lfm_public_b200.tools.tune --n 256 --steps 1 --tile morton > gpurun_out/${TAG}_ncu256.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/${TAG}_ncu256.log | cut -c1-200
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "rc=$?"
LFM_FULL_SIZE_TESTS=1 timeout 450 python -m pytest tests/test_zz_large.py -m gpu -q -k 256 > gpurun_out/${TAG}_pytest_256cube.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_pytest_256cube.log
