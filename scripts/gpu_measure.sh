#!/bin/bash
# The measurement pass behind profiles/: GPU test suite, bench (fp64, fp32), ncu launch list and full capture of the
# tile kernels on the bench workload.  Run on a B200 box:  gpurun --timeout 2400 -- 'bash scripts/gpu_measure.sh <tag>'
set -x
TAG=${1:-final}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
(time python bench.py --steps 10 --warmup 3) > gpurun_out/${TAG}_bench_fp64.log 2>&1; grep -E '^\{"metric' gpurun_out/${TAG}_bench_fp64.log | cut -c1-400
(time python bench.py --steps 10 --warmup 3 --precision 4 --no-cpu-baseline) > gpurun_out/${TAG}_bench_fp32.log 2>&1; grep -E '^\{"metric' gpurun_out/${TAG}_bench_fp32.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_ -s 8 -c 4 -o gpurun_out/${TAG}_prof256 python -m lfm_public_b200.tools.tune --n 256 --steps 1 > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
