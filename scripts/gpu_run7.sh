set -x
mkdir -p gpurun_out
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r7_bench_fp64.log 2>&1; tail -5 gpurun_out/r7_bench_fp64.log | cut -c1-1500
(time python bench.py --steps 10 --warmup 3 --precision 4 --no-cpu-baseline) > gpurun_out/r7_bench_fp32.log 2>&1; tail -5 gpurun_out/r7_bench_fp32.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r7_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r7_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_ -s 8 -c 4 -o gpurun_out/r7_prof256 python -m lfm_public_b200.tools.tune --n 256 --steps 1 > gpurun_out/r7_ncu.log 2>&1; tail -2 gpurun_out/r7_ncu.log
