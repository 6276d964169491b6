set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipe_io.py -m gpu -x -q > gpurun_out/r25_pytest.log 2>&1; tail -6 gpurun_out/r25_pytest.log
(time python bench.py --steps 10 --warmup 3) > gpurun_out/r25_bench_fp64.log 2>&1; grep -E '^\{"metric' gpurun_out/r25_bench_fp64.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','roofline')})"; tail -4 gpurun_out/r25_bench_fp64.log | cut -c1-200
