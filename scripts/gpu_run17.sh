set -x
mkdir -p gpurun_out
(time LFMGPU_PLAN_STATS=1 python scripts/run_example_dropin.py cvu 0.02) > gpurun_out/r17_example_cvu.log 2>&1; tail -22 gpurun_out/r17_example_cvu.log
