set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -m gpu -x -q > gpurun_out/r23_pytest_gpu.log 2>&1; tail -4 gpurun_out/r23_pytest_gpu.log
for c in cylinder_vortex tv1; do tar xzf tmp_cases/$c.tgz -C /tmp; LFMGPU_PLAN_STATS=1 python -m lfm_public_b200.tools.run_case /tmp/$c 10 > gpurun_out/r23_run_$c.log 2>&1; cat gpurun_out/r23_run_$c.log | cut -c1-700; done
python -m lfm_public_b200.tools.tune --n 128 --steps 3 > gpurun_out/r23_tune.log 2>&1; cat gpurun_out/r23_tune.log | cut -c1-300
