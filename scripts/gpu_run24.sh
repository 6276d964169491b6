set -x
mkdir -p gpurun_out
for c in cv_hilbert; do tar xzf tmp_cases/$c.tgz -C /tmp; LFMGPU_PLAN_STATS=1 python -m lfm_public_b200.tools.run_case /tmp/$c 10 > gpurun_out/r24_run_$c.log 2>&1; cat gpurun_out/r24_run_$c.log | cut -c1-700; done
