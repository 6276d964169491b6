set -x
mkdir -p gpurun_out
(time python scripts/run_example_dropin.py cvu_hilbert 0.02) > gpurun_out/r22_example_cvu_hilbert.log 2>&1; tail -12 gpurun_out/r22_example_cvu_hilbert.log
tar xzf tmp_cases/cvu.tgz -C /tmp
for c in cvu cvu_hilbert; do LFMGPU_PLAN_STATS=1 python -m lfm_public_b200.tools.run_case /tmp/$c 10 > gpurun_out/r22_run_$c.log 2>&1; cat gpurun_out/r22_run_$c.log | cut -c1-600; done
