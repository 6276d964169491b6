set -x
mkdir -p gpurun_out
(time python scripts/run_example_dropin.py cylinder_vortex 0.06) > gpurun_out/r16_example_cylinder_vortex.log 2>&1; tail -25 gpurun_out/r16_example_cylinder_vortex.log
