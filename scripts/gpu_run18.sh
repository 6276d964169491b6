set -x
mkdir -p gpurun_out
(time python scripts/run_example_dropin.py tandem_vortex 0.09 2) > gpurun_out/r18_example_tandem_2gpu_nccl.log 2>&1; tail -30 gpurun_out/r18_example_tandem_2gpu_nccl.log
