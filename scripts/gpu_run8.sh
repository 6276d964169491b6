set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py -m gpu -x -q > gpurun_out/r8_dropin.log 2>&1; tail -30 gpurun_out/r8_dropin.log
timeout 900 python -m lfm_public_b200.tools.tune --n 256 --steps 3 --set LFMGPU_STAGE_CFG=6 > gpurun_out/r8_tune_morton.log 2>&1; cat gpurun_out/r8_tune_morton.log
timeout 900 python -m lfm_public_b200.tools.tune --n 256 --steps 3 --brick-order lex > gpurun_out/r8_tune_lex.log 2>&1; cat gpurun_out/r8_tune_lex.log
