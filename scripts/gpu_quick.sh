#!/bin/bash
# quick GPU pass: parity suite + per-kernel timing of the bench workload.  usage: gpu_quick.sh <tag> [pytest -k expr] [tune args...]
set -x
TAG=${1:-q}; shift
KEXPR=${1:-}; shift
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
fi
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python -m lfm_public_b200.tools.tune --n 128 --steps 5 "$@" > gpurun_out/${TAG}_tune128.log 2>&1; tail -4 gpurun_out/${TAG}_tune128.log
timeout 400 python -m lfm_public_b200.tools.tune --n 256 --steps 5 "$@" > gpurun_out/${TAG}_tune256.log 2>&1; tail -4 gpurun_out/${TAG}_tune256.log
