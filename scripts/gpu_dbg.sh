#!/bin/bash
mkdir -p gpurun_out
for P in 2 1; do
echo "== sanitizer PIPE=$P n=128"; LFMGPU_PIPE=$P timeout 400 compute-sanitizer --tool memcheck python -m lfm_public_b200.tools.tune --n 128 --steps 1 2>&1 | grep -v "^#" | head -40 | cut -c1-300
done
