#!/bin/bash
# parity tests + quick bench + ncu full capture at half scale: tools/gpucheck.sh <tag>
T=${1:-x}
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_$T.log; tail -4 gpurun_out/pytest_gpu_$T.log
timeout 200 python tools/quickbench.py uniform3d 1.0 2>&1 | tail -3 | head -1 | cut -c1-330
timeout 300 ncu --set full --clock-control none --import-source on -k regex:push_deposit --launch-skip 10 --launch-count 1 -o gpurun_out/prof_push_$T -f python tools/quickbench.py uniform3d 0.5 > gpurun_out/ncu_$T.log 2>&1
