#!/bin/bash
# one GPU session: parity tests, the bench line, 2D workload, dram traffic of a bench-size push launch, launch list
# usage: tools/gpu_round.sh <tag>
T=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_$T.log; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 900 python bench.py > gpurun_out/bench_n1_$T.json 2> gpurun_out/bench_n1_$T.err; cut -c1-1500 gpurun_out/bench_n1_$T.json
timeout 300 python tools/quickbench.py weibel2d 1.0 2>&1 | tail -2 | cut -c1-400
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:push_deposit --launch-skip 5 --launch-count 1 --csv --log-file gpurun_out/traffic_$T.csv python tools/quickbench.py uniform3d 1.0 > gpurun_out/traffic_$T.log 2>&1; tail -4 gpurun_out/traffic_$T.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/launches_$T.log 2>&1; tail -2 gpurun_out/launches_$T.csv | cut -c1-300
