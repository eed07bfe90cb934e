#!/bin/bash
# one GPU session: parity tests, the bench line, other workloads, dram traffic of a bench-size push launch,
# launch list, one full ncu capture.   usage: tools/gpu_round.sh <tag>
T=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_$T.log; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 900 python bench.py > gpurun_out/bench_n1_$T.json 2> gpurun_out/bench_n1_$T.err; cut -c1-1200 gpurun_out/bench_n1_$T.json
timeout 600 python bench.py --prec f64 --e2e-steps 0 --no-cpu-baseline --steps 10 > gpurun_out/bench_n1_f64_$T.json 2>> gpurun_out/bench_n1_$T.err; cut -c1-400 gpurun_out/bench_n1_f64_$T.json
timeout 600 python bench.py --workload weibel2d --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_n1_weibel2d_$T.json 2>> gpurun_out/bench_n1_$T.err; cut -c1-400 gpurun_out/bench_n1_weibel2d_$T.json
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:push_deposit --launch-skip 10 --launch-count 1 --csv --log-file gpurun_out/traffic_$T.csv python tools/quickbench.py uniform3d 1.0 > gpurun_out/traffic_$T.log 2>&1; tail -3 gpurun_out/traffic_$T.csv | cut -c180-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/launches_$T.log 2>&1; tail -2 gpurun_out/launches_$T.csv | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:push_deposit --launch-skip 10 --launch-count 1 -o gpurun_out/prof_push_$T -f python tools/quickbench.py uniform3d 0.5 > gpurun_out/ncu_$T.log 2>&1
