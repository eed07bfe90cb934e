#!/bin/bash
# multi-GPU session (gpurun --gpus 8): N-rank parity tests, then the weak-scaling bench.  usage: tools/gpu_multi.sh <tag> [N ...]
T=${1:-x}; shift
NS=${@:-2 4 8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
(timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_multi_$T.log; tail -3 gpurun_out/pytest_multi_$T.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline > gpurun_out/scale_n1_$T.json 2> gpurun_out/scale_$T.err; cut -c1-180 gpurun_out/scale_n1_$T.json
for N in $NS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n${N}_$T.json 2>> gpurun_out/scale_$T.err
  tail -1 gpurun_out/scale_n${N}_$T.json | cut -c1-180
done
