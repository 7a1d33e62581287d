#!/bin/sh
# multi-GPU measurements of one box: tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=$1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN tools/pcie_ceiling.py 2>/dev/null | tail -1 > gpurun_out/pcie_ceiling_n$N.log; cut -c1-400 gpurun_out/pcie_ceiling_n$N.log
$RUN bench.py --gpus $N --steps 10 --warmup 3 --no-extra 2>gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_n$N.json; cut -c1-300 gpurun_out/bench_n$N.json; tail -2 gpurun_out/bench_n$N.err
$RUN bench.py --gpus $N --workload mixed --steps 3 --warmup 1 2>gpurun_out/mixed_n$N.err | tail -1 > gpurun_out/mixed_n$N.json; cut -c1-200 gpurun_out/mixed_n$N.json; tail -2 gpurun_out/mixed_n$N.err
