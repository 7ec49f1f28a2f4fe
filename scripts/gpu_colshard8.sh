#!/usr/bin/env bash
set -u
N=${1:-8}
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for cfg in cfg2 cfg4; do
  timeout 300 $RUN bench.py --gpus $N --config $cfg --steps 100 --warmup 10 --mode colshard \
      > $OUT/bench_${cfg}_${N}gpu_colshard.json 2> $OUT/bench_${cfg}_${N}gpu_colshard.err
done
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_driverlike_${N}gpu_c.json 2> $OUT/bench_driverlike_${N}gpu_c.err
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 --mode colshard > $OUT/bench_driverlike_${N}gpu_colshard.json 2> $OUT/bench_driverlike_${N}gpu_colshard.err
ls -la $OUT | tail -6
