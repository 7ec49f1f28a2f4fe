#!/usr/bin/env bash
# N-GPU call: the column-sharded (tensor-parallel) scheme against the replicated column-parallel one.
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s 2>&1 | tail -30 > $OUT/pytest_multi_${N}gpu.log
fi
for cfg in cfg4 cfg2; do
  timeout 300 $RUN bench.py --gpus $N --config $cfg --steps 100 --warmup 10 --mode colshard \
      > $OUT/bench_${cfg}_${N}gpu_colshard.json 2> $OUT/bench_${cfg}_${N}gpu_colshard.err
done
timeout 300 $RUN bench.py --gpus $N --config cfg4 --steps 100 --warmup 10 --mode colpar \
    > $OUT/bench_cfg4_${N}gpu_colpar_peer_b.json 2> $OUT/bench_cfg4_${N}gpu_colpar_peer_b.err
# the driver's own invocation (default flags, steps 20 / warmup 5): value and e2e must now agree
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_driverlike_${N}gpu_b.json 2> $OUT/bench_driverlike_${N}gpu_b.err
ls -la $OUT | tail -8
