#!/usr/bin/env bash
# N-GPU sanity of the defaults: multi-GPU parity tests, the default scheme chosen by table size, the driver's invocation.
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = 2 ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -12 > $OUT/pytest_multi_${N}gpu_final.log
fi
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_final_${N}gpu_cfg2.json 2> $OUT/bench_final_${N}gpu_cfg2.err
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 --config cfg4 > $OUT/bench_final_${N}gpu_cfg4.json 2> $OUT/bench_final_${N}gpu_cfg4.err
timeout 300 $RUN bench.py --gpus $N --impl reference --steps 3 --warmup 1 --cpu-budget 20 > $OUT/bench_final_${N}gpu_reference.json 2> $OUT/bench_final_${N}gpu_reference.err
ls -la $OUT | tail -6
