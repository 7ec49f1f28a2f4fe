#!/usr/bin/env bash
# N-GPU call: handshake A/B of the column-parallel step (NVLink peer flags vs NCCL), config 2 and 4.
set -u
N=${1:-2}
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -s 2>&1 | tail -30 > $OUT/pytest_multi_${N}gpu.log
fi
nvidia-smi topo -m > $OUT/topo_${N}gpu.txt 2>&1
for cfg in cfg2 cfg4; do
  for hs in peer nccl; do
    if [ "$N" -ge 4 ] && [ $cfg = cfg4 ] && [ $hs = nccl ]; then continue; fi
    timeout 300 $RUN bench.py --gpus $N --config $cfg --steps 100 --warmup 10 --mode colpar --handshake $hs \
        > $OUT/bench_${cfg}_${N}gpu_colpar_${hs}.json 2> $OUT/bench_${cfg}_${N}gpu_colpar_${hs}.err
  done
done
# the driver's own invocation (default flags, steps 20 / warmup 5)
timeout 300 $RUN bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_driverlike_${N}gpu.json 2> $OUT/bench_driverlike_${N}gpu.err
if [ "$N" -eq 4 ]; then
  for mode in rowshard; do
    nvidia-smi nvlink -gt d -i 0 > $OUT/nvlink_cfg4_${N}gpu_${mode}_before.txt 2>&1
    timeout 300 $RUN bench.py --gpus $N --config cfg4 --steps 100 --warmup 10 --mode $mode \
        > $OUT/bench_cfg4_${N}gpu_${mode}.json 2> $OUT/bench_cfg4_${N}gpu_${mode}.err
    nvidia-smi nvlink -gt d -i 0 > $OUT/nvlink_cfg4_${N}gpu_${mode}_after.txt 2>&1
  done
fi
ls -la $OUT
