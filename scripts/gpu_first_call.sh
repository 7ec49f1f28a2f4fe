#!/usr/bin/env bash
# First GPU call of a round: everything that was written without a GPU gets validated and measured in ONE
# box acquisition (1 GPU).  Usage:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_first_call.sh'
# Outputs land in gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/rNN/.
set -u
mkdir -p gpurun_out
OUT=gpurun_out
export PYTHONUNBUFFERED=1

# 1. parity: the whole -m gpu suite (new files: test_gpu_rows_next, test_gpu_sharded, test_gpu_train_byent)
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $OUT/pytest_gpu.log  # no -x: the whole picture in one box acquisition
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1

# 2. bench: measured default vs the atomics-free by-entity backward, all four training configs
for cfg in cfg2 cfg1 cfg3 cfg4; do
  timeout 300 python bench.py --config $cfg --steps 200 --warmup 10 $([ $cfg = cfg2 ] || echo --no-cpu-baseline) \
      > $OUT/bench_${cfg}_scatter.json 2> $OUT/bench_${cfg}_scatter.err
  timeout 300 python bench.py --config $cfg --steps 200 --warmup 10 --no-cpu-baseline --backward by_entity \
      > $OUT/bench_${cfg}_byent.json 2> $OUT/bench_${cfg}_byent.err
done
# config 3 with the reference's shared pool: gather kernels vs the tensor-core GEMM formulation
timeout 300 python bench.py --config cfg3 --steps 200 --warmup 10 --no-cpu-baseline --pool reference \
    > $OUT/bench_cfg3_pool_gather.json 2> $OUT/bench_cfg3_pool_gather.err
timeout 300 python bench.py --config cfg3 --steps 200 --warmup 10 --no-cpu-baseline --pool reference --pooled-gemm \
    > $OUT/bench_cfg3_pool_gemm.json 2> $OUT/bench_cfg3_pool_gemm.err
# row-sharded kernels with all shards local (addressing overhead only; NVLink needs scripts/gpu_multi_call.sh)
timeout 300 python bench.py --config cfg4 --steps 100 --warmup 10 --no-cpu-baseline --virtual-shards 4 \
    > $OUT/bench_cfg4_vshard4.json 2> $OUT/bench_cfg4_vshard4.err
timeout 300 python scripts/evalbench.py > $OUT/evalbench.log 2>&1

# 3. ncu: launch list of a by-entity step, then full captures of its two heavy kernels and of top-k
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_byent.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --backward by_entity > $OUT/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:byent_apply_kernel -s 6 -c 1 \
    -o $OUT/ncu_byent_apply python bench.py --steps 3 --warmup 3 --no-cpu-baseline --backward by_entity \
    > $OUT/ncu_byent_apply.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_bwd_kernel -s 6 -c 1 \
    -o $OUT/ncu_dq_pass python bench.py --steps 3 --warmup 3 --no-cpu-baseline --backward by_entity \
    > $OUT/ncu_dq_pass.log 2>&1
# the default (scatter) step: one full capture of every kernel of a step (sampler, fwd, bwd, adam x2)
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 40 --launch-count 10 \
    -o $OUT/ncu_default_step python bench.py --steps 3 --warmup 8 --no-cpu-baseline \
    > $OUT/ncu_default_step.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --launch-skip 40 --launch-count 10 \
    -o $OUT/ncu_cfg4_step python bench.py --config cfg4 --steps 3 --warmup 8 --no-cpu-baseline \
    > $OUT/ncu_cfg4_step.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $OUT/nvidia_smi.csv
ls -la $OUT
