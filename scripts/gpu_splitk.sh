#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -k "pooled or rank or cfg5 or Evaluation or evaluation" 2>&1 | tail -8 > $OUT/pytest_splitk.log
B="python bench.py --config cfg3 --pool reference --pooled-gemm --no-cpu-baseline --no-hbm-config --steps 100 --warmup 10"
KGE_DOT_SPLITK=0 timeout 200 $B > $OUT/bench_cfg3_pool_gemm_nosplit.json 2> $OUT/bench_cfg3_pool_gemm_nosplit.err
timeout 200 $B > $OUT/bench_cfg3_pool_gemm_splitk.json 2> $OUT/bench_cfg3_pool_gemm_splitk.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rank_tc|pooled|transpose" --launch-skip 100 --launch-count 24 --csv --log-file $OUT/launches_pooled_splitk.csv $B --steps 3 > /dev/null 2>&1
ls $OUT | tail -5
