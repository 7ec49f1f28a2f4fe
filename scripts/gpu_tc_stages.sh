#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
KGE_TC_STAGES=1 timeout 600 python -m pytest tests -m gpu -q -k "rank or cfg5 or valuation or pooled" 2>&1 | tail -5 > $OUT/pytest_tc_stages1.log
for v in 2 1; do
  KGE_TC_STAGES=$v timeout 200 python scripts/evalbench.py --model ComplEx,DistMult 2>&1 | grep -E "ComplEx|DistMult" > $OUT/evalbench_tc_stages$v.log
  KGE_TC_STAGES=$v timeout 200 python bench.py --config cfg3 --pool reference --pooled-gemm --no-cpu-baseline --no-hbm-config --steps 100 --warmup 10 > $OUT/bench_cfg3_pool_stages$v.json 2>/dev/null
done
KGE_TC_STAGES=1 timeout 300 ncu --set full --clock-control none -k regex:"rank_tc" --launch-skip 3 --launch-count 1 -o $OUT/ncu_rank_tc_stages1 \
    python scripts/evalbench.py --model ComplEx > /dev/null 2>&1
cat $OUT/evalbench_tc_stages2.log $OUT/evalbench_tc_stages1.log
