#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -k "rank or cfg5 or valuation or pooled or rows_next" 2>&1 | tail -8 > $OUT/pytest_rank_epilogue.log
timeout 300 python scripts/evalbench.py --model ComplEx,DistMult > $OUT/evalbench_tc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rank_tc" --launch-skip 3 --launch-count 1 -o $OUT/ncu_rank_tc_v2 \
    python scripts/evalbench.py --model ComplEx > $OUT/ncu_rank_tc_v2.log 2>&1
head -6 $OUT/evalbench_tc.log
