#!/usr/bin/env bash
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -k "rank or cfg5 or valuation or rows_next or shard or smoke" 2>&1 | tail -6 > $OUT/pytest_rank_tile.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 300 python scripts/evalbench.py --model RotatE,TransE,ComplEx,DistMult > $OUT/evalbench_v3.log 2>&1
timeout 300 python scripts/evalbench.py --cfg5 > $OUT/evalbench_cfg5_v3.json 2> $OUT/evalbench_cfg5_v3.err
head -12 $OUT/evalbench_v3.log
