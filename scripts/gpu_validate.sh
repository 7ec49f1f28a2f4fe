#!/usr/bin/env bash
# 1-GPU validation call: the whole -m gpu suite, smoke, the default bench line (what the driver runs), the
# reference arm, the config-5 evaluation line, ncu of the ranking kernels.
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
timeout 300 python scripts/evalbench.py --cfg5 > $OUT/evalbench_cfg5.json 2> $OUT/evalbench_cfg5.err
timeout 300 python scripts/evalbench.py > $OUT/evalbench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_tc -c 2 -o $OUT/ncu_rank_tc \
    python scripts/evalbench.py --model ComplEx > $OUT/ncu_rank_tc.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"sample_negatives|topk_rows|kl_fwd|kl_bwd|adam_kernel" -c 6 -o $OUT/ncu_small \
    python -m pytest tests/test_gpu_rows_next.py tests/test_gpu_parity.py -q -k "kl or topk or TopK or sampler_invariants or dense_adam" > $OUT/ncu_small.log 2>&1
ls -la $OUT | tail -8
