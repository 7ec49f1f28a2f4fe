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
true
true
ls -la $OUT | tail -8
timeout 200 python bench.py --config cfg3 --pool reference --pooled-gemm --no-cpu-baseline --no-hbm-config --steps 100 --warmup 10 \
    > $OUT/bench_cfg3_pool_gemm_final.json 2> $OUT/bench_cfg3_pool_gemm_final.err
