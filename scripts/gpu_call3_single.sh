#!/usr/bin/env bash
# 1-GPU call: the whole -m gpu suite (new: cfg5 at size, full-shape spot checks, ADVICE regressions), the
# default bench line (new roofline object, reference arm) and the reference arm on its own.
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -70 > $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
ls -la $OUT
