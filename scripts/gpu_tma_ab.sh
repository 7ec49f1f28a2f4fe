#!/usr/bin/env bash
# 1-GPU call: K2-TMA (rows staged through cp.async.bulk) against the LDG forward — parity first, then the A/B.
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tma.py -x -q 2>&1 | tail -15 > $OUT/pytest_tma.log
timeout 300 python -m pytest tests/test_gpu_cfg5.py -q -s 2>&1 | grep -E "cfg5|passed|failed" > $OUT/pytest_cfg5.log
B="python bench.py --no-cpu-baseline --no-hbm-config --steps 100 --warmup 10"
timeout 120 compute-sanitizer --tool memcheck python -m pytest "tests/test_gpu_tma.py" -x -q -k "DistMult or 33" 2>&1 | tail -12 > $OUT/sanitizer_tma.log
for cfg in cfg2 cfg4 cfg3 cfg1; do
  timeout 200 $B --config $cfg > $OUT/ab_${cfg}_ldg.json 2> $OUT/ab_${cfg}_ldg.err
  KGE_BWD_TMA=1 timeout 200 $B --config $cfg > $OUT/ab_${cfg}_bwdtma.json 2> $OUT/ab_${cfg}_bwdtma.err
  KGE_FWD_TMA=1 KGE_BWD_TMA=1 timeout 200 $B --config $cfg > $OUT/ab_${cfg}_bothtma.json 2> $OUT/ab_${cfg}_bothtma.err
  KGE_FWD_TMA=1 timeout 200 $B --config $cfg > $OUT/ab_${cfg}_tma_b1.json 2> $OUT/ab_${cfg}_tma_b1.err
  KGE_FWD_TMA=1 KGE_TMA_MINB=2 timeout 200 $B --config $cfg > $OUT/ab_${cfg}_tma_b2.json 2> $OUT/ab_${cfg}_tma_b2.err
done
KGE_FWD_TMA=1 KGE_TMA_STAGES=2 timeout 200 $B --config cfg2 > $OUT/ab_cfg2_tma_b1_s2.json 2> $OUT/ab_cfg2_tma_b1_s2.err
KGE_FWD_TMA=1 KGE_TMA_STAGES=1 timeout 200 $B --config cfg2 > $OUT/ab_cfg2_tma_b1_s1.json 2> $OUT/ab_cfg2_tma_b1_s1.err
KGE_FWD_TMA=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_neg_tma -s 4 -c 1 \
    -o $OUT/ncu_fwd_tma_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-config > $OUT/ncu_fwd_tma.log 2>&1
KGE_FWD_TMA=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_neg_tma -s 4 -c 1 \
    -o $OUT/ncu_fwd_tma_cfg4 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-config > $OUT/ncu_fwd_tma4.log 2>&1
KGE_BWD_TMA=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_bwd_tma -s 4 -c 1 \
    -o $OUT/ncu_bwd_tma_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-config > $OUT/ncu_bwd_tma.log 2>&1
KGE_BWD_TMA=1 KGE_TMA_STAGES=1 timeout 200 $B --config cfg2 > $OUT/ab_cfg2_bwdtma_s1.json 2> $OUT/ab_cfg2_bwdtma_s1.err
ls -la $OUT | tail -30
