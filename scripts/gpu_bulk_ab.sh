#!/usr/bin/env bash
# 1-GPU call: K3b (LDG loads + staged bulk-reduction row gradients) against K3 (RED.128) — parity, A/B, ncu.
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_tma.py -x -q -k backward 2>&1 | tail -15 > $OUT/pytest_bulk.log
KGE_BWD_BULK=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullshape.py tests/test_gpu_advice.py -q 2>&1 | tail -8 > $OUT/pytest_parity_bulk.log
B="python bench.py --no-cpu-baseline --no-hbm-config --steps 100 --warmup 10"
for cfg in cfg2 cfg4 cfg3 cfg1; do
  KGE_BWD_BULK=0 timeout 200 $B --config $cfg > $OUT/ab2_${cfg}_red.json 2> $OUT/ab2_${cfg}_red.err
  KGE_BWD_BULK=1 timeout 200 $B --config $cfg > $OUT/ab2_${cfg}_bulk.json 2> $OUT/ab2_${cfg}_bulk.err
done
KGE_BWD_BULK=1 KGE_KS=1 timeout 200 $B --config cfg2 > $OUT/ab2_cfg2_bulk_ks1.json 2> $OUT/ab2_cfg2_bulk_ks1.err
KGE_BWD_BULK=1 KGE_KS=2 timeout 200 $B --config cfg2 > $OUT/ab2_cfg2_bulk_ks2.json 2> $OUT/ab2_cfg2_bulk_ks2.err
KGE_BWD_BULK=1 KGE_KS=8 timeout 200 $B --config cfg2 > $OUT/ab2_cfg2_bulk_ks8.json 2> $OUT/ab2_cfg2_bulk_ks8.err
KGE_BWD_BULK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_bwd_kernel -s 4 -c 1 \
    -o $OUT/ncu_bwd_bulk_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-config > $OUT/ncu_bwd_bulk.log 2>&1
KGE_BWD_BULK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:score_bwd_kernel -s 4 -c 1 \
    -o $OUT/ncu_bwd_bulk_cfg4 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-config > $OUT/ncu_bwd_bulk4.log 2>&1
ls -la $OUT | tail -5
