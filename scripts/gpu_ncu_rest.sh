#!/usr/bin/env bash
# 1-GPU call: ncu --set full of the kernels that had no capture yet (pooled GEMM flow, top-k, KL, row-sharded
# variants with local shards, the tensor-core ranking at full size).
set -u
mkdir -p gpurun_out; OUT=gpurun_out; export PYTHONUNBUFFERED=1
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:"pooled|dot_nt|rank_tc|transpose|gather_pool" --launch-skip 60 --launch-count 14 -o $OUT/ncu_pooled \
    python bench.py --config cfg3 --pool reference --pooled-gemm --steps 3 --warmup 6 --no-cpu-baseline --no-hbm-config > $OUT/ncu_pooled.log 2>&1
timeout 300 $NCU -k regex:"score_neg_kernel|score_bwd_kernel" --launch-skip 12 --launch-count 2 -o $OUT/ncu_vshard4 \
    python bench.py --config cfg4 --virtual-shards 4 --steps 3 --warmup 6 --no-cpu-baseline --no-hbm-config > $OUT/ncu_vshard4.log 2>&1
cat > /tmp/small_kernels.py <<'PY'
import torch, sys
sys.path.insert(0, ".")
from mkb_b200 import ops, losses
x = torch.randn(1024, 40943, device="cuda")
for _ in range(2):
    ops.topk_rows(x, 64)
s = torch.randn(1024, 256, device="cuda", requires_grad=True); t = torch.randn(1024, 256, device="cuda")
for _ in range(2):
    l = losses.KlDivergence()(s, t, T=3); l.backward()
torch.cuda.synchronize()
PY
timeout 300 $NCU -k regex:"topk_rows|kl_fwd|kl_bwd" --launch-count 6 -o $OUT/ncu_topk_kl python /tmp/small_kernels.py > $OUT/ncu_topk_kl.log 2>&1
timeout 300 $NCU -k regex:"rank_tc|rank_tile" --launch-skip 2 --launch-count 2 -o $OUT/ncu_rank_full \
    python scripts/evalbench.py --model ComplEx,RotatE --Q 1024 > $OUT/ncu_rank_full.log 2>&1
ls -la $OUT/*.ncu-rep | tail
