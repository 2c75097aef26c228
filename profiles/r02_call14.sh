#!/bin/bash
# Round-2 GPU call 14: ncu of the Conformer element-wise kernels that are far from their HBM floor
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dwconv_fwd_kernel|layernorm_bwd_kernel|dwconv_bwd_w_kernel|bgemm_kernel|residual_kernel|colsum" -s 40 -c 14 -f -o $O/prof_conformer_ew python bench.py --mode conformer --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_conformer_ew.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu_conformer_ew.log
