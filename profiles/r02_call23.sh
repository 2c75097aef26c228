#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd_kernel|cast_colsum_kernel|act_bwd_kernel|softmax_mask_bwd|dwconv_bwd_w_kernel" -s 4 -c 10 -f -o $O/prof_conformer_ln python bench.py --mode conformer --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_conformer_ln.log 2>&1; echo "ncu rc=$?"
