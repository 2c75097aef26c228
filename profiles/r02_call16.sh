#!/bin/bash
# Round-2 GPU call 16: ncu of the batched attention GEMM (register double-buffered form)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bgemm_kernel" -s 10 -c 6 -f -o $O/prof_bgemm python bench.py --mode conformer --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_bgemm.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_bgemm.log | cut -c1-200
