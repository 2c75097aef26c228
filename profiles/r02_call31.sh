#!/bin/bash
# Round-2 GPU call 31: ncu --set full of the fast front-end forward (first launch = benchmark shape with noise), with source-level counters
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:frontend_fwd_fast -c 1 -o $O/k1_fast -f python scratch/frontend_ab.py child > $O/ncu_k1_fast.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu_k1_fast.log; ls -la $O/k1_fast.ncu-rep
