#!/bin/bash
# Round-2 GPU call 27: K3 timing bounds with the debug switches (WRONG RESULTS by design; timing only): which link of the per-step chain costs what
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
for f in 0 4 8 12 2 6 14 1 15; do echo -n "NSD_GRU_DEBUG=$f  "; NSD_GRU_DEBUG=$f timeout 120 python scratch/gru_time.py 2>&1 | tail -1; done | tee $O/k3_bounds.log
