#!/bin/bash
# Round-2 GPU call 36: ncu --set full + source counters of ctc_kernel (which phase holds the 155 us)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
NSD_GRU_NO_COOP=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:ctc_kernel -s 3 -c 1 -f -o $O/prof_ctc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_ctc.log 2>&1; echo "rc=$?"; ls -la $O/prof_ctc.ncu-rep
