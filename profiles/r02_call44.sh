#!/bin/bash
# Round-2 GPU call 44: does the library notice Nsight Compute's injection and launch K3 non-cooperatively by itself?  (no NSD_GRU_NO_COOP set)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
env | sort > $O/env_plain.txt; timeout 60 ncu --metrics gpu__time_duration.sum env 2>/dev/null | sort > $O/env_ncu.txt; comm -13 $O/env_plain.txt $O/env_ncu.txt | cut -c1-120 | head -20
TP=8 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/ncu_autodetect.csv python scratch/gru_time.py > $O/ncu_autodetect.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_autodetect.log; grep -c "gru_.*_ts_kernel" $O/ncu_autodetect.csv
