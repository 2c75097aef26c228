#!/bin/bash
# Round-2 profile capture (GPU box): after a plain run exited 0 --
#  1. launch list of the headline bench command INCLUDING the K3 cluster kernels (NSD_GRU_NO_COOP=1: non-cooperative cluster launch, the
#     form Nsight Compute can replay), --metrics gpu__time_duration.sum --clock-control none;
#  2. ncu --set full of the first three K2 launches of a timed step (traffic of the dominant kernel);
#  3. ncu --set full of K1 forward / backward, Adam, CTC.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; grep smoke $O/smoke.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
NSD_GRU_NO_COOP=1 $CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
python -c "
import json; d=json.loads([l for l in open('$O/plain.log') if l.startswith('{')][-1]); print('plain (non-cooperative K3):', d['value'], d['ms_per_step'])"
NSD_GRU_NO_COOP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc2?_kernel" -s 69 -c 3 -f -o $O/prof_gemm $CMD > $O/ncu_gemm.log 2>&1; echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:frontend_fwd_kernel|frontend_bwd|adam_kernel|ctc_kernel' -s 12 -c 4 -f -o $O/prof_misc $CMD > $O/ncu_misc.log 2>&1; echo "misc capture rc=$?"
ls -la $O | grep -E "launches|prof_gemm|prof_misc"
