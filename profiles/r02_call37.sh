#!/bin/bash
# Round-2 GPU call 37: validation at HEAD -- full GPU suite, smoke(), the default bench line (with cpu_baseline), then the profile capture
# (launch list incl. K3 in non-cooperative mode, ncu --set full of the K2 GEMMs and of K1 / Adam / CTC) on commands that exited 0 without ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/final_gputests.log 2>&1; echo "tests rc=$?"; tail -3 $O/final_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke $O/final_smoke.log
timeout 600 python bench.py --breakdown > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; grep '^{' $O/final_bench.json | cut -c1-400; grep calls $O/final_bench.err | head -12
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
NSD_GRU_NO_COOP=1 $CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
NSD_GRU_NO_COOP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/final_launches.csv $CMD > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc2?_kernel" -s 69 -c 3 -f -o $O/final_prof_gemm $CMD > $O/ncu_gemm.log 2>&1; echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:frontend_fwd|frontend_bwd|adam_kernel|ctc_kernel' -s 12 -c 4 -f -o $O/final_prof_misc $CMD > $O/ncu_misc.log 2>&1; echo "misc capture rc=$?"
ls -la $O | grep -E "final_"
