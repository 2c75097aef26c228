#!/bin/bash
# Round-2 GPU call 2: operand-ring K3 (per-box zone waits, 64-row chains for B > 64): parity, timing, trace, ncu sections.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gru_tc.py -m gpu -x -q > $O/k3_small.log 2>&1; echo "k3 small rc=$?"; tail -3 $O/k3_small.log
timeout 900 python -m pytest tests/test_gpu_fullshape.py -m gpu -x -q -s -k "gru_tc" > $O/k3_full.log 2>&1; echo "k3 full rc=$?"; tail -3 $O/k3_full.log
python scratch/gru_time.py > $O/gru_time.txt 2>&1
B=128 TP=60 python scratch/gru_time.py >> $O/gru_time.txt 2>&1
B=256 TP=60 python scratch/gru_time.py >> $O/gru_time.txt 2>&1
NSD_GRU_WPC=2 python scratch/gru_time.py >> $O/gru_time.txt 2>&1
B=32 python scratch/gru_time.py >> $O/gru_time.txt 2>&1
cat $O/gru_time.txt
NSD_GRU_TRACE=1 timeout 120 python tests/trace_gru.py > /dev/null 2> $O/gru_trace.log; echo "trace rc=$?"; head -18 $O/gru_trace.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullshape.py::test_gru_tc_benchmark_shape > $O/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/gputests.log
timeout 300 python bench.py --breakdown > $O/bench_bi.json 2> $O/bench_bi.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_bi.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'])"; head -8 $O/bench_bi.err
timeout 600 python bench.py --T 2000 --batch 256 --steps 3 --warmup 3 --no-cpu-baseline --breakdown > $O/long_1.json 2> $O/long_1.err; echo "long rc=$?"; python -c "
import json; d=json.load(open('$O/long_1.json')); print(d['value'], d['ms_per_step'])"; head -6 $O/long_1.err
# K3 under Nsight Compute (non-cooperative cluster launch): first a 1-pass timing run, then sections without SASS patching
NSD_GRU_NO_COOP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gru_ --csv --log-file $O/ncu_k3_time.csv python tests/trace_gru.py > $O/ncu_k3_time.log 2>&1; echo "ncu k3 time rc=$?"; tail -5 $O/ncu_k3_time.csv
NSD_GRU_NO_COOP=1 timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none -k regex:gru_ -s 1 -c 2 -f -o $O/prof_k3 python tests/trace_gru.py > $O/ncu_k3.log 2>&1; echo "ncu k3 sections rc=$?"; tail -4 $O/ncu_k3.log
ls -la $O | head -40
