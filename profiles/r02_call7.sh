#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gru_tc.py tests/test_gpu_streaming.py -m gpu -x -q > $O/k_small.log 2>&1; echo "small rc=$?"; tail -2 $O/k_small.log | cut -c1-300
{ python scratch/gru_time.py; B=128 TP=60 python scratch/gru_time.py; B=256 TP=60 python scratch/gru_time.py; B=32 python scratch/gru_time.py; } > $O/gru_time.txt 2>&1
grep -v "timeout" $O/gru_time.txt | head -20
NSD_GRU_TRACE=1 timeout 120 python tests/trace_gru.py > /dev/null 2> $O/gru_trace.log; echo "trace rc=$?"; grep -v timeout $O/gru_trace.log | head -8; grep -A8 "gru_bwd_bf16 trace" $O/gru_trace.log | head -9
timeout 600 python -m pytest tests/test_gpu_fullshape.py -m gpu -x -q -k "gru_tc" > $O/k3_full.log 2>&1; echo "k3 full rc=$?"; tail -2 $O/k3_full.log | cut -c1-200
timeout 300 python bench.py --breakdown > $O/bench_bi.json 2> $O/bench_bi.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_bi.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; head -4 $O/bench_bi.err
timeout 600 python bench.py --T 2000 --batch 256 --steps 3 --warmup 3 --no-cpu-baseline --breakdown > $O/long_1.json 2> $O/long_1.err; echo "long rc=$?"; python -c "
import json; d=json.load(open('$O/long_1.json')); print(d['value'], d['ms_per_step'])"; head -4 $O/long_1.err
