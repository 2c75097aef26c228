#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gru_tc.py -m gpu -x -q > $O/k_small.log 2>&1; echo "small rc=$?"; tail -2 $O/k_small.log | cut -c1-300
NSD_GRU_DEBUG=4 timeout 600 python -m pytest tests/test_gpu_gru_tc.py -m gpu -x -q > $O/k_small_weak.log 2>&1; echo "small weak rc=$?"; tail -2 $O/k_small_weak.log | cut -c1-300
{ echo "# relaxed.gpu LL loads"; python scratch/gru_time.py; B=128 TP=60 python scratch/gru_time.py; B=256 TP=60 python scratch/gru_time.py; B=32 python scratch/gru_time.py;
  echo "# weak .cg LL loads (NSD_GRU_DEBUG=4)"; NSD_GRU_DEBUG=4 python scratch/gru_time.py; NSD_GRU_DEBUG=4 B=256 TP=60 python scratch/gru_time.py;
  echo "# counter + TMA form (NSD_GRU_LL=0)"; NSD_GRU_LL=0 python scratch/gru_time.py; } > $O/gru_time.txt 2>&1
grep -v "timeout" $O/gru_time.txt | head -20
NSD_GRU_TRACE=1 timeout 120 python tests/trace_gru.py > /dev/null 2> $O/gru_trace.log; echo "trace rc=$?"; grep -v timeout $O/gru_trace.log | head -10
NSD_GRU_DEBUG=4 NSD_GRU_TRACE=1 timeout 120 python tests/trace_gru.py > /dev/null 2> $O/gru_trace_weak.log; grep -v timeout $O/gru_trace_weak.log | head -8
