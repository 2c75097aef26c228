#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gru_tc.py tests/test_gpu_kernels.py -m gpu -x -q > $O/k_small.log 2>&1; echo "small rc=$?"; tail -3 $O/k_small.log
{ python scratch/gru_time.py; B=128 TP=60 python scratch/gru_time.py; B=256 TP=60 python scratch/gru_time.py; B=32 python scratch/gru_time.py;
  echo "# debug timing (wrong results): NSD_GRU_DEBUG=1 no zone waits; =2 relaxed publish; =3 both"; NSD_GRU_DEBUG=1 python scratch/gru_time.py; NSD_GRU_DEBUG=2 python scratch/gru_time.py; NSD_GRU_DEBUG=3 python scratch/gru_time.py; } > $O/gru_time.txt 2>&1
cat $O/gru_time.txt
NSD_GRU_TRACE=1 timeout 120 python tests/trace_gru.py > /dev/null 2> $O/gru_trace.log; echo "trace rc=$?"; head -8 $O/gru_trace.log
NSD_GRU_DEBUG=3 NSD_GRU_TRACE=1 timeout 120 python tests/trace_gru.py > /dev/null 2> $O/gru_trace_nosync.log; head -8 $O/gru_trace_nosync.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/gputests.log
timeout 300 python bench.py --breakdown > $O/bench_bi.json 2> $O/bench_bi.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_bi.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'])"; head -8 $O/bench_bi.err
NSD_K1_BWD_V1=1 timeout 300 python bench.py --breakdown --no-cpu-baseline > $O/bench_bi_k1v1.json 2> $O/bench_bi_k1v1.err; grep frontend $O/bench_bi_k1v1.err
timeout 600 python bench.py --T 2000 --batch 256 --steps 3 --warmup 3 --no-cpu-baseline --breakdown > $O/long_1.json 2> $O/long_1.err; echo "long rc=$?"; python -c "
import json; d=json.load(open('$O/long_1.json')); print(d['value'], d['ms_per_step'])"; head -6 $O/long_1.err
