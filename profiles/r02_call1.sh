#!/bin/bash
# Round-2 GPU call 1: parity at the benchmarked shapes, measurement holes (uni, infer, stream, reference-cuda), K3 ncu counters.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 900 python -m pytest tests/test_gpu_fullshape.py -m gpu -x -q -s > $O/fullshape.log 2>&1; echo "fullshape rc=$?"
tail -3 $O/fullshape.log
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullshape.py > $O/gputests.log 2>&1; echo "gpu tests rc=$?"
tail -3 $O/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 300 python bench.py --breakdown > $O/bench_bi.json 2> $O/bench_bi.err; echo "bench rc=$?"; tail -c 1500 $O/bench_bi.json
timeout 300 python bench.py --uni --breakdown > $O/bench_uni.json 2> $O/bench_uni.err; echo "bench uni rc=$?"
timeout 400 python bench.py --impl reference --uni --steps 3 --warmup 1 > $O/ref_uni.json 2>/dev/null; echo "ref uni rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/ref_bi.json 2>/dev/null; echo "ref bi rc=$?"
timeout 300 python bench.py --impl reference-cuda --steps 10 --warmup 3 > $O/ref_cuda_bi.json 2> $O/ref_cuda_bi.err; echo "ref cuda rc=$?"; cat $O/ref_cuda_bi.json
timeout 300 python bench.py --impl reference-cuda --uni --steps 10 --warmup 3 > $O/ref_cuda_uni.json 2>/dev/null; echo "ref cuda uni rc=$?"
timeout 300 python bench.py --mode infer --steps 20 > $O/infer.jsonl 2> $O/infer.err; echo "infer rc=$?"; cat $O/infer.jsonl
timeout 300 python bench.py --mode stream > $O/stream.jsonl 2> $O/stream.err; echo "stream rc=$?"; cat $O/stream.jsonl
python scratch/gru_time.py > $O/gru_time.txt 2>&1; cat $O/gru_time.txt
# K3 under Nsight Compute: non-cooperative cluster launch (NSD_GRU_NO_COOP=1), one forward + one BPTT launch at T'=20
NSD_GRU_NO_COOP=1 timeout 120 python tests/trace_gru.py > $O/k3_nocoop_plain.log 2>&1; echo "k3 nocoop plain rc=$?"
NSD_GRU_NO_COOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_ -s 2 -c 1 -f -o $O/prof_k3_fwd python tests/trace_gru.py > $O/ncu_k3_fwd.log 2>&1; echo "ncu k3 fwd rc=$?"; tail -3 $O/ncu_k3_fwd.log
NSD_GRU_NO_COOP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gru_bwd -s 1 -c 1 -f -o $O/prof_k3_bwd python tests/trace_gru.py > $O/ncu_k3_bwd.log 2>&1; echo "ncu k3 bwd rc=$?"; tail -3 $O/ncu_k3_bwd.log
ls -la $O
