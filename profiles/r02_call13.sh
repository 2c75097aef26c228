#!/bin/bash
# Round-2 GPU call 13: Conformer tests at tightened tolerances + bench lines (ours bf16, breakdown; reference-cuda port)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_conformer.py -m gpu -q -s > $O/conformer_tests.log 2>&1; echo "conformer tests rc=$?"; grep -E "passed|failed" $O/conformer_tests.log
timeout 600 python bench.py --mode conformer --breakdown > $O/conformer_bench.json 2> $O/conformer_bench.err; echo "bench rc=$?"; cut -c1-900 $O/conformer_bench.json; grep -v Warn $O/conformer_bench.err | head -40
timeout 600 python bench.py --mode conformer --impl reference-cuda --steps 5 --warmup 2 > $O/conformer_ref_cuda.json 2> $O/conformer_ref_cuda.err; echo "ref-cuda rc=$?"; cut -c1-400 $O/conformer_ref_cuda.json; tail -3 $O/conformer_ref_cuda.err
