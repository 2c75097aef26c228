#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_conformer.py tests/test_gpu_conformer_kernels.py -m gpu -q > $O/conformer_tests.log 2>&1; echo "conformer tests rc=$?"; grep -E "passed|failed" $O/conformer_tests.log
timeout 600 python bench.py --mode conformer --graph --no-cpu-baseline > $O/conformer_bench_graph.json 2> $O/conformer_bench_graph.err; echo "graph bench rc=$?"; cut -c1-260 $O/conformer_bench_graph.json
timeout 600 python bench.py --mode conformer --no-cpu-baseline --breakdown > $O/conformer_bench.json 2> $O/conformer_bench.err; echo "bench rc=$?"; cut -c1-260 $O/conformer_bench.json; grep -E "nsd_|sum of" $O/conformer_bench.err | head -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd_kernel|cast_colsum_kernel|act_bwd_kernel|softmax_mask_bwd|dwconv_bwd_w_kernel" -s 4 -c 10 -f -o $O/prof_conformer_ln python bench.py --mode conformer --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_conformer_ln.log 2>&1; echo "ncu rc=$?"
