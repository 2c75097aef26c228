#!/bin/bash
# Round-2 GPU call 11 (8 GPUs of one box): weak scaling 2/4/8, NCCL CTA cap + GEMM SM reserve variants at 8, long-sequence config at 2/4/8, NCCL DP equivalence
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $n --no-cpu-baseline "$@"; }
timeout 300 python bench.py --no-cpu-baseline > $O/scale_1.json 2>/dev/null; echo "n=1 rc=$?"
for n in 2 4 8; do run $n > $O/scale_$n.json 2> $O/scale_$n.err; echo "n=$n rc=$?"; done
run 8 --nccl-ctas 8 > $O/scale_8_ctas8.json 2>/dev/null; echo "ctas8 rc=$?"
run 8 --nccl-ctas 16 > $O/scale_8_ctas16.json 2>/dev/null; echo "ctas16 rc=$?"
NCCL_DEBUG=INFO run 8 --steps 3 > /dev/null 2> $O/nccl_info_8.log; grep -E "NVLS|Using network|Channel 00/|nChannels|via P2P|comm .* nranks" $O/nccl_info_8.log | head -12 > $O/nccl_summary_8.txt; rm -f $O/nccl_info_8.log
for n in 1 2 4 8; do if [ $n = 1 ]; then timeout 600 python bench.py --no-cpu-baseline --T 2000 --batch 256 --steps 3 --warmup 2 > $O/long_1.json 2>/dev/null; else run $n --T 2000 --batch 256 --steps 3 --warmup 2 > $O/long_$n.json 2>/dev/null; fi; echo "long n=$n rc=$?"; done
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -s -k nccl > $O/dp_nccl.log 2>&1; echo "dp nccl rc=$?"; grep -E "passed|failed|EQUIV|rel" $O/dp_nccl.log | head -5
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_*.json')+glob.glob('gpurun_out/long_*.json')):
    try:
        d=json.load(open(f)); print(f, d['n_gpus'], d['ms_per_step'], round(d['value'],1), d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
