#!/bin/bash
# Round-2 GPU call 26 (2 GPUs): Adam split under the last all-reduce (finish_early): split-step test, 2-rank NCCL equivalence, 2-GPU bench lines + timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "adam" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -s > $O/dp_equiv.log 2>&1; echo "dp rc=$?"; grep -E "rel|passed|failed|Error" $O/dp_equiv.log | head
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --no-cpu-baseline "$@" > $O/scale_2_$tag.json 2>$O/scale_2_$tag.err; echo "$tag rc=$?"; grep '^{' $O/scale_2_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$tag', d['ms_per_step'], d['value'], d['e2e'])"; }
run split
run split_tl --timeline $O/tl_split.json
timeout 300 python bench.py --no-cpu-baseline > $O/bench_1.json 2>$O/bench_1.err; grep '^{' $O/bench_1.json | cut -c1-300
run split2
