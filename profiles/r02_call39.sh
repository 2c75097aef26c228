#!/bin/bash
# Round-2 GPU call 39 (8 GPUs): weak-scaling line at 8 GPUs with the optimizer split under the all-reduce tail
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 --no-cpu-baseline > $O/final_scale_8.json 2>$O/final_scale_8.err; echo "N=8 rc=$?"; grep '^{' $O/final_scale_8.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('N=8', d['ms_per_step'], d['value'], d['e2e'])"; tail -3 $O/final_scale_8.err
