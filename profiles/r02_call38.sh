#!/bin/bash
# Round-2 GPU call 38 (2 GPUs): data-parallel path at HEAD: 2-rank NCCL == 1-rank test, bench at 1 and 2 GPUs on the same box
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -s > $O/dp_equiv.log 2>&1; echo "dp rc=$?"; grep -E "worst|passed|failed|Error" $O/dp_equiv.log | head -4
timeout 300 python bench.py --no-cpu-baseline > $O/final_scale_1.json 2>$O/final_scale_1.err; grep '^{' $O/final_scale_1.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('N=1', d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 2 --no-cpu-baseline > $O/final_scale_2.json 2>$O/final_scale_2.err; echo "N=2 rc=$?"; grep '^{' $O/final_scale_2.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('N=2', d['ms_per_step'], d['value'], d['e2e'])"
