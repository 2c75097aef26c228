#!/bin/bash
# Round-2 GPU call 34: long-sequence config (T=2000, B=256) breakdown with four 32-row chains (default) vs 64-row chains (NSD_GRU_WPC=2), and with PDL off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --T 2000 --batch 256 --steps 4 --breakdown > $O/long_$tag.json 2> $O/long_$tag.err; grep '^{' $O/long_$tag.json | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$tag', d['ms_per_step'], d['value'], d['e2e']['value'])"; grep "calls" $O/long_$tag.err | head -8; }
run four NSD_GRU_WPC=0
run wpc2 NSD_GRU_WPC=2
run four_nopdl NSD_GRU_WPC=0 NSD_PDL=0
