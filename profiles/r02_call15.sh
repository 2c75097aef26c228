#!/bin/bash
# Round-2 GPU call 15: whole GPU suite + smoke + headline bench at HEAD, Conformer bench after the depthwise-conv fix
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $O/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
timeout 300 python bench.py --breakdown > $O/bench_bi.json 2> $O/bench_bi.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$O/bench_bi.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'])"; grep nsd_ $O/bench_bi.err | head -12
timeout 600 python bench.py --mode conformer --breakdown --no-cpu-baseline > $O/conformer_bench.json 2> $O/conformer_bench.err; echo "conformer rc=$?"; python -c "
import json; d=json.load(open('$O/conformer_bench.json')); print(d['value'], d['ms_per_step'], d['host_enqueue_ms_per_step'], d['e2e']['value'])"; grep nsd_ $O/conformer_bench.err | head -12
