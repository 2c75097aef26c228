#!/bin/bash
# Round-2 GPU call 43: validation at HEAD (optimizer under the recurrence on): full GPU suite, smoke(), the default bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/final_gputests.log 2>&1; echo "tests rc=$?"; tail -3 $O/final_gputests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke $O/final_smoke.log
timeout 300 python bench.py --breakdown > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; grep '^{' $O/final_bench.json | cut -c1-300; grep calls $O/final_bench.err | head -12
