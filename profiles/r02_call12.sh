#!/bin/bash
# Round-2 GPU call 12: first run of the Conformer path (configs[2]) against the reference fixtures and the oracle port
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_conformer.py -m gpu -q -s > $O/conformer_tests.log 2>&1; echo "conformer tests rc=$?"
grep -E "conformer_|competition|passed|failed|Error|error|assert" $O/conformer_tests.log | head -40
tail -30 $O/conformer_tests.log | cut -c1-250
