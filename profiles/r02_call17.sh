#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conformer_kernels.py -m gpu -q > $O/conformer_kernel_tests.log 2>&1; echo "rc=$?"; tail -25 $O/conformer_kernel_tests.log | cut -c1-220
