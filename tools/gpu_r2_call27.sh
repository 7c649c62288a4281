#!/bin/bash
# round 2, the last GPU seconds: parity of the generator fixtures after the EPI_MRF epilogue change of the per-layer kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 45 python -m pytest tests/test_mrf16_gpu.py tests/test_fused_gpu.py -q -x -m gpu -k "fixture or fused" > gpurun_out/r3e_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r3e_tests.log)"
