#!/bin/bash
# 1 GPU: peer / gemm tests (simulated ranks) incl. the fused scatter.
tag=${1:-r2e}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_peer_gpu.py tests/test_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; grep -aE "passed|failed|^E  |^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -20
