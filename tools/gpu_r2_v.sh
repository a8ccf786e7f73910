#!/bin/bash
tag=${1:-r2v}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vae_gpu.py -m gpu -q -s -x > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; grep -aE "\[vae\]|passed|failed|^E  |Error" gpurun_out/${tag}_pytest.log | tail -20
