#!/bin/bash
# GPU parity suite only (no -x), compact report.
tag=${1:-r2t}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -s ${2:-} > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; grep -E "^\[e2e\]|passed|failed|^E  |^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -60
