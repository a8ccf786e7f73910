#!/bin/bash
# GPU parity suite only, compact report.
tag=${1:-r2t}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${tag}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; grep -aE "passed|failed|^E  |^FAILED|^ERROR" gpurun_out/${tag}_pytest.log | tail -8
