#!/bin/bash
# Round-2 first GPU call (1 GPU): smoke, full GPU parity suite, full bench line.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2a}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$? $(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/${tag}_smoke.log
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$? $(( $(date +%s) - t0 ))s"; grep -E "^\[e2e\]|passed|failed|Error|error" gpurun_out/${tag}_pytest.log | tail -40
timeout 900 python bench.py --steps 3 --warmup 3 --shapes-out gpurun_out/${tag}_gemm_shapes.md > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(( $(date +%s) - t0 ))s"; cut -c1-3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
